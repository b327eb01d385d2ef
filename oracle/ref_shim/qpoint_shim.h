// Stand-ins for the deal.II types the reference's quadrature-point kernel touches
// (NavierStokesOperator::do_vmult_cell and symm_scalar_product_add, include/operator_ns.cc:880-1182), written for
// this repository (TEST INFRASTRUCTURE): Tensor, VectorizedArray (one lane), Table / AlignedVector, EvaluationFlags
// and an "FEEvaluation" whose evaluate / integrate are no-ops on per-point arrays handed in by the harness -- the
// sum-factorised evaluation, the geometry and the vector access are deal.II's and are NOT exercised here, only what
// the reference itself computes between get_value / get_gradient and submit_value / submit_gradient.
// deal.II is not available in this image; nothing of it is copied here, the semantics follow its documentation:
//   Tensor<2, dim> * Tensor<1, dim> contracts the last index of the first with the vector;
//   `T x = {}` and `Tensor<...>()` are zero;  scalar * Tensor and Tensor * scalar scale every entry.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <map>
#include <set>
#include <vector>

#define DEAL_II_ALWAYS_INLINE __attribute__((always_inline))

namespace dealii
{
  // ---- VectorizedArray with a single lane ----
  template <typename Number>
  struct VectorizedArray
  {
    Number data = Number(0);

    VectorizedArray() = default;
    VectorizedArray(const Number v)
      : data(v)
    {}

    Number
    operator[](const unsigned int) const
    {
      return data;
    }
    Number &
    operator[](const unsigned int)
    {
      return data;
    }

    VectorizedArray &
    operator+=(const VectorizedArray &o)
    {
      data += o.data;
      return *this;
    }
    VectorizedArray &
    operator-=(const VectorizedArray &o)
    {
      data -= o.data;
      return *this;
    }
    VectorizedArray &
    operator*=(const VectorizedArray &o)
    {
      data *= o.data;
      return *this;
    }
  };

#define GLSB_VA_BINARY(OP)                                                                       \
  template <typename Number>                                                                     \
  inline VectorizedArray<Number> operator OP(const VectorizedArray<Number> &a,                   \
                                             const VectorizedArray<Number> &b)                   \
  {                                                                                              \
    return VectorizedArray<Number>(a.data OP b.data);                                            \
  }                                                                                              \
  template <typename Number>                                                                     \
  inline VectorizedArray<Number> operator OP(const VectorizedArray<Number> &a, const double b)   \
  {                                                                                              \
    return VectorizedArray<Number>(a.data OP Number(b));                                         \
  }                                                                                              \
  template <typename Number>                                                                     \
  inline VectorizedArray<Number> operator OP(const double a, const VectorizedArray<Number> &b)   \
  {                                                                                              \
    return VectorizedArray<Number>(Number(a) OP b.data);                                         \
  }
  GLSB_VA_BINARY(+)
  GLSB_VA_BINARY(-)
  GLSB_VA_BINARY(*)
  GLSB_VA_BINARY(/)
#undef GLSB_VA_BINARY

  // ---- Tensor ----
  template <int rank, int dim, typename Number>
  class Tensor;

  template <int dim, typename Number>
  class Tensor<1, dim, Number>
  {
  public:
    Number v[dim] = {};

    Number &
    operator[](const unsigned int i)
    {
      return v[i];
    }
    const Number &
    operator[](const unsigned int i) const
    {
      return v[i];
    }
    Tensor &
    operator+=(const Tensor &o)
    {
      for (int i = 0; i < dim; ++i)
        v[i] += o.v[i];
      return *this;
    }
    Tensor &
    operator-=(const Tensor &o)
    {
      for (int i = 0; i < dim; ++i)
        v[i] -= o.v[i];
      return *this;
    }
    // l2 norm (used on Tensor<1, dim, VectorizedArray<Number>> by compute_penalty_parameters)
    Number
    norm() const
    {
      Number s = v[0] * v[0];
      for (int i = 1; i < dim; ++i)
        s += v[i] * v[i];
      return Number(std::sqrt(s.data));
    }
  };

  template <int dim, typename Number>
  class Tensor<2, dim, Number>
  {
  public:
    Tensor<1, dim, Number> v[dim] = {};

    Tensor<1, dim, Number> &
    operator[](const unsigned int i)
    {
      return v[i];
    }
    const Tensor<1, dim, Number> &
    operator[](const unsigned int i) const
    {
      return v[i];
    }
    Tensor &
    operator+=(const Tensor &o)
    {
      for (int i = 0; i < dim; ++i)
        v[i] += o.v[i];
      return *this;
    }
  };

  template <int dim, typename Number>
  inline Tensor<1, dim, Number>
  operator+(const Tensor<1, dim, Number> &a, const Tensor<1, dim, Number> &b)
  {
    Tensor<1, dim, Number> r;
    for (int i = 0; i < dim; ++i)
      r[i] = a[i] + b[i];
    return r;
  }

  // scalar * Tensor and Tensor * scalar (S = VectorizedArray<Number> or a plain number)
  template <int dim, typename Number, typename S>
  inline Tensor<1, dim, Number>
  scaled(const Tensor<1, dim, Number> &t, const S &s)
  {
    Tensor<1, dim, Number> r;
    for (int i = 0; i < dim; ++i)
      r[i] = t[i] * s;
    return r;
  }
  template <int dim, typename Number>
  inline Tensor<1, dim, Number>
  operator*(const Number &s, const Tensor<1, dim, Number> &t)
  {
    return scaled(t, s);
  }
  template <int dim, typename Number>
  inline Tensor<1, dim, Number>
  operator*(const Tensor<1, dim, Number> &t, const Number &s)
  {
    return scaled(t, s);
  }
  template <int dim, typename Number>
  inline Tensor<1, dim, Number>
  operator*(const Tensor<1, dim, Number> &t, const double s)
  {
    return scaled(t, s);
  }
  template <int dim, typename Number>
  inline Tensor<2, dim, Number>
  operator*(const Number &s, const Tensor<2, dim, Number> &t)
  {
    Tensor<2, dim, Number> r;
    for (int i = 0; i < dim; ++i)
      r[i] = scaled(t[i], s);
    return r;
  }

  // Tensor<1> * Tensor<1>: scalar product
  template <int dim, typename Number>
  inline Number
  operator*(const Tensor<1, dim, Number> &a, const Tensor<1, dim, Number> &b)
  {
    Number s = a[0] * b[0];
    for (int i = 1; i < dim; ++i)
      s += a[i] * b[i];
    return s;
  }

  // Tensor<2> * Tensor<1>: (A * b)[i] = sum_j A[i][j] b[j]
  template <int dim, typename Number>
  inline Tensor<1, dim, Number>
  operator*(const Tensor<2, dim, Number> &A, const Tensor<1, dim, Number> &b)
  {
    Tensor<1, dim, Number> r;
    for (int i = 0; i < dim; ++i)
      {
        Number s = A[i][0] * b[0];
        for (int j = 1; j < dim; ++j)
          s += A[i][j] * b[j];
        r[i] = s;
      }
    return r;
  }

  // ---- containers ----
  template <typename T>
  using AlignedVector = std::vector<T>;

  template <int N, typename T>
  class Table;

  template <typename T>
  class Table<2, T>
  {
  public:
    std::size_t    n0 = 0, n1 = 0;
    std::vector<T> data;

    void
    reinit(const std::size_t a, const std::size_t b)
    {
      n0 = a;
      n1 = b;
      data.assign(a * b, T());
    }
    std::size_t
    size(const unsigned int i) const
    {
      return i == 0 ? n0 : n1;
    }
    T *
    operator[](const std::size_t i)
    {
      return data.data() + i * n1;
    }
    const T *
    operator[](const std::size_t i) const
    {
      return data.data() + i * n1;
    }
  };

  template <typename T>
  class Table<1, T>
  {
  public:
    std::vector<T> data;
    void
    reinit(const std::size_t a)
    {
      data.assign(a, T());
    }
    T &
    operator[](const std::size_t i)
    {
      return data[i];
    }
    const T &
    operator[](const std::size_t i) const
    {
      return data[i];
    }
  };

  // ---- FEEvaluation: per-point arrays in, per-point arrays out ----
  namespace EvaluationFlags
  {
    enum EvaluationFlags
    {
      nothing   = 0,
      values    = 1,
      gradients = 2
    };
    inline EvaluationFlags
    operator|(const EvaluationFlags a, const EvaluationFlags b)
    {
      return static_cast<EvaluationFlags>(static_cast<int>(a) | static_cast<int>(b));
    }
  } // namespace EvaluationFlags

  template <int dim, int fe_degree, int n_q_points_1d, int n_components, typename Number>
  class FEEvaluation
  {
  public:
    using value_type    = Tensor<1, n_components, VectorizedArray<Number>>;
    using gradient_type = Tensor<1, n_components, Tensor<1, dim, VectorizedArray<Number>>>;

    std::vector<value_type>    values_in, values_out;
    std::vector<gradient_type> gradients_in, gradients_out;
    unsigned int               cell = 0;

    unsigned int
    get_current_cell_index() const
    {
      return cell;
    }
    void
    evaluate(const EvaluationFlags::EvaluationFlags)
    {}
    void
    integrate(const EvaluationFlags::EvaluationFlags)
    {}
    std::vector<unsigned int>
    quadrature_point_indices() const
    {
      std::vector<unsigned int> r(values_in.size());
      for (unsigned int i = 0; i < r.size(); ++i)
        r[i] = i;
      return r;
    }
    value_type
    get_value(const unsigned int q) const
    {
      return values_in[q];
    }
    gradient_type
    get_gradient(const unsigned int q) const
    {
      return gradients_in[q];
    }
    void
    submit_value(const value_type &v, const unsigned int q)
    {
      values_out[q] = v;
    }
    void
    submit_gradient(const gradient_type &g, const unsigned int q)
    {
      gradients_out[q] = g;
    }
  };
  // boundary faces: the same, plus the normal vectors, the boundary id and the dof values the reference zeroes
  // on faces without outflow terms
  template <int dim, int fe_degree, int n_q_points_1d, int n_components, typename Number>
  class FEFaceEvaluation : public FEEvaluation<dim, fe_degree, n_q_points_1d, n_components, Number>
  {
  public:
    std::vector<Tensor<1, dim, VectorizedArray<Number>>> normals;
    std::vector<VectorizedArray<Number>>                 dof_values;
    unsigned int                                         id            = 0;
    unsigned int                                         dofs_per_cell = 0;

    unsigned int
    boundary_id() const
    {
      return id;
    }
    Tensor<1, dim, VectorizedArray<Number>>
    get_normal_vector(const unsigned int q) const
    {
      return normals[q];
    }
    VectorizedArray<Number> *
    begin_dof_values()
    {
      return dof_values.data();
    }
  };
} // namespace dealii

namespace dealii
{
  namespace Utilities
  {
    template <int N, typename T>
    inline T
    fixed_power(const T t)
    {
      static_assert(N == 2, "only squares are needed");
      return t * t;
    }
  } // namespace Utilities
} // namespace dealii

// deal.II overloads these two for VectorizedArray (lane-wise)
namespace std
{
  template <typename Number>
  inline dealii::VectorizedArray<Number>
  sqrt(const dealii::VectorizedArray<Number> &x)
  {
    return dealii::VectorizedArray<Number>(std::sqrt(x.data));
  }
  template <typename Number>
  inline dealii::VectorizedArray<Number>
  max(const dealii::VectorizedArray<Number> &a, const dealii::VectorizedArray<Number> &b)
  {
    return dealii::VectorizedArray<Number>(a.data > b.data ? a.data : b.data);
  }
  template <typename Number>
  inline dealii::VectorizedArray<Number>
  pow(const dealii::VectorizedArray<Number> &x, const Number p)
  {
    return dealii::VectorizedArray<Number>(std::pow(x.data, p));
  }
  template <typename Number>
  inline dealii::VectorizedArray<Number>
  min(const dealii::VectorizedArray<Number> &a, const dealii::VectorizedArray<Number> &b)
  {
    return dealii::VectorizedArray<Number>(a.data < b.data ? a.data : b.data);
  }
} // namespace std
