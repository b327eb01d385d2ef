// Minimal stand-in for the slice of the deal.II API that dealii_ns_gls_b200/cpp/dealii_adapter.h uses.
// PURPOSE: syntax, type and linkage check of the adapter in an image without deal.II (tests/test_cpp_host.py
// compiles tests/cpp/test_adapter_compiles.cpp against it).  Only the signatures matter (names, argument and
// return types as documented in deal.II's public API); the bodies are trivial single-rank placeholders and the
// program is never run as a solver.  With a real deal.II tree on the include path this directory is not used.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../mpi.h"

namespace dealii
{
namespace types
{
using global_dof_index = std::uint64_t;
using boundary_id      = unsigned int;
} // namespace types
namespace numbers
{
constexpr unsigned int invalid_unsigned_int = static_cast<unsigned int>(-1);
}
class Subscriptor
{
public:
  virtual ~Subscriptor() = default;
};
struct ExcMessage
{
  explicit ExcMessage(const std::string &s) : what(s) {}
  std::string what;
};
struct ExcInternalError
{
  std::string what = "internal error";
};
struct ExcNotImplemented
{
  std::string what = "not implemented";
};
#define AssertThrow(cond, exc)                     \
  do                                               \
    {                                              \
      if (!(cond))                                 \
        throw std::runtime_error((exc).what);      \
    }                                              \
  while (false)
#define Assert(cond, exc) AssertThrow(cond, exc)

namespace MemorySpace
{
struct Host
{};
struct Default
{};
} // namespace MemorySpace

template <int rank, int dim, typename Number = double>
class Tensor;
template <int dim, typename Number>
class Tensor<1, dim, Number>
{
public:
  Number       &operator[](unsigned int i) { return v[i]; }
  const Number &operator[](unsigned int i) const { return v[i]; }
  Number        norm() const
  {
    Number s = 0;
    for (int i = 0; i < dim; ++i)
      s += v[i] * v[i];
    return std::sqrt(s);
  }

private:
  Number v[dim] = {};
};
template <int dim, typename Number>
class Tensor<2, dim, Number>
{
public:
  Tensor<1, dim, Number>       &operator[](unsigned int i) { return r[i]; }
  const Tensor<1, dim, Number> &operator[](unsigned int i) const { return r[i]; }
  Number                        determinant() const { return Number(1); }

private:
  Tensor<1, dim, Number> r[dim];
};
template <int dim, typename Number = double>
using DerivativeForm1 = Tensor<2, dim, Number>;

template <int dim>
class Point : public Tensor<1, dim, double>
{};

template <typename Number>
struct VectorizedArray
{
  Number        v[1] = {};
  Number       &operator[](unsigned int) { return v[0]; }
  const Number &operator[](unsigned int) const { return v[0]; }
};

class IndexSet
{
public:
  IndexSet() = default;
  explicit IndexSet(types::global_dof_index n) : n(n) {}
  bool                    is_element(types::global_dof_index g) const { return std::binary_search(e.begin(), e.end(), g); }
  types::global_dof_index index_within_set(types::global_dof_index g) const
  {
    return std::lower_bound(e.begin(), e.end(), g) - e.begin();
  }
  std::vector<types::global_dof_index>::const_iterator begin() const { return e.begin(); }
  std::vector<types::global_dof_index>::const_iterator end() const { return e.end(); }

private:
  types::global_dof_index              n = 0;
  std::vector<types::global_dof_index> e;
};

namespace Utilities
{
template <typename T>
constexpr T pow(T base, int e)
{
  return e == 0 ? T(1) : base * pow(base, e - 1);
}
namespace MPI
{
template <typename T>
T max(const T &v, MPI_Comm)
{
  return v;
}
inline unsigned int this_mpi_process(MPI_Comm) { return 0; }
inline unsigned int n_mpi_processes(MPI_Comm) { return 1; }

class Partitioner
{
public:
  unsigned int                                        locally_owned_size() const { return 0; }
  unsigned int                                        n_ghost_indices() const { return 0; }
  unsigned int                                        global_to_local(types::global_dof_index g) const { return (unsigned int)g; }
  const std::vector<std::pair<unsigned int, unsigned int>> &import_indices() const { return ranges; }
  const std::vector<std::pair<unsigned int, unsigned int>> &import_targets() const { return ranges; }
  const std::vector<std::pair<unsigned int, unsigned int>> &ghost_targets() const { return ranges; }
  unsigned int                                        n_import_indices() const { return 0; }
  unsigned int                                        this_mpi_process() const { return 0; }
  unsigned int                                        n_mpi_processes() const { return 1; }
  MPI_Comm                                            get_mpi_communicator() const { return MPI_COMM_WORLD; }

private:
  std::vector<std::pair<unsigned int, unsigned int>> ranges;
};
} // namespace MPI
} // namespace Utilities

namespace LinearAlgebra
{
namespace distributed
{
template <typename Number, typename MemorySpaceType = MemorySpace::Host>
class Vector
{
public:
  using value_type   = Number;
  using memory_space = MemorySpaceType;
  Vector()           = default;
  void          reinit(const Vector &o, bool omit_zeroing_entries = false) { (void)omit_zeroing_entries, data.resize(o.data.size()); }
  void          reinit(const std::shared_ptr<const Utilities::MPI::Partitioner> &p) { data.resize(p->locally_owned_size() + p->n_ghost_indices()); }
  Number       *get_values() { return data.data(); }
  const Number *get_values() const { return data.data(); }
  Number       *begin() { return data.data(); }
  const Number *begin() const { return data.data(); }
  unsigned int  locally_owned_size() const { return data.size(); }
  bool          has_ghost_elements() const { return false; }
  void          zero_out_ghost_values() const {}

private:
  std::vector<Number> data;
};
} // namespace distributed
} // namespace LinearAlgebra

template <typename Number>
class AffineConstraints
{
public:
  using Entries = std::vector<std::pair<types::global_dof_index, Number>>;
  bool           is_constrained(types::global_dof_index) const { return false; }
  const Entries *get_constraint_entries(types::global_dof_index) const { return nullptr; }
  template <typename V>
  void distribute(V &) const
  {}
};

template <int dim, typename RangeNumber = double>
class Function
{
public:
  virtual ~Function() = default;
  virtual RangeNumber value(const Point<dim> &, unsigned int component = 0) const { return component * 0.0; }
};

template <int dim>
class Quadrature
{
public:
  Quadrature() = default;
  Quadrature(const std::vector<Point<dim>> &p, const std::vector<double> &w) : pts(p), wts(w) {}
  unsigned int      size() const { return wts.size(); }
  double            weight(unsigned int q) const { return wts[q]; }
  const Point<dim> &point(unsigned int q) const { return pts[q]; }

protected:
  std::vector<Point<dim>> pts;
  std::vector<double>     wts;
};
template <int dim>
class QGauss : public Quadrature<dim>
{
public:
  explicit QGauss(unsigned int n)
  {
    this->pts.resize(Utilities::pow(n, dim));
    this->wts.assign(Utilities::pow(n, dim), 1.0);
  }
};

template <int dim>
class Mapping
{
public:
  virtual ~Mapping() = default;
};
template <int dim>
class FiniteElement
{
public:
  unsigned int tensor_degree() const { return 1; }
  unsigned int n_dofs_per_cell() const { return (dim + 1) << dim; }
};

template <int dim>
class Triangulation
{
public:
  struct FaceAccessor
  {
    bool               at_boundary() const { return false; }
    types::boundary_id boundary_id() const { return 0; }
  };
  struct CellAccessor
  {
    void                            get_dof_indices(std::vector<types::global_dof_index> &) const {}
    void                            get_mg_dof_indices(std::vector<types::global_dof_index> &) const {}
    double                          minimum_vertex_distance() const { return 1; }
    double                          measure() const { return 1; }
    std::array<unsigned int, 2 * dim> face_indices() const
    {
      std::array<unsigned int, 2 * dim> a{};
      for (unsigned int i = 0; i < 2 * dim; ++i)
        a[i] = i;
      return a;
    }
    const FaceAccessor *face(unsigned int) const { return &f; }
    FaceAccessor        f;
  };
  struct cell_iterator
  {
    cell_iterator() = default;
    const CellAccessor *operator->() const { return &c; }
    CellAccessor        c;
  };
  unsigned int n_global_levels() const { return 1; }
};

template <int dim>
class DoFHandler
{
public:
  using cell_iterator = typename Triangulation<dim>::cell_iterator;
  const FiniteElement<dim> &get_fe() const { return fe; }
  const Triangulation<dim> &get_triangulation() const { return tria; }
  types::global_dof_index   n_dofs() const { return 0; }
  types::global_dof_index   n_dofs(unsigned int) const { return 0; }
  const IndexSet           &locally_owned_mg_dofs(unsigned int) const { return owned; }
  MPI_Comm                  get_communicator() const { return MPI_COMM_WORLD; }

private:
  FiniteElement<dim> fe;
  Triangulation<dim> tria;
  IndexSet           owned;
};

enum UpdateFlags
{
  update_default           = 0,
  update_values            = 1,
  update_gradients         = 2,
  update_quadrature_points = 4,
  update_JxW_values        = 8,
  update_jacobians         = 16,
  update_inverse_jacobians = 32,
  update_normal_vectors    = 64
};
inline UpdateFlags operator|(UpdateFlags a, UpdateFlags b) { return static_cast<UpdateFlags>(int(a) | int(b)); }

template <int dim>
class FEValues
{
public:
  FEValues(const Mapping<dim> &, const FiniteElement<dim> &, const Quadrature<dim> &, UpdateFlags) {}
  void                         reinit(const typename Triangulation<dim>::cell_iterator &) {}
  const DerivativeForm1<dim>  &inverse_jacobian(unsigned int) const { return J; }
  const DerivativeForm1<dim>  &jacobian(unsigned int) const { return J; }
  const Point<dim>            &quadrature_point(unsigned int) const { return x; }

private:
  DerivativeForm1<dim> J;
  Point<dim>           x;
};

namespace internal
{
namespace MatrixFreeFunctions
{
enum GeometryType : unsigned char
{
  cartesian = 0,
  affine    = 1,
  flat_faces = 2,
  general   = 3
};
template <typename Number>
struct ShapeInfo
{
  std::vector<unsigned int> lexicographic_numbering;
};
struct MappingInfoStub
{
  GeometryType get_cell_type(unsigned int) const { return cartesian; }
};
} // namespace MatrixFreeFunctions
} // namespace internal

template <int dim, typename Number, typename VectorizedArrayType = VectorizedArray<Number>>
class MatrixFree
{
public:
  struct AdditionalData
  {
    UpdateFlags  mapping_update_flags = update_default;
    unsigned int mg_level             = numbers::invalid_unsigned_int;
  };
  template <typename Q>
  void reinit(const Mapping<dim> &, const DoFHandler<dim> &, const AffineConstraints<Number> &, const Q &,
              const AdditionalData & = AdditionalData())
  {}
  const std::shared_ptr<const Utilities::MPI::Partitioner> &get_vector_partitioner(unsigned int = 0) const { return part; }
  const internal::MatrixFreeFunctions::ShapeInfo<Number>   &get_shape_info(unsigned int = 0, unsigned int = 0) const { return shape; }
  const internal::MatrixFreeFunctions::MappingInfoStub     &get_mapping_info() const { return mapping; }
  unsigned int                                              n_cell_batches() const { return 0; }
  unsigned int                               n_active_entries_per_cell_batch(unsigned int) const { return 0; }
  typename DoFHandler<dim>::cell_iterator    get_cell_iterator(unsigned int, unsigned int, unsigned int = 0) const { return {}; }
  const std::vector<unsigned int>           &get_constrained_dofs(unsigned int = 0) const { return constrained; }
  template <typename V>
  void initialize_dof_vector(V &v, unsigned int = 0) const
  {
    v.reinit(part);
  }

private:
  std::shared_ptr<const Utilities::MPI::Partitioner> part = std::make_shared<Utilities::MPI::Partitioner>();
  internal::MatrixFreeFunctions::ShapeInfo<Number>   shape;
  internal::MatrixFreeFunctions::MappingInfoStub     mapping;
  std::vector<unsigned int>                          constrained;
};

template <int dim, int fe_degree, int n_q_points_1d, int n_components, typename Number,
          typename VectorizedArrayType = VectorizedArray<Number>>
class FEEvaluation
{
public:
  explicit FEEvaluation(const MatrixFree<dim, Number, VectorizedArrayType> &, unsigned int = 0, unsigned int = 0) {}
  void                                     reinit(unsigned int) {}
  Tensor<2, dim, VectorizedArrayType>      inverse_jacobian(unsigned int) const { return {}; }
  VectorizedArrayType                      JxW(unsigned int) const { return {}; }
};

namespace MGTools
{
template <int dim>
void extract_inner_interface_dofs(const DoFHandler<dim> &, std::vector<IndexSet> &)
{}
} // namespace MGTools
} // namespace dealii
