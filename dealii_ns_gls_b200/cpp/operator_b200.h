// C++ host side above the C ABI (include/glsb200.h): the reference's operator interface for the
// hot path, with the same names, argument meaning and error behaviour, so that solver_l /
// solver_nl / time_integration style callers and the parity tests read like the reference's code.
//
//   glsb::OperatorBase<Number>               <-> OperatorBase<Number>          (include/operator_base.h:13-73)
//   glsb::NavierStokesOperator<dim, Number>  <-> NavierStokesOperator<dim, N>  (include/operator_ns.h:17-189)
//   glsb::TimeIntegratorData{BDF,Theta,None} <-> include/time_integration.h:10-139
//   glsb::SolutionHistory<Number>            <-> include/time_integration.h:145-164
//   glsb::DeviceVector<Number>               <-> LinearAlgebra::distributed::Vector<Number> in device memory
//                                                (owned block followed by ghost block, config.h:9-10)
//
// deal.II is not needed to compile this header; dealii_adapter.h (same directory) is the thin
// subclass that fills a MeshDescription from MatrixFree/DoFHandler when deal.II is present.
// Errors of the C ABI are turned into exceptions, the counterpart of the reference's AssertThrow.
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/glsb200.h"

namespace glsb
{
struct Error : std::runtime_error
{
  using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char *what)
{
  if (e != cudaSuccess)
    throw Error(std::string(what) + ": " + cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------------------------
// device-resident distributed vector: [owned | ghost]
// ---------------------------------------------------------------------------------------------
template <typename Number>
class DeviceVector
{
public:
  DeviceVector() = default;
  explicit DeviceVector(std::size_t n) { reinit(n); }
  DeviceVector(const DeviceVector &o) { *this = o; }
  DeviceVector &operator=(const DeviceVector &o)
  {
    if (this != &o)
      {
        reinit(o.n_, true);
        if (n_)
          cuda_check(cudaMemcpy(d_, o.d_, n_ * sizeof(Number), cudaMemcpyDeviceToDevice), "DeviceVector copy");
      }
    return *this;
  }
  ~DeviceVector() { clear(); }

  void reinit(std::size_t n, bool omit_zeroing_entries = false)
  {
    if (n != n_)
      {
        clear();
        if (n)
          cuda_check(cudaMalloc(&d_, n * sizeof(Number)), "DeviceVector alloc");
        n_ = n;
      }
    if (!omit_zeroing_entries && n_)
      cuda_check(cudaMemset(d_, 0, n_ * sizeof(Number)), "DeviceVector zero");
  }
  void reinit(const DeviceVector &o, bool omit_zeroing_entries = false) { reinit(o.n_, omit_zeroing_entries); }
  void clear()
  {
    if (d_)
      cudaFree(d_);
    d_ = nullptr;
    n_ = 0;
  }
  std::size_t   size() const { return n_; }
  Number       *data() { return d_; }
  const Number *data() const { return d_; }
  void copy_from_host(const std::vector<Number> &h)
  {
    reinit(h.size(), true);
    cuda_check(cudaMemcpy(d_, h.data(), n_ * sizeof(Number), cudaMemcpyHostToDevice), "DeviceVector h2d");
  }
  std::vector<Number> to_host() const
  {
    std::vector<Number> h(n_);
    if (n_)
      cuda_check(cudaMemcpy(h.data(), d_, n_ * sizeof(Number), cudaMemcpyDeviceToHost), "DeviceVector d2h");
    return h;
  }
  void copy_locally_owned_data_from(const DeviceVector &o) { *this = o; }

private:
  Number     *d_ = nullptr;
  std::size_t n_ = 0;
};

// ---------------------------------------------------------------------------------------------
// time integration scalars (include/time_integration.{h,cc})
// ---------------------------------------------------------------------------------------------
class TimeIntegratorData
{
public:
  virtual ~TimeIntegratorData()                          = default;
  virtual void                       update_dt(double)   = 0;
  virtual double                     get_primary_weight() const = 0;
  virtual const std::vector<double> &get_weights() const = 0;
  virtual unsigned int               get_order() const   = 0;
  virtual double                     get_current_dt() const = 0;
  virtual double                     get_theta() const   = 0;
};

class TimeIntegratorDataBDF : public TimeIntegratorData
{
public:
  explicit TimeIntegratorDataBDF(unsigned int order) : order(order), dt(order, 0.0), weights(order + 1, 0.0) {}
  void update_dt(double dt_new) override
  {
    for (int i = (int)order - 2; i >= 0; --i)
      dt[i + 1] = dt[i];
    dt[0] = dt_new;
    std::fill(weights.begin(), weights.end(), 0.0);
    const auto eff = std::count_if(dt.begin(), dt.end(), [](double v) { return v > 0; });
    if (eff == 3)
      {
        weights[1] = -(dt[0] + dt[1]) * (dt[0] + dt[1] + dt[2]) / (dt[0] * dt[1] * (dt[1] + dt[2]));
        weights[2] = dt[0] * (dt[0] + dt[1] + dt[2]) / (dt[1] * dt[2] * (dt[0] + dt[1]));
        weights[3] = -dt[0] * (dt[0] + dt[1]) / (dt[2] * (dt[1] + dt[2]) * (dt[0] + dt[1] + dt[2]));
        weights[0] = -(weights[1] + weights[2] + weights[3]);
      }
    else if (eff == 2)
      {
        weights[0] = (2 * dt[0] + dt[1]) / (dt[0] * (dt[0] + dt[1]));
        weights[1] = -(dt[0] + dt[1]) / (dt[0] * dt[1]);
        weights[2] = dt[0] / (dt[1] * (dt[0] + dt[1]));
      }
    else if (eff == 1)
      {
        weights[0] = 1.0 / dt[0];
        weights[1] = -1.0 / dt[0];
      }
    else
      throw Error("TimeIntegratorDataBDF: not implemented");
  }
  double                     get_primary_weight() const override { return weights[0]; }
  const std::vector<double> &get_weights() const override { return weights; }
  unsigned int               get_order() const override { return order; }
  double                     get_current_dt() const override { return dt[0]; }
  double                     get_theta() const override { return 1.0; }

private:
  unsigned int        order;
  std::vector<double> dt, weights;
};

class TimeIntegratorDataTheta : public TimeIntegratorData
{
public:
  explicit TimeIntegratorDataTheta(double theta) : theta(theta), dt(0), weights(2, 0.0) {}
  void update_dt(double dt_new) override
  {
    dt         = dt_new;
    weights[0] = +1.0 / dt;
    weights[1] = -1.0 / dt;
  }
  double                     get_primary_weight() const override { return weights[0]; }
  const std::vector<double> &get_weights() const override { return weights; }
  unsigned int               get_order() const override { return 1; }
  double                     get_current_dt() const override { return dt; }
  double                     get_theta() const override { return theta; }

private:
  double              theta, dt;
  std::vector<double> weights;
};

class TimeIntegratorDataNone : public TimeIntegratorData
{
public:
  void                       update_dt(double) override {}
  double                     get_primary_weight() const override { return 0.0; }
  const std::vector<double> &get_weights() const override { return weights; }
  unsigned int               get_order() const override { return 0; }
  double                     get_current_dt() const override { return 1.0; }
  double                     get_theta() const override { return 1.0; }

private:
  std::vector<double> weights;
};

template <typename Number>
class SolutionHistory
{
public:
  explicit SolutionHistory(unsigned int size) : solutions(size) {}
  DeviceVector<Number>                    &get_current_solution() { return solutions[0]; }
  std::vector<DeviceVector<Number>>       &get_vectors() { return solutions; }
  const std::vector<DeviceVector<Number>> &get_vectors() const { return solutions; }
  void commit_solution()
  {
    for (int i = (int)solutions.size() - 2; i >= 0; --i)
      solutions[i + 1].copy_locally_owned_data_from(solutions[i]);
  }
  std::vector<DeviceVector<Number>> solutions;
};

// ---------------------------------------------------------------------------------------------
// OperatorBase (include/operator_base.h:13-73), hot-path subset
// ---------------------------------------------------------------------------------------------
template <typename Number>
class OperatorBase
{
public:
  using value_type = Number;
  using size_type  = std::uint64_t;
  using VectorType = DeviceVector<Number>;

  virtual ~OperatorBase()                                                    = default;
  virtual size_type m() const                                                = 0;
  virtual void      compute_inverse_diagonal(VectorType &diagonal) const     = 0;
  virtual void      invalidate_system()                                      = 0;
  virtual void      set_previous_solution(const SolutionHistory<Number> &h)  = 0;
  virtual void      set_linearization_point(const VectorType &src)           = 0;
  virtual void      evaluate_rhs(VectorType &dst) const                      = 0;
  virtual void      evaluate_residual(VectorType &dst, const VectorType &src) const = 0;
  virtual void      vmult(VectorType &dst, const VectorType &src) const      = 0;
  void              Tvmult(VectorType &dst, const VectorType &src) const { vmult(dst, src); } // operator_base.cc:12-18
  virtual void      initialize_dof_vector(VectorType &vec) const             = 0;
  virtual double    get_max_u(const VectorType &) const { return 1.0; }                      // operator_base.cc:52-56
  // operator_base.cc:29-50: the base class throws ExcNotImplemented
  virtual void vmult_interface_down(VectorType &, const VectorType &) const { throw Error("ExcNotImplemented"); }
  virtual void vmult_interface_up(VectorType &, const VectorType &) const { throw Error("ExcNotImplemented"); }
};

// the arrays the C ABI takes (what the deal.II adapter extracts)
struct MeshDescription
{
  int                   dim = 0, degree = 0, geometry_type = GLSB_GEOM_CARTESIAN;
  std::uint64_t         n_cells = 0, n_owned = 0, n_ghost = 0, n_global_dofs = 0;
  std::vector<uint32_t> dof_indices, row_dof, row_ptr, entry_col, constrained_indices, export_indices;
  std::vector<double>   entry_val, inv_jac, jxw, cell_h_min, cell_measure;
  // GMG-LS level operators only (operator_ns.cc:131-152)
  std::vector<uint32_t> edge_constrained_indices;
  bool                  has_edge_constrained_indices = false;
};

template <int dim, typename Number>
class NavierStokesOperator : public OperatorBase<Number>
{
public:
  using VectorType = DeviceVector<Number>;

  // argument order follows operator_ns.h:24-41; mapping/dof_handler/constraints/quadrature arrive
  // flattened in `mesh`
  NavierStokesOperator(const MeshDescription &mesh, const Number nu, const Number c_1, const Number c_2,
                       const TimeIntegratorData &time_integrator_data, const bool consider_time_derivative,
                       const bool increment_form, const bool cell_wise_stabilization, const int device = 0,
                       cudaStream_t stream = nullptr)
    : time_integrator_data(time_integrator_data), n_local(mesh.n_owned + mesh.n_ghost), n_global(mesh.n_global_dofs),
      stream(stream)
  {
    if (mesh.dim != dim)
      throw Error("NavierStokesOperator: mesh.dim != dim");
    glsb_desc d{};
    d.abi_version              = GLSB_ABI_VERSION;
    d.device                   = device;
    d.dim                      = dim;
    d.degree                   = mesh.degree;
    d.number_type              = sizeof(Number) == 8 ? GLSB_F64 : GLSB_F32;
    d.increment_form           = increment_form;
    d.consider_time_derivative = consider_time_derivative;
    d.cell_wise_stabilization  = cell_wise_stabilization;
    d.time_order               = (int)time_integrator_data.get_order();
    d.nu = nu, d.c1 = c_1, d.c2 = c_2, d.theta = time_integrator_data.get_theta();
    d.n_cells = mesh.n_cells, d.n_owned = mesh.n_owned, d.n_ghost = mesh.n_ghost;
    d.dof_indices           = mesh.dof_indices.data();
    d.n_constraint_rows     = (uint32_t)mesh.row_dof.size();
    d.row_dof               = mesh.row_dof.data();
    d.row_ptr               = mesh.row_ptr.data();
    d.entry_col             = mesh.entry_col.data();
    d.entry_val             = mesh.entry_val.data();
    d.n_constrained_indices = (uint32_t)mesh.constrained_indices.size();
    d.constrained_indices   = mesh.constrained_indices.data();
    d.geometry_type         = mesh.geometry_type;
    d.inv_jac               = mesh.inv_jac.data();
    d.jxw                   = mesh.jxw.data();
    d.cell_h_min            = mesh.cell_h_min.data();
    d.cell_measure          = mesh.cell_measure.data();
    d.n_export              = mesh.export_indices.size();
    d.export_indices        = mesh.export_indices.data();
    d.n_edge_constrained_indices   = (uint32_t)mesh.edge_constrained_indices.size();
    d.edge_constrained_indices     = mesh.edge_constrained_indices.data();
    d.has_edge_constrained_indices = mesh.has_edge_constrained_indices;
    if (glsb_create(&d, &op) != 0)
      throw Error(std::string("glsb_create: ") + glsb_last_error(nullptr));
  }
  ~NavierStokesOperator() override { glsb_destroy(op); }
  NavierStokesOperator(const NavierStokesOperator &) = delete;
  NavierStokesOperator &operator=(const NavierStokesOperator &) = delete;

  typename OperatorBase<Number>::size_type m() const override { return n_global; }

  void compute_inverse_diagonal(VectorType &diagonal) const override
  {
    initialize_dof_vector(diagonal);
    check(glsb_compute_inverse_diagonal(op, diagonal.data(), time_integrator_data.get_primary_weight(), stream));
  }
  void invalidate_system() override { check(glsb_invalidate_system(op)); }

  void set_previous_solution(const SolutionHistory<Number> &history) override
  {
    const unsigned int order = time_integrator_data.get_order();
    if (order == 0) // operator_ns.cc:242-243
      return;
    std::vector<const void *> ptr;
    for (unsigned int i = 0; i <= order; ++i)
      ptr.push_back(history.get_vectors()[i].data());
    check(glsb_set_previous_solution(op, ptr.data(), time_integrator_data.get_weights().data(), (int)order, stream));
  }
  void set_linearization_point(const VectorType &vec) override
  {
    check(glsb_set_linearization_point(op, vec.data(), time_integrator_data.get_current_dt(), stream));
  }
  // src must already carry the inhomogeneous boundary values (operator_ns.cc:655-656)
  void evaluate_residual(VectorType &dst, const VectorType &src) const override
  {
    check(glsb_evaluate_residual(op, dst.data(), src.data(), time_integrator_data.get_primary_weight(), stream));
  }
  void evaluate_rhs(VectorType &dst) const override
  {
    VectorType src(dst.size());
    evaluate_residual(dst, src);
  }
  void vmult(VectorType &dst, const VectorType &src) const override
  {
    check(glsb_vmult(op, dst.data(), src.data(), time_integrator_data.get_primary_weight(), stream));
  }
  // vmult for callers whose vectors live in host memory (the reference's VectorType, config.h:9-10)
  void vmult_host(Number *dst_host, const Number *src_host) const
  {
    check(glsb_vmult_host(op, dst_host, src_host, time_integrator_data.get_primary_weight(), stream));
  }
  void vmult_interface_down(VectorType &dst, const VectorType &src) const override
  {
    check(glsb_vmult_interface_down(op, dst.data(), src.data(), time_integrator_data.get_primary_weight(), stream));
  }
  void vmult_interface_up(VectorType &dst, const VectorType &src) const override
  {
    check(glsb_vmult_interface_up(op, dst.data(), src.data(), time_integrator_data.get_primary_weight(), stream));
  }
  void initialize_dof_vector(VectorType &vec) const override { vec.reinit(n_local); }
  double get_max_u(const VectorType &src) const override
  {
    double v = 0;
    check(glsb_get_max_u(op, src.data(), &v, stream));
    return v;
  }
  const char *vmult_variant() const { return glsb_vmult_variant(op); }
  void        synchronize() const { cuda_check(cudaStreamSynchronize(stream), "synchronize"); }

private:
  void check(int rc) const
  {
    if (rc != 0)
      throw Error(glsb_last_error(op));
  }
  const TimeIntegratorData &time_integrator_data; // read at call time, like the reference (operator_ns.cc:958)
  glsb_op                  *op = nullptr;
  std::uint64_t             n_local, n_global;
  cudaStream_t              stream;
};

// ---------------------------------------------------------------------------------------------
// the device pieces of the callers, C++ side
// ---------------------------------------------------------------------------------------------
template <typename Number>
constexpr int number_type_of()
{
  return std::is_same<Number, double>::value ? GLSB_F64 : GLSB_F32;
}

// deal.II's MGTwoLevelTransfer<dim, VectorType> between a level and the next coarser one (main.cc:540-563):
// reinit() takes what the two DoFHandlers give -- the coarse cells' dof indices (GLSB_CONSTRAINED_BIT | row for
// constrained ones), their children's fine dof indices, the coarse constraint rows and deal.II's weights.
template <int dim, typename Number>
class MGTwoLevelTransfer
{
public:
  MGTwoLevelTransfer() = default;
  MGTwoLevelTransfer(const MGTwoLevelTransfer &) = delete;
  ~MGTwoLevelTransfer() { glsb_transfer_destroy(t_); }

  void reinit(int degree, std::uint64_t n_fine_dofs, std::uint64_t n_coarse_dofs,
              const std::vector<std::uint32_t> &coarse_dof_indices, const std::vector<std::uint32_t> &fine_dof_indices,
              const std::vector<double> &weights, const std::vector<std::uint32_t> &row_ptr = {},
              const std::vector<std::uint32_t> &entry_col = {}, const std::vector<double> &entry_val = {})
  {
    glsb_transfer_destroy(t_);
    t_ = nullptr;
    const std::uint64_t ndof = (dim + 1) * (dim == 2 ? (degree + 1) * (degree + 1) : (degree + 1) * (degree + 1) * (degree + 1));
    glsb_transfer_desc  d{};
    d.abi_version = GLSB_ABI_VERSION;
    cuda_check(cudaGetDevice(&d.device), "cudaGetDevice");
    d.dim = dim, d.degree = degree, d.number_type = number_type_of<Number>();
    d.n_coarse_cells = coarse_dof_indices.size() / ndof;
    d.n_fine_dofs = n_fine_dofs, d.n_coarse_dofs = n_coarse_dofs;
    d.coarse_dof_indices = coarse_dof_indices.data(), d.fine_dof_indices = fine_dof_indices.data();
    d.n_constraint_rows = row_ptr.empty() ? 0 : (std::uint32_t)row_ptr.size() - 1;
    d.row_ptr = row_ptr.data(), d.entry_col = entry_col.data(), d.entry_val = entry_val.data();
    d.weights = weights.empty() ? nullptr : weights.data();
    if (glsb_transfer_create(&d, &t_) != 0)
      throw Error(std::string("glsb_transfer_create: ") + glsb_transfer_last_error(nullptr));
  }
  void prolongate_and_add(DeviceVector<Number> &dst_fine, const DeviceVector<Number> &src_coarse) const
  {
    check(glsb_transfer_prolongate_and_add(t_, dst_fine.data(), src_coarse.data(), nullptr));
  }
  void restrict_and_add(DeviceVector<Number> &dst_coarse, const DeviceVector<Number> &src_fine) const
  {
    check(glsb_transfer_restrict_and_add(t_, dst_coarse.data(), src_fine.data(), nullptr));
  }
  void interpolate(DeviceVector<Number> &dst_coarse, const DeviceVector<Number> &src_fine) const
  {
    check(glsb_transfer_interpolate(t_, dst_coarse.data(), src_fine.data(), nullptr));
  }

private:
  void check(int rc) const
  {
    if (rc != 0)
      throw Error(std::string("glsb_transfer: ") + glsb_transfer_last_error(t_));
  }
  glsb_transfer *t_ = nullptr;
};

// the vector operations deal.II's SolverGMRES / Multigrid do on LinearAlgebra::distributed::Vector
// (solver_l.cc:46-74), on device vectors
template <typename Number>
struct DeviceVectorOps
{
  static void chk(int rc, const char *what)
  {
    if (rc != 0)
      throw Error(std::string(what) + " failed");
  }
  // y = a x + b y
  static void axpby(DeviceVector<Number> &y, double a, const DeviceVector<Number> &x, double b)
  {
    chk(glsb_vec_axpby(y.data(), a, x.data(), b, y.size(), number_type_of<Number>(), nullptr), "glsb_vec_axpby");
  }
  // out[j] = V_j . w, j < k, V = k vectors of length w.size() stored back to back
  static std::vector<double> multi_dot(const DeviceVector<Number> &V, int k, const DeviceVector<Number> &w)
  {
    DeviceVector<double> out(k);
    chk(glsb_vec_multi_dot(out.data(), V.data(), w.size(), k, w.data(), w.size(), number_type_of<Number>(), nullptr),
        "glsb_vec_multi_dot");
    return out.to_host();
  }
  // w += scale * sum_j coef[j] V_j
  static void multi_axpy(DeviceVector<Number> &w, const DeviceVector<Number> &V, const std::vector<double> &coef,
                         double scale)
  {
    DeviceVector<double> c;
    c.copy_from_host(coef);
    chk(glsb_vec_multi_axpy(w.data(), V.data(), w.size(), (int)coef.size(), c.data(), scale, w.size(),
                            number_type_of<Number>(), nullptr),
        "glsb_vec_multi_axpy");
  }
  template <typename Other>
  static void convert(DeviceVector<Number> &dst, const DeviceVector<Other> &src)
  {
    chk(glsb_vec_convert(dst.data(), number_type_of<Number>(), src.data(), number_type_of<Other>(), dst.size(), nullptr),
        "glsb_vec_convert");
  }
};

} // namespace glsb
