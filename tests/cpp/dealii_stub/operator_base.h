// Stand-in for the reference's include/operator_base.h:13-73, include/config.h:3-12, the TimeIntegratorData
// interface (include/time_integration.h:10-31), SolutionHistory (:146-164) and the timer scope types
// (include/timer.h): declarations only, so that every `override` in dealii_adapter.h is checked against the
// reference's virtual signatures.  With the reference's own headers on the include path this file is not used.
#pragma once
#include <deal.II/lac/affine_constraints.h>

using namespace dealii;

using Number   = double;
using MGNumber = float;
template <typename Number>
using VectorType = dealii::LinearAlgebra::distributed::Vector<Number>;
struct SparseMatrixType
{};

class TimeIntegratorData
{
public:
  virtual ~TimeIntegratorData()                          = default;
  virtual void                       update_dt(const Number dt_new) = 0;
  virtual Number                     get_primary_weight() const     = 0;
  virtual const std::vector<Number> &get_weights() const            = 0;
  virtual unsigned int               get_order() const              = 0;
  virtual Number                     get_current_dt() const         = 0;
  virtual Number                     get_theta() const              = 0;
};

template <typename Number>
class SolutionHistory
{
public:
  VectorType<Number>                    &get_current_solution() { return solutions[0]; }
  std::vector<VectorType<Number>>       &get_vectors() { return solutions; }
  const std::vector<VectorType<Number>> &get_vectors() const { return solutions; }

private:
  std::vector<VectorType<Number>> solutions;
};

class MyTimerOutput
{
public:
  explicit MyTimerOutput(const bool = true) {}
};
class MyScope
{
public:
  MyScope(MyTimerOutput &, const std::string &, const bool = true) {}
};

template <typename Number = double>
class OperatorBase : public Subscriptor
{
public:
  using value_type = Number;
  using size_type  = types::global_dof_index;
  virtual types::global_dof_index          m() const                                                       = 0;
  virtual void                             compute_inverse_diagonal(VectorType<Number> &diagonal) const    = 0;
  virtual void                             invalidate_system()                                             = 0;
  virtual void                             set_previous_solution(const SolutionHistory<Number> &vec)       = 0;
  virtual void                             set_linearization_point(const VectorType<Number> &src)          = 0;
  virtual void                             evaluate_rhs(VectorType<Number> &dst) const                     = 0;
  virtual void                             evaluate_residual(VectorType<Number> &dst, const VectorType<Number> &src) const = 0;
  virtual void                             vmult(VectorType<Number> &dst, const VectorType<Number> &src) const = 0;
  virtual void                             vmult_interface_down(VectorType<Number> &, const VectorType<Number> &) const {}
  virtual void                             vmult_interface_up(VectorType<Number> &, const VectorType<Number> &) const {}
  virtual std::vector<std::vector<bool>>   extract_constant_modes() const { return {}; }
  virtual const AffineConstraints<Number> &get_constraints() const                                         = 0;
  virtual const SparseMatrixType          &get_system_matrix() const                                       = 0;
  virtual void                             initialize_dof_vector(VectorType<Number> &src) const            = 0;
  virtual double                           get_max_u(const VectorType<Number> &) const { return 0; }
};
