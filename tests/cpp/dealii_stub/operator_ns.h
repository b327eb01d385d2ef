// Stand-in for the reference's include/operator_ns.h:24-92 (the CPU NavierStokesOperator the adapter retains for
// get_system_matrix / extract_constant_modes / get_constraints): constructor signature and the members used.
#pragma once
#include "operator_base.h"
#include <deal.II/matrix_free/matrix_free.h>

template <int dim, typename Number>
class NavierStokesOperator : public OperatorBase<Number>
{
public:
  NavierStokesOperator(const Mapping<dim> &, const DoFHandler<dim> &, const AffineConstraints<Number> &,
                       const AffineConstraints<Number> &constraints, const AffineConstraints<Number> &, const Quadrature<dim> &,
                       const Number, const Number, const Number, const std::set<unsigned int> &,
                       const std::map<unsigned int, std::shared_ptr<Function<dim, double>>> &, const TimeIntegratorData &,
                       const bool, const bool, const bool, const unsigned int = numbers::invalid_unsigned_int)
    : constraints(constraints)
  {}
  types::global_dof_index          m() const override { return 0; }
  void                             compute_inverse_diagonal(VectorType<Number> &) const override {}
  void                             invalidate_system() override {}
  void                             set_previous_solution(const SolutionHistory<Number> &) override {}
  void                             set_linearization_point(const VectorType<Number> &) override {}
  void                             evaluate_rhs(VectorType<Number> &) const override {}
  void                             evaluate_residual(VectorType<Number> &, const VectorType<Number> &) const override {}
  void                             vmult(VectorType<Number> &, const VectorType<Number> &) const override {}
  const AffineConstraints<Number> &get_constraints() const override { return constraints; }
  const SparseMatrixType          &get_system_matrix() const override { return matrix; }
  void                             initialize_dof_vector(VectorType<Number> &) const override {}

private:
  const AffineConstraints<Number> &constraints;
  SparseMatrixType                 matrix;
};
