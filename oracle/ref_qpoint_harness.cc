// Harness around the reference's OWN quadrature-point kernel: NavierStokesOperator<dim, Number>::do_vmult_cell
// (both branches) and symm_scalar_product_add, include/operator_ns.cc:880-1182.  TEST INFRASTRUCTURE.
//
// operator_ns.cc as a whole needs deal.II (MatrixFree, FEEvaluation, Trilinos, ...) and cannot be built here.  The
// recipe in oracle/Makefile (target _ref) cuts exactly those lines out of /root/reference/include/operator_ns.cc at
// build time into a scratch file (_ref/qpoint_extract.inc, deleted after the compile, never committed) and
// compiles them, unmodified, inside this file: the class below declares the members the two function bodies read,
// with the names and types of include/operator_ns.h:97-140, on top of the stand-in types of
// ref_shim/qpoint_shim.h.  What runs is the reference's own arithmetic between get_value / get_gradient and
// submit_value / submit_gradient; deal.II's evaluate / integrate, geometry and vector access are not part of it.
// tests/test_reference_qpoint.py compares oracle/gls_oracle.py's _cell_newton / _cell_fixed_point with it.
#include "qpoint_shim.h"

#include <utility>

using namespace dealii;

struct TimeIntegratorData
{
  double weight = 0.0;
  double
  get_primary_weight() const
  {
    return weight;
  }
};

template <int dim, typename Number>
class NavierStokesOperator
{
public:
  using FECellIntegrator = FEEvaluation<dim, -1, 0, dim + 1, Number>;
  using FEFaceIntegrator = FEFaceEvaluation<dim, -1, 0, dim + 1, Number>;

  NavierStokesOperator(const TimeIntegratorData &t, const Number theta, const Number nu, const bool ctd,
                       const bool increment_form, const bool cell_wise)
    : theta(theta)
    , nu(nu)
    , time_integrator_data(t)
    , consider_time_derivative(ctd)
    , increment_form(increment_form)
    , cell_wise_stabilization(cell_wise)
  {}

  // include/operator_ns.h:103-132
  const VectorizedArray<Number> theta;
  const VectorizedArray<Number> nu;
  const TimeIntegratorData     &time_integrator_data;
  const bool                    consider_time_derivative;
  const bool                    increment_form;
  const bool                    cell_wise_stabilization;

  AlignedVector<VectorizedArray<Number>> delta_1;
  AlignedVector<VectorizedArray<Number>> delta_2;

  Table<2, VectorizedArray<Number>> delta_1_q;
  Table<2, VectorizedArray<Number>> delta_2_q;

  Table<2, Tensor<1, dim, VectorizedArray<Number>>> u_star_value;
  Table<2, Tensor<2, dim, VectorizedArray<Number>>> u_star_gradient;
  Table<2, Tensor<1, dim, VectorizedArray<Number>>> p_star_gradient;

  Table<2, Tensor<1, dim, VectorizedArray<Number>>> u_time_derivative_old;
  Table<2, Tensor<2, dim, VectorizedArray<Number>>> u_old_gradient;
  Table<2, Tensor<1, dim, VectorizedArray<Number>>> p_old_gradient;

  // boundary faces with outflow terms (include/operator_ns.h:109-111, :134-136)
  struct
  {
    unsigned int
    n_inner_face_batches() const
    {
      return 0;
    }
  } matrix_free;
  std::set<unsigned int>                                all_outflow_bcs_cut;
  std::map<unsigned int, int>                           all_outflow_bcs_nitsche; // the reference maps to a Function
  Table<1, VectorizedArray<Number>>                     effective_beta_face;
  Table<2, Tensor<1, dim + 1, VectorizedArray<Number>>> face_target_velocity;
  Table<2, Tensor<1, dim, VectorizedArray<Number>>>     face_velocity;

  template <bool evaluate_residual>
  void
  do_vmult_cell(FECellIntegrator &integrator) const;

  template <bool evaluate_residual>
  void
  do_vmult_boundary(FEFaceIntegrator &integrator) const;
};

// ---- the reference's code, lines 880-1182 of include/operator_ns.cc, as extracted by the Makefile ----
#include "qpoint_extract.inc"

// ---- do_vmult_boundary ("cut" and Nitsche outflow faces), include/operator_ns.cc:1195-1301 ----
#include "boundary_extract.inc"

// ---- compute_penalty_parameters: the body of the cell loop, include/operator_ns.cc:348-421 ----
// (tau / stau, the cell-wise delta of :364-388 and the q-point-wise delta after Lethe of :390-420).  The lines
// before it set up deal.II objects; here they are replaced by locals of the same names on the stand-in types.
struct CellIterator
{
  double h_min, vol;
  const CellIterator *
  operator->() const
  {
    return this;
  }
  double
  minimum_vertex_distance() const
  {
    return h_min;
  }
  double
  measure() const
  {
    return vol;
  }
};

struct PenaltyMatrixFree
{
  std::vector<CellIterator> cells;
  unsigned int              degree = 0;
  unsigned int
  n_active_entries_per_cell_batch(const unsigned int) const
  {
    return 1; // one lane per "batch"
  }
  CellIterator
  get_cell_iterator(const unsigned int cell, const unsigned int) const
  {
    return cells[cell];
  }
  // boundary faces: face f belongs to cell f of `cells`
  unsigned int
  n_active_entries_per_face_batch(const unsigned int) const
  {
    return 1;
  }
  std::pair<CellIterator, unsigned int>
  get_face_iterator(const unsigned int face, const unsigned int) const
  {
    return {cells[face], 0u};
  }
  // matrix_free.get_dof_handler().get_fe().tensor_degree()
  const PenaltyMatrixFree &
  get_dof_handler() const
  {
    return *this;
  }
  const PenaltyMatrixFree &
  get_fe() const
  {
    return *this;
  }
  unsigned int
  tensor_degree() const
  {
    return degree;
  }
};

struct PenaltyTime
{
  double dt;
  double
  get_current_dt() const
  {
    return dt;
  }
};

template <int dim, typename Number>
struct PenaltyIntegrator
{
  const double *u; // [cell][q][dim]
  unsigned int  n_q, cell = 0;
  void
  reinit(const unsigned int c)
  {
    cell = c;
  }
  void
  read_dof_values_plain(const int &)
  {}
  void
  evaluate(const EvaluationFlags::EvaluationFlags)
  {}
  std::vector<unsigned int>
  quadrature_point_indices() const
  {
    std::vector<unsigned int> r(n_q);
    for (unsigned int i = 0; i < n_q; ++i)
      r[i] = i;
    return r;
  }
  Tensor<1, dim, VectorizedArray<Number>>
  get_value(const unsigned int q) const
  {
    Tensor<1, dim, VectorizedArray<Number>> t;
    for (int i = 0; i < dim; ++i)
      t[i] = u[((std::size_t)cell * n_q + q) * dim + i];
    return t;
  }
};

template <int dim, typename Number>
struct PenaltyHarness
{
  PenaltyTime                            time_integrator_data;
  PenaltyMatrixFree                      matrix_free;
  VectorizedArray<Number>                nu;
  Number                                 c_1, c_2;
  AlignedVector<VectorizedArray<Number>> delta_1, delta_2;
  Table<2, VectorizedArray<Number>>      delta_1_q, delta_2_q;

  // effective_beta_face of the outflow faces, include/operator_ns.cc:428-457 (one "face" per entry of
  // matrix_free.cells)
  Table<1, VectorizedArray<Number>> effective_beta_face;
  void
  compute_beta()
  {
    const unsigned int n_inner_faces = 0, n_boundary_faces = matrix_free.cells.size();
#include "beta_extract.inc"
  }

  void
  compute(const double *u, const unsigned int n_cells, const unsigned int n_quadrature_points,
          const unsigned int fe_degree)
  {
    PenaltyIntegrator<dim, Number> integrator{u, n_quadrature_points};
    const int                      vec = 0;
    (void)vec;
#include "penalty_extract.inc"
  }
};

template <int dim>
static void
run_penalty(const double dt, const double nu, const double c1, const double c2, const int degree, const int n_cells,
            const int n_q, const double *u, const double *h_min, const double *measure, double *d1_cell,
            double *d2_cell, double *d1_q, double *d2_q)
{
  PenaltyHarness<dim, double> h;
  h.time_integrator_data.dt = dt;
  h.nu                      = nu;
  h.c_1                     = c1;
  h.c_2                     = c2;
  h.matrix_free.cells.resize(n_cells);
  for (int c = 0; c < n_cells; ++c)
    h.matrix_free.cells[c] = CellIterator{h_min[c], measure[c]};
  h.compute(u, n_cells, n_q, degree);
  for (int c = 0; c < n_cells; ++c)
    {
      d1_cell[c] = h.delta_1[c].data;
      d2_cell[c] = h.delta_2[c].data;
      for (int q = 0; q < n_q; ++q)
        {
          d1_q[c * n_q + q] = h.delta_1_q[c][q].data;
          d2_q[c * n_q + q] = h.delta_2_q[c][q].data;
        }
    }
}

// velocity values u[cell][q][dim] at the quadrature points, h_min / measure per cell -> the four delta tables
extern "C" int
refq_penalty(int dim, double dt, double nu, double c1, double c2, int degree, int n_cells, int n_q, const double *u,
             const double *h_min, const double *measure, double *d1_cell, double *d2_cell, double *d1_q,
             double *d2_q)
{
  if (dim == 2)
    run_penalty<2>(dt, nu, c1, c2, degree, n_cells, n_q, u, h_min, measure, d1_cell, d2_cell, d1_q, d2_q);
  else if (dim == 3)
    run_penalty<3>(dt, nu, c1, c2, degree, n_cells, n_q, u, h_min, measure, d1_cell, d2_cell, d1_q, d2_q);
  else
    return 1;
  return 0;
}

template <int dim, typename Number = double>
static void
run(const bool residual, const bool increment_form, const bool ctd, const bool cell_wise, const double theta,
    const double nu, const double weight, const int n_q, const double *value, const double *grad,
    const double *u_star, const double *u_star_grad, const double *p_star_grad, const double *u_tdo,
    const double *u_old_grad, const double *p_old_grad, const double *d1, const double *d2, double *value_out,
    double *grad_out)
{
  constexpr int      C = dim + 1;
  TimeIntegratorData ti;
  ti.weight = weight;
  NavierStokesOperator<dim, Number> op(ti, Number(theta), Number(nu), ctd, increment_form, cell_wise);
  op.delta_1.assign(1, VectorizedArray<Number>(Number(d1[0])));
  op.delta_2.assign(1, VectorizedArray<Number>(Number(d2[0])));
  op.delta_1_q.reinit(1, n_q);
  op.delta_2_q.reinit(1, n_q);
  op.u_star_value.reinit(1, n_q);
  op.u_star_gradient.reinit(1, n_q);
  op.p_star_gradient.reinit(1, n_q);
  if (u_tdo)
    op.u_time_derivative_old.reinit(1, n_q);
  if (u_old_grad)
    {
      op.u_old_gradient.reinit(1, n_q);
      op.p_old_gradient.reinit(1, n_q);
    }
  typename NavierStokesOperator<dim, Number>::FECellIntegrator phi;
  phi.values_in.resize(n_q);
  phi.values_out.resize(n_q);
  phi.gradients_in.resize(n_q);
  phi.gradients_out.resize(n_q);
  for (int q = 0; q < n_q; ++q)
    {
      if (!cell_wise)
        {
          op.delta_1_q[0][q] = Number(d1[q]);
          op.delta_2_q[0][q] = Number(d2[q]);
        }
      for (int i = 0; i < dim; ++i)
        {
          op.u_star_value[0][q][i]    = Number(u_star[q * dim + i]);
          op.p_star_gradient[0][q][i] = Number(p_star_grad[q * dim + i]);
          if (u_tdo)
            op.u_time_derivative_old[0][q][i] = Number(u_tdo[q * dim + i]);
          if (u_old_grad)
            op.p_old_gradient[0][q][i] = Number(p_old_grad[q * dim + i]);
          for (int j = 0; j < dim; ++j)
            {
              op.u_star_gradient[0][q][i][j] = Number(u_star_grad[(q * dim + i) * dim + j]);
              if (u_old_grad)
                op.u_old_gradient[0][q][i][j] = Number(u_old_grad[(q * dim + i) * dim + j]);
            }
        }
      for (int c = 0; c < C; ++c)
        {
          phi.values_in[q][c] = Number(value[q * C + c]);
          for (int j = 0; j < dim; ++j)
            phi.gradients_in[q][c][j] = Number(grad[(q * C + c) * dim + j]);
        }
    }
  if (residual)
    op.template do_vmult_cell<true>(phi);
  else
    op.template do_vmult_cell<false>(phi);
  for (int q = 0; q < n_q; ++q)
    for (int c = 0; c < C; ++c)
      {
        value_out[q * C + c] = phi.values_out[q][c].data;
        for (int j = 0; j < dim; ++j)
          grad_out[(q * C + c) * dim + j] = phi.gradients_out[q][c][j].data;
      }
}

extern "C" int
refq_apply(int dim, int residual, int increment_form, int ctd, int cell_wise, double theta, double nu, double weight,
           int n_q, const double *value, const double *grad, const double *u_star, const double *u_star_grad,
           const double *p_star_grad, const double *u_tdo, const double *u_old_grad, const double *p_old_grad,
           const double *d1, const double *d2, double *value_out, double *grad_out)
{
  if (dim == 2)
    run<2>(residual, increment_form, ctd, cell_wise, theta, nu, weight, n_q, value, grad, u_star, u_star_grad,
           p_star_grad, u_tdo, u_old_grad, p_old_grad, d1, d2, value_out, grad_out);
  else if (dim == 3)
    run<3>(residual, increment_form, ctd, cell_wise, theta, nu, weight, n_q, value, grad, u_star, u_star_grad,
           p_star_grad, u_tdo, u_old_grad, p_old_grad, d1, d2, value_out, grad_out);
  else
    return 1;
  return 0;
}

// the same with Number = float (the multigrid level operators, include/config.h:7): inputs are rounded to float,
// the time-integration weight stays a double as in the reference (TimeIntegratorData returns the global Number)
extern "C" int
refq_apply_f32(int dim, int residual, int increment_form, int ctd, int cell_wise, double theta, double nu, double weight,
               int n_q, const double *value, const double *grad, const double *u_star, const double *u_star_grad,
               const double *p_star_grad, const double *u_tdo, const double *u_old_grad, const double *p_old_grad,
               const double *d1, const double *d2, double *value_out, double *grad_out)
{
  if (dim == 2)
    run<2, float>(residual, increment_form, ctd, cell_wise, theta, nu, weight, n_q, value, grad, u_star, u_star_grad,
                  p_star_grad, u_tdo, u_old_grad, p_old_grad, d1, d2, value_out, grad_out);
  else if (dim == 3)
    run<3, float>(residual, increment_form, ctd, cell_wise, theta, nu, weight, n_q, value, grad, u_star, u_star_grad,
                  p_star_grad, u_tdo, u_old_grad, p_old_grad, d1, d2, value_out, grad_out);
  else
    return 1;
  return 0;
}

template <int dim>
static void
run_boundary(const bool residual, const int kind, const double nu, const double beta, const int n_q,
             const double *value, const double *grad, const double *normal, const double *face_velocity,
             const double *target, double *value_out, double *grad_out, double *dof_values, const int n_dof_values)
{
  constexpr int      C = dim + 1;
  TimeIntegratorData ti;
  NavierStokesOperator<dim, double> op(ti, 1.0, nu, true, true, true);
  if (kind == 1)
    op.all_outflow_bcs_cut.insert(7);
  if (kind == 2)
    op.all_outflow_bcs_nitsche[7] = 0;
  op.effective_beta_face.reinit(1);
  op.effective_beta_face[0] = beta;
  op.face_velocity.reinit(1, n_q);
  op.face_target_velocity.reinit(1, n_q);
  typename NavierStokesOperator<dim, double>::FEFaceIntegrator phi;
  phi.id = 7;
  phi.values_in.resize(n_q);
  phi.values_out.resize(n_q);
  phi.gradients_in.resize(n_q);
  phi.gradients_out.resize(n_q);
  phi.normals.resize(n_q);
  phi.dofs_per_cell = n_dof_values;
  phi.dof_values.resize(n_dof_values);
  for (int i = 0; i < n_dof_values; ++i)
    phi.dof_values[i] = dof_values[i];
  for (int q = 0; q < n_q; ++q)
    {
      for (int i = 0; i < dim; ++i)
        {
          phi.normals[q][i]         = normal[q * dim + i];
          op.face_velocity[0][q][i] = face_velocity[q * dim + i];
        }
      for (int c = 0; c < C; ++c)
        {
          op.face_target_velocity[0][q][c] = target[q * C + c];
          phi.values_in[q][c]              = value[q * C + c];
          for (int j = 0; j < dim; ++j)
            phi.gradients_in[q][c][j] = grad[(q * C + c) * dim + j];
        }
    }
  if (residual)
    op.template do_vmult_boundary<true>(phi);
  else
    op.template do_vmult_boundary<false>(phi);
  for (int q = 0; q < n_q; ++q)
    for (int c = 0; c < C; ++c)
      {
        value_out[q * C + c] = phi.values_out[q][c].data;
        for (int j = 0; j < dim; ++j)
          grad_out[(q * C + c) * dim + j] = phi.gradients_out[q][c][j].data;
      }
  for (int i = 0; i < n_dof_values; ++i)
    dof_values[i] = phi.dof_values[i].data;
}

// one boundary face: kind 1 = its boundary id is in all_outflow_bcs_cut, 2 = in all_outflow_bcs_nitsche, 0 = in
// neither (the reference then zeroes the face's dof values and returns); target[q][dim + 1]
extern "C" int
refq_boundary(int dim, int residual, int kind, double nu, double beta, int n_q, const double *value,
              const double *grad, const double *normal, const double *face_velocity, const double *target,
              double *value_out, double *grad_out, double *dof_values, int n_dof_values)
{
  if (dim == 2)
    run_boundary<2>(residual, kind, nu, beta, n_q, value, grad, normal, face_velocity, target, value_out, grad_out,
                    dof_values, n_dof_values);
  else if (dim == 3)
    run_boundary<3>(residual, kind, nu, beta, n_q, value, grad, normal, face_velocity, target, value_out, grad_out,
                    dof_values, n_dof_values);
  else
    return 1;
  return 0;
}

// beta of the outflow faces (operator_ns.cc:428-457) from the measures of the cells behind them
extern "C" int
refq_face_beta(int dim, int degree, int n_faces, const double *measure, double *beta)
{
  auto run = [&](auto &h) {
    h.matrix_free.degree = degree;
    h.matrix_free.cells.resize(n_faces);
    for (int f = 0; f < n_faces; ++f)
      h.matrix_free.cells[f] = CellIterator{0.0, measure[f]};
    h.compute_beta();
    for (int f = 0; f < n_faces; ++f)
      beta[f] = h.effective_beta_face[f].data;
  };
  if (dim == 2)
    {
      PenaltyHarness<2, double> h;
      run(h);
    }
  else if (dim == 3)
    {
      PenaltyHarness<3, double> h;
      run(h);
    }
  else
    return 1;
  return 0;
}
