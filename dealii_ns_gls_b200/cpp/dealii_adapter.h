// deal.II-side binding of libglsb200.so: NavierStokesOperatorB200<dim, Number>, a drop-in subclass of
// the reference's OperatorBase<Number> (include/operator_base.h:13-73) with the constructor
// signature of NavierStokesOperator<dim, Number> (include/operator_ns.h:24-41).  solver_l, solver_nl,
// multigrid, time_integration and main.cc stay unchanged; main.cc:333-348 / :513-529 / :683-702 and
// performance.cc:48-62 only swap the class name.
//
// deal.II (>= 9.6, with p4est and Trilinos), MPI and the reference's own headers are not available to this
// build, so the class cannot be RUN here.  It is compiled, instantiated for
// <2|3, double|float> and linked against libglsb200.so + NCCL by tests/cpp/test_adapter_compiles.cpp on top of
// tests/cpp/dealii_stub/ (the used deal.II / reference signatures only), which checks syntax, every
// `override` against OperatorBase<Number>, and the C-ABI calls.  The file is header-only and written
// against the public deal.II API; every accessor used is listed in INTEGRATION.md so a maintainer with
// a deal.II tree can check it in one pass.  What it does is mechanical: flatten
// MatrixFree / DoFHandler / AffineConstraints / Partitioner into the arrays of glsb_desc
//  once, then forward each virtual to the matching C-ABI call.
#pragma once

#if __has_include(<deal.II/matrix_free/matrix_free.h>)

#include <deal.II/base/mpi.h>
#include <deal.II/base/partitioner.h>
#include <deal.II/lac/affine_constraints.h>
#include <deal.II/lac/la_parallel_vector.h>
#include <deal.II/matrix_free/fe_evaluation.h>
#include <deal.II/matrix_free/matrix_free.h>
#include <deal.II/multigrid/mg_tools.h>

#include <cuda_runtime.h>
#include <nccl.h>

#include "../../include/glsb200.h"
#include "operator_base.h" // the reference's header
#include "operator_ns.h"   // retained CPU operator for get_system_matrix() on the coarse level

namespace glsb
{
using namespace dealii;

/**
 * Number = double for the Krylov operator, float for the multigrid level operators
 * (include/config.h:6-7).  VectorType<Number> is the reference's alias (config.h:9-10).  With a
 * Kokkos-CUDA deal.II and `MemorySpace::Default` vectors the device pointers are used directly;
 * with host vectors (the reference's default alias) a pinned mirror is copied per call -- the
 * "plumbing" mode that pays 16 B/DoF over PCIe and is only meant for bring-up.
 */
template <int dim, typename Number>
class NavierStokesOperatorB200 : public OperatorBase<Number>
{
public:
  using FECellIntegrator = FEEvaluation<dim, -1, 0, dim + 1, Number>;

  NavierStokesOperatorB200(const Mapping<dim>              &mapping,
                           const DoFHandler<dim>           &dof_handler,
                           const AffineConstraints<Number> &constraints_homogeneous,
                           const AffineConstraints<Number> &constraints,
                           const AffineConstraints<Number> &constraints_inhomogeneous,
                           const Quadrature<dim>           &quadrature,
                           const Number                     nu,
                           const Number                     c_1,
                           const Number                     c_2,
                           const std::set<unsigned int>    &all_outflow_bcs_cut,
                           const std::map<unsigned int, std::shared_ptr<Function<dim, double>>> &all_outflow_bcs_nitsche,
                           const TimeIntegratorData &time_integrator_data,
                           const bool                consider_time_derivative,
                           const bool                increment_form,
                           const bool                cell_wise_stabilization,
                           const unsigned int        mg_level = numbers::invalid_unsigned_int)
    : constraints_inhomogeneous(constraints_inhomogeneous)
    , time_integrator_data(time_integrator_data)
    , cpu_operator(mapping, dof_handler, constraints_homogeneous, constraints, constraints_inhomogeneous, quadrature,
                   nu, c_1, c_2, all_outflow_bcs_cut, all_outflow_bcs_nitsche, time_integrator_data,
                   consider_time_derivative, increment_form, cell_wise_stabilization, mg_level)
  {
    // boundary-face outflow terms (operator_ns.cc:1195-1301): the faces are collected below (outflow_faces)

    typename MatrixFree<dim, Number>::AdditionalData ad;
    ad.mapping_update_flags = update_values | update_gradients; // operator_ns.cc:112
    ad.mg_level             = mg_level;
    matrix_free.reinit(mapping, dof_handler, constraints_homogeneous, quadrature, ad);

    const auto &part   = *matrix_free.get_vector_partitioner();
    const auto &shape  = matrix_free.get_shape_info();
    const auto &lex    = shape.lexicographic_numbering; // component-blocked lexicographic -> cell dof
    const unsigned int degree = dof_handler.get_fe().tensor_degree();
    const unsigned int n_q    = quadrature.size();
    const unsigned int ndof   = dof_handler.get_fe().n_dofs_per_cell();

    // ---- constraint rows (homogeneous part only, operator_ns.cc:107-108) -------------------
    std::vector<uint32_t> row_dof, row_ptr(1, 0), entry_col;
    std::vector<double>   entry_val;
    std::map<types::global_dof_index, uint32_t> row_of;
    auto row_for = [&](const types::global_dof_index g) -> uint32_t {
      auto it = row_of.find(g);
      if (it != row_of.end())
        return it->second;
      const uint32_t r = row_dof.size();
      row_of[g]        = r;
      row_dof.push_back(part.global_to_local(g));
      if (const auto *entries = constraints_homogeneous.get_constraint_entries(g))
        for (const auto &e : *entries)
          {
            entry_col.push_back(part.global_to_local(e.first));
            entry_val.push_back(e.second);
          }
      row_ptr.push_back(entry_col.size());
      return r;
    };

    // ---- cells: dof indices, geometry, h, measure ---------------------------------------------
    std::vector<uint32_t> dof_indices;
    std::vector<double>   inv_jac, jxw, h_min, measure;
    bool                  all_cartesian = true;
    for (unsigned int b = 0; b < matrix_free.n_cell_batches(); ++b)
      if (matrix_free.get_mapping_info().get_cell_type(b) != internal::MatrixFreeFunctions::cartesian)
        all_cartesian = false;
    std::vector<types::global_dof_index> cell_dofs(ndof);
    FECellIntegrator                     phi(matrix_free);
    std::uint64_t                        n_cells = 0;
    for (unsigned int b = 0; b < matrix_free.n_cell_batches(); ++b)
      {
        phi.reinit(b);
        for (unsigned int v = 0; v < matrix_free.n_active_entries_per_cell_batch(b); ++v, ++n_cells)
          {
            const auto cell = matrix_free.get_cell_iterator(b, v);
            if (mg_level == numbers::invalid_unsigned_int)
              cell->get_dof_indices(cell_dofs);
            else
              cell->get_mg_dof_indices(cell_dofs);
            for (unsigned int i = 0; i < ndof; ++i)
              {
                const auto g = cell_dofs[lex[i]];
                dof_indices.push_back(constraints_homogeneous.is_constrained(g) ?
                                        (GLSB_CONSTRAINED_BIT | row_for(g)) :
                                        part.global_to_local(g));
              }
            h_min.push_back(cell->minimum_vertex_distance()); // operator_ns.cc:374
            measure.push_back(cell->measure());               // operator_ns.cc:399
            if (all_cartesian)
              {
                const auto J = phi.inverse_jacobian(0); // J^{-T}, diagonal for Cartesian cells
                for (unsigned int d = 0; d < dim; ++d)
                  inv_jac.push_back(J[d][d][v]);
                // JxW(q) = det J * w_q  =>  det J = JxW(0) / w_0
                jxw.push_back(phi.JxW(0)[v] / quadrature.weight(0));
              }
            else
              for (unsigned int q = 0; q < n_q; ++q)
                {
                  const auto J = phi.inverse_jacobian(q); // J[j][e] = (J^{-1})_{e j}
                  for (unsigned int e = 0; e < dim; ++e)
                    for (unsigned int j = 0; j < dim; ++j)
                      inv_jac.push_back(J[j][e][v]);
                  jxw.push_back(phi.JxW(q)[v]);
                }
          }
      }

    std::vector<uint32_t> constrained_indices(matrix_free.get_constrained_dofs().begin(),
                                              matrix_free.get_constrained_dofs().end()); // operator_ns.cc:123-124
    // owned entries this rank exports, neighbour by neighbour (Partitioner::import_indices)
    std::vector<uint32_t> export_indices;
    for (const auto &range : part.import_indices())
      for (unsigned int i = range.first; i < range.second; ++i)
        export_indices.push_back(i);

    // GMG-LS: owned dofs on the refinement edge of this level (operator_ns.cc:131-152, :1436-1455)
    std::vector<uint32_t> edge_constrained_indices;
    bool                  has_edge = false;
    if (mg_level != numbers::invalid_unsigned_int)
      {
        std::vector<IndexSet> refinement_edge_indices(dof_handler.get_triangulation().n_global_levels());
        for (unsigned int l = 0; l < refinement_edge_indices.size(); ++l)
          refinement_edge_indices[l] = IndexSet(dof_handler.n_dofs(l));
        MGTools::extract_inner_interface_dofs(dof_handler, refinement_edge_indices);
        const IndexSet &owned = dof_handler.locally_owned_mg_dofs(mg_level);
        for (const auto g : refinement_edge_indices[mg_level])
          if (owned.is_element(g))
            edge_constrained_indices.push_back(owned.index_within_set(g));
        has_edge = Utilities::MPI::max(edge_constrained_indices.size(), dof_handler.get_communicator()) > 0;
      }

    // ---- boundary faces with outflow terms: what MatrixFree::loop hands to do_vmult_boundary -------------
    // Face quadrature in the library's order (QGauss(p+1) in the tangential directions, ascending direction
    // fastest), geometry from FEValues on the cell at the projected points: n = J^-T n_ref / |.|,
    // JxW = |det J| |J^-T n_ref| w_q (what update_normal_vectors | update_JxW_values give on the face).
    std::vector<uint32_t> face_cell, face_no, face_kind;
    std::vector<double>   face_normal, face_jxw, face_inv_jac, face_target;
    if (!(all_outflow_bcs_cut.empty() && all_outflow_bcs_nitsche.empty()))
      {
        const QGauss<1>           q1(degree + 1);
        const unsigned int        n1 = q1.size();
        std::vector<unsigned int> cell_number; // position of (batch, lane) in the cell order used above
        std::uint64_t             k = 0;
        for (unsigned int b = 0; b < matrix_free.n_cell_batches(); ++b)
          for (unsigned int v = 0; v < matrix_free.n_active_entries_per_cell_batch(b); ++v, ++k)
            {
              const auto cell = matrix_free.get_cell_iterator(b, v);
              for (const unsigned int f : cell->face_indices())
                {
                  if (!cell->face(f)->at_boundary())
                    continue;
                  const auto id   = cell->face(f)->boundary_id();
                  const bool cut  = all_outflow_bcs_cut.count(id) > 0;
                  const auto nit  = all_outflow_bcs_nitsche.find(id);
                  if (!cut && nit == all_outflow_bcs_nitsche.end())
                    continue;
                  const unsigned int dir = f / 2, side = f % 2;
                  std::vector<Point<dim>> pts;
                  std::vector<double>     wts;
                  const unsigned int      nqf = Utilities::pow(n1, dim - 1);
                  for (unsigned int q = 0; q < nqf; ++q)
                    {
                      Point<dim>   xi;
                      double       w  = 1;
                      unsigned int qq = q;
                      for (unsigned int e = 0; e < dim; ++e)
                        if (e == dir)
                          xi[e] = side;
                        else
                          {
                            xi[e] = q1.point(qq % n1)[0];
                            w *= q1.weight(qq % n1);
                            qq /= n1;
                          }
                      pts.push_back(xi);
                      wts.push_back(w);
                    }
                  FEValues<dim> fev(mapping, dof_handler.get_fe(), Quadrature<dim>(pts, wts),
                                    update_inverse_jacobians | update_jacobians | update_quadrature_points);
                  fev.reinit(typename Triangulation<dim>::cell_iterator(cell));
                  face_cell.push_back(k), face_no.push_back(f), face_kind.push_back(cut ? 1 : 2);
                  for (unsigned int q = 0; q < nqf; ++q)
                    {
                      const auto    &Ji = fev.inverse_jacobian(q); // Ji[e][j] = (J^-1)_{e j}
                      Tensor<1, dim> nn;
                      for (unsigned int j = 0; j < dim; ++j)
                        nn[j] = Ji[dir][j] * (side ? 1.0 : -1.0);
                      const double ln = nn.norm();
                      for (unsigned int j = 0; j < dim; ++j)
                        face_normal.push_back(nn[j] / ln);
                      face_jxw.push_back(std::abs(fev.jacobian(q).determinant()) * ln * wts[q]);
                      for (unsigned int e = 0; e < dim; ++e)
                        for (unsigned int j = 0; j < dim; ++j)
                          face_inv_jac.push_back(Ji[e][j]);
                      for (unsigned int c = 0; c < dim; ++c) // operator_ns.cc:495-521
                        face_target.push_back(cut ? 0.0 : nit->second->value(fev.quadrature_point(q), c));
                    }
                }
            }
      }

    glsb_desc d{};
    d.abi_version = GLSB_ABI_VERSION;
    cudaGetDevice(&d.device);
    d.n_outflow_faces = face_cell.size();
    d.face_cell = face_cell.data(), d.face_no = face_no.data(), d.face_kind = face_kind.data();
    d.face_normal = face_normal.data(), d.face_jxw = face_jxw.data(), d.face_inv_jac = face_inv_jac.data();
    d.face_target_velocity = face_target.data();
    d.n_edge_constrained_indices   = edge_constrained_indices.size();
    d.edge_constrained_indices     = edge_constrained_indices.data();
    d.has_edge_constrained_indices = has_edge;
    d.dim = dim, d.degree = degree, d.number_type = std::is_same<Number, double>::value ? GLSB_F64 : GLSB_F32;
    d.increment_form = increment_form, d.consider_time_derivative = consider_time_derivative;
    d.cell_wise_stabilization = cell_wise_stabilization, d.time_order = time_integrator_data.get_order();
    d.nu = nu, d.c1 = c_1, d.c2 = c_2, d.theta = time_integrator_data.get_theta();
    d.n_cells = n_cells, d.n_owned = part.locally_owned_size(), d.n_ghost = part.n_ghost_indices();
    d.dof_indices = dof_indices.data();
    d.n_constraint_rows = row_dof.size(), d.row_dof = row_dof.data(), d.row_ptr = row_ptr.data();
    d.entry_col = entry_col.data(), d.entry_val = entry_val.data();
    d.n_constrained_indices = constrained_indices.size(), d.constrained_indices = constrained_indices.data();
    d.geometry_type = all_cartesian ? GLSB_GEOM_CARTESIAN : GLSB_GEOM_GENERAL;
    d.inv_jac = inv_jac.data(), d.jxw = jxw.data(), d.cell_h_min = h_min.data(), d.cell_measure = measure.data();
    d.n_export = export_indices.size(), d.export_indices = export_indices.data();
    AssertThrow(glsb_create(&d, &op) == 0, ExcMessage(glsb_last_error(nullptr)));
    n_owned = d.n_owned, n_ghost = d.n_ghost;
    n_local = d.n_owned + d.n_ghost;
    has_edge_constrained_indices = has_edge;

    // ---- streams, exchange buffers, NCCL communicator -----------------------------------------------------
    partitioner = matrix_free.get_vector_partitioner();
    cuda_check(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    cuda_check(cudaStreamCreateWithFlags(&comm_stream, cudaStreamNonBlocking));
    cuda_check(cudaEventCreateWithFlags(&ev_compute, cudaEventDisableTiming));
    cuda_check(cudaEventCreateWithFlags(&ev_comm, cudaEventDisableTiming));
    n_export = export_indices.size();
    if (n_export > 0)
      {
        cuda_check(cudaMalloc(&d_export, n_export * sizeof(Number)));
        cuda_check(cudaMalloc(&d_import, n_export * sizeof(Number)));
      }
    if (partitioner->n_mpi_processes() > 1)
      nccl = shared_communicator(partitioner->get_mpi_communicator());
  }

  ~NavierStokesOperatorB200() override
  {
    glsb_destroy(op);
    for (auto &m : mirrors)
      {
        cudaFree(m.second.d);
        if (m.second.registered)
          glsb_host_unregister(const_cast<Number *>(m.first));
      }
    cudaFree(d_export), cudaFree(d_import);
    cudaEventDestroy(ev_compute), cudaEventDestroy(ev_comm);
    cudaStreamDestroy(stream), cudaStreamDestroy(comm_stream);
  }

  // ---- OperatorBase<Number> -------------------------------------------------------------------
  types::global_dof_index m() const override { return cpu_operator.m(); }
  const AffineConstraints<Number> &get_constraints() const override { return cpu_operator.get_constraints(); }
  std::vector<std::vector<bool>>   extract_constant_modes() const override { return cpu_operator.extract_constant_modes(); }
  // assembled matrix for the coarse-level AMG/direct solvers stays on the CPU operator (Trilinos)
  const SparseMatrixType &get_system_matrix() const override { return cpu_operator.get_system_matrix(); }
  void initialize_dof_vector(VectorType<Number> &vec) const override { matrix_free.initialize_dof_vector(vec); }
  void invalidate_system() override
  {
    cpu_operator.invalidate_system();
    check(glsb_invalidate_system(op));
  }

  void vmult(VectorType<Number> &dst, const VectorType<Number> &src) const override
  {
    MyScope scope(timer, "ns::vmult"); // same section names as the reference (operator_ns.cc:689)
    const double w = time_integrator_data.get_primary_weight();
    if (!device_vectors && !has_edge_constrained_indices)
      {
        // host VectorType (config.h:9-10): upload, cells and download pipelined inside the library
        vmult_host_vectors(dst, src, w);
        return;
      }
    check(glsb_edge_begin(op, const_cast<void *>(dev(src)), stream)); // operator_ns.cc:692-700
    check(glsb_vmult_begin(op, dev(dst), stream));
    update_ghost_values_start(src); // NCCL send/recv of the export block into the ghost block
    check(glsb_vmult_cells_part(op, dev(dst), dev(src), w, GLSB_CELLS_INTERIOR, 0, 2, stream));
    update_ghost_values_finish();
    check(glsb_vmult_cells(op, dev(dst), dev(src), w, GLSB_CELLS_BOUNDARY, stream));
    compress_start(dst);
    check(glsb_vmult_cells_part(op, dev(dst), dev(src), w, GLSB_CELLS_INTERIOR, 1, 2, stream));
    compress_finish(dst); // glsb_unpack_add
    check(glsb_vmult_finish(op, dev(dst), dev(src), stream));
    check(glsb_edge_finish(op, dev(dst), const_cast<void *>(dev(src)), stream)); // operator_ns.cc:724-731
    finish(dst);
  }

  void vmult_interface_down(VectorType<Number> &dst, const VectorType<Number> &src) const override
  {
    MyScope scope(timer, "ns::vmult_interface_down");
    const double w = time_integrator_data.get_primary_weight();
    check(glsb_vmult_begin(op, dev(dst), stream));
    update_ghost_values_start(src);
    update_ghost_values_finish();
    check(glsb_vmult_cells(op, dev(dst), dev(src), w, GLSB_CELLS_ALL, stream));
    compress_start(dst);
    compress_finish(dst);
    check(glsb_vmult_finish(op, dev(dst), dev(src), stream));
    finish(dst);
  }

  void vmult_interface_up(VectorType<Number> &dst, const VectorType<Number> &src) const override
  {
    MyScope scope(timer, "ns::vmult_interface_up");
    check(glsb_vmult_begin(op, dev(dst), stream)); // dst = 0
    if (has_edge_constrained_indices)
      {
        VectorType<Number> src_cpy;
        src_cpy.reinit(src, /*omit_zeroing_entries=*/true);
        check(glsb_edge_extract(op, dev(src_cpy), dev(src), stream)); // operator_ns.cc:768-774
        update_ghost_values_start(src_cpy);
        update_ghost_values_finish();
        check(glsb_vmult_cells(op, dev(dst), dev(src_cpy), time_integrator_data.get_primary_weight(), GLSB_CELLS_ALL,
                               stream));
        compress_start(dst);
        compress_finish(dst);
      }
    finish(dst);
  }

  void set_linearization_point(const VectorType<Number> &vec) override
  {
    MyScope scope(timer, "ns::set_linearization_point");
    AssertThrow(vec.has_ghost_elements() == false, ExcInternalError()); // operator_ns.cc:589-591
    update_ghost_values_start(vec);
    update_ghost_values_finish();
    check(glsb_set_linearization_point(op, dev(vec), time_integrator_data.get_current_dt(), stream));
    end_operation();
  }

  void set_previous_solution(const SolutionHistory<Number> &history) override
  {
    MyScope scope(timer, "ns::set_previous_solution");
    const unsigned int order = time_integrator_data.get_order();
    if (order == 0)
      return;
    std::vector<const void *> ptr;
    for (unsigned int i = 0; i <= order; ++i)
      {
        if (i > 0)
          {
            update_ghost_values_start(history.get_vectors()[i]);
            update_ghost_values_finish();
          }
        ptr.push_back(dev(history.get_vectors()[i]));
      }
    check(glsb_set_previous_solution(op, ptr.data(), time_integrator_data.get_weights().data(), order, stream));
    end_operation();
  }

  void evaluate_residual(VectorType<Number> &dst, const VectorType<Number> &src) const override
  {
    MyScope scope(timer, "ns::evaluate_residual");
    VectorType<Number> tmp = src;
    constraints_inhomogeneous.distribute(tmp); // operator_ns.cc:655-656
    update_ghost_values_start(tmp);
    update_ghost_values_finish();
    check(glsb_evaluate_residual(op, dev(dst), dev(tmp), time_integrator_data.get_primary_weight(), stream));
    compress_start(dst);
    compress_finish(dst);
    finish(dst);
  }

  void evaluate_rhs(VectorType<Number> &dst) const override
  {
    VectorType<Number> src;
    src.reinit(dst);
    evaluate_residual(dst, src);
  }

  void compute_inverse_diagonal(VectorType<Number> &diagonal) const override
  {
    MyScope scope(timer, "ns::compute_inverse_diagonal");
    matrix_free.initialize_dof_vector(diagonal);
    check(glsb_diagonal_cells(op, dev(diagonal), time_integrator_data.get_primary_weight(), stream));
    compress_start(diagonal);
    compress_finish(diagonal);
    check(glsb_diagonal_finish(op, dev(diagonal), stream));
    finish(diagonal);
  }

  double get_max_u(const VectorType<Number> &vec) const override
  {
    update_ghost_values_start(vec);
    update_ghost_values_finish();
    double v = 0;
    check(glsb_get_max_u(op, dev(vec), &v, stream)); // synchronises `stream`
    end_operation();
    return Utilities::MPI::max(v, MPI_COMM_WORLD); // operator_ns.cc:567
  }

private:
  void check(const int rc) const { AssertThrow(rc == 0, ExcMessage(glsb_last_error(op))); }

  static void cuda_check(const cudaError_t e) { AssertThrow(e == cudaSuccess, ExcMessage(cudaGetErrorString(e))); }
  static void nccl_check(const ncclResult_t r) { AssertThrow(r == ncclSuccess, ExcMessage(ncclGetErrorString(r))); }

  // ---- vector residency ---------------------------------------------------------------------------------
  // MemorySpace::Default vectors (Kokkos-CUDA deal.II, alias flipped in config.h:9-10): get_values() is
  // device memory and is passed through.  Host vectors (the reference's default alias): every host array gets
  // a device mirror of n_owned + n_ghost values, page-locked once (glsb_host_register); dev(const) uploads the
  // owned block at its first use inside an operation, finish() downloads the owned block of the result.
  static constexpr bool device_vectors =
    std::is_same<typename VectorType<Number>::memory_space, MemorySpace::Default>::value;

  struct Mirror
  {
    Number       *d          = nullptr;
    unsigned long uploaded   = 0; // epoch of the last upload
    bool          registered = false;
  };

  Mirror &mirror_of(const Number *host) const
  {
    auto it = mirrors.find(host);
    if (it == mirrors.end())
      {
        Mirror m;
        cuda_check(cudaMalloc(&m.d, n_local * sizeof(Number)));
        cuda_check(cudaMemsetAsync(m.d, 0, n_local * sizeof(Number), stream));
        m.registered = glsb_host_register(const_cast<Number *>(host), n_owned * sizeof(Number)) == 0;
        it           = mirrors.emplace(host, m).first;
      }
    return it->second;
  }

  const void *dev(const VectorType<Number> &v) const
  {
    if (device_vectors)
      return v.get_values();
    Mirror &m = mirror_of(v.get_values());
    if (m.uploaded != epoch)
      {
        cuda_check(cudaMemcpyAsync(m.d, v.get_values(), n_owned * sizeof(Number), cudaMemcpyHostToDevice, stream));
        m.uploaded = epoch;
      }
    return m.d;
  }

  // results: no upload, the kernels overwrite the owned block
  void *dev(VectorType<Number> &v) const
  {
    if (device_vectors)
      return v.get_values();
    Mirror &m  = mirror_of(v.get_values());
    m.uploaded = epoch;
    return m.d;
  }

  // the operation is over: make the result visible to the (host-side) caller
  void finish(VectorType<Number> &v) const
  {
    if (!device_vectors)
      cuda_check(cudaMemcpyAsync(v.get_values(), mirror_of(v.get_values()).d, n_owned * sizeof(Number),
                                 cudaMemcpyDeviceToHost, stream));
    end_operation();
  }

  void end_operation() const
  {
    cuda_check(cudaStreamSynchronize(stream));
    ++epoch; // host vectors may change before the next call: upload again
  }

  // host vectors, no edge indices: the pipelined entry points
  void vmult_host_vectors(VectorType<Number> &dst, const VectorType<Number> &src, const double w) const
  {
    if (nccl == nullptr && n_ghost == 0)
      check(glsb_vmult_host(op, dst.get_values(), src.get_values(), w, stream));
    else
      {
        Mirror &ms = mirror_of(src.get_values()), &md = mirror_of(dst.get_values());
        check(glsb_vmult_host_begin(op, md.d, ms.d, dst.get_values(), src.get_values(), w, stream));
        ms.uploaded = md.uploaded = epoch; // _begin uploaded src and zeroed dst
        update_ghost_values_start(src);
        update_ghost_values_finish();
        check(glsb_vmult_cells(op, md.d, ms.d, w, GLSB_CELLS_BOUNDARY, stream));
        compress_start(dst);
        compress_finish(dst);
        check(glsb_vmult_host_finish(op, md.d, dst.get_values(), stream));
      }
    end_operation();
  }

  // ---- ghost exchange over NCCL ---------------------------------------------------------------------------
  // update_ghost_values: pack the owned entries other ranks read (Partitioner::import_indices, grouped by
  // import_targets) and send them; receive the ghost block slice by slice from ghost_targets.
  // compress(add): the same lists the other way round, added by glsb_unpack_add.  One MPI rank per GPU; the
  // communicator is created once per process from an id broadcast over MPI.
  static ncclComm_t shared_communicator(const MPI_Comm comm)
  {
    static ncclComm_t shared = nullptr;
    if (shared == nullptr)
      {
        int rank = 0, size = 1;
        MPI_Comm_rank(comm, &rank);
        MPI_Comm_size(comm, &size);
        ncclUniqueId id;
        if (rank == 0)
          nccl_check(ncclGetUniqueId(&id));
        MPI_Bcast(&id, sizeof(id), MPI_BYTE, 0, comm);
        nccl_check(ncclCommInitRank(&shared, size, id, rank));
      }
    return shared;
  }

  static constexpr ncclDataType_t nccl_type = std::is_same<Number, double>::value ? ncclFloat64 : ncclFloat32;

  void update_ghost_values_start(const VectorType<Number> &v) const
  {
    if (nccl == nullptr)
      return;
    Number *vec = static_cast<Number *>(const_cast<void *>(dev(v)));
    cuda_check(cudaEventRecord(ev_compute, stream));
    cuda_check(cudaStreamWaitEvent(comm_stream, ev_compute, 0));
    if (n_export > 0)
      check(glsb_pack_export(op, d_export, vec, comm_stream));
    nccl_check(ncclGroupStart());
    std::size_t off = 0;
    for (const auto &t : partitioner->import_targets()) // (rank, number of owned entries it reads)
      {
        nccl_check(ncclSend(d_export + off, t.second, nccl_type, t.first, nccl, comm_stream));
        off += t.second;
      }
    off = n_owned;
    for (const auto &t : partitioner->ghost_targets()) // (owner rank, number of ghosts it owns), ghost block order
      {
        nccl_check(ncclRecv(vec + off, t.second, nccl_type, t.first, nccl, comm_stream));
        off += t.second;
      }
    nccl_check(ncclGroupEnd());
    cuda_check(cudaEventRecord(ev_comm, comm_stream));
  }

  void update_ghost_values_finish() const
  {
    if (nccl != nullptr)
      cuda_check(cudaStreamWaitEvent(stream, ev_comm, 0));
  }

  void compress_start(VectorType<Number> &v) const
  {
    if (nccl == nullptr)
      return;
    Number *vec = static_cast<Number *>(dev(v));
    cuda_check(cudaEventRecord(ev_compute, stream));
    cuda_check(cudaStreamWaitEvent(comm_stream, ev_compute, 0));
    nccl_check(ncclGroupStart());
    std::size_t off = n_owned;
    for (const auto &t : partitioner->ghost_targets())
      {
        nccl_check(ncclSend(vec + off, t.second, nccl_type, t.first, nccl, comm_stream));
        off += t.second;
      }
    off = 0;
    for (const auto &t : partitioner->import_targets())
      {
        nccl_check(ncclRecv(d_import + off, t.second, nccl_type, t.first, nccl, comm_stream));
        off += t.second;
      }
    nccl_check(ncclGroupEnd());
    cuda_check(cudaEventRecord(ev_comm, comm_stream));
  }

  void compress_finish(VectorType<Number> &v) const
  {
    if (nccl == nullptr)
      return;
    Number *vec = static_cast<Number *>(dev(v));
    cuda_check(cudaStreamWaitEvent(stream, ev_comm, 0));
    if (n_export > 0)
      check(glsb_unpack_add(op, vec, d_import, stream));
    if (n_ghost > 0) // compress() leaves the ghost block zeroed
      cuda_check(cudaMemsetAsync(vec + n_owned, 0, n_ghost * sizeof(Number), stream));
  }

  const AffineConstraints<Number>  &constraints_inhomogeneous;
  const TimeIntegratorData         &time_integrator_data;
  NavierStokesOperator<dim, Number> cpu_operator; // system matrix / constant modes / constraints only
  MatrixFree<dim, Number>           matrix_free;
  glsb_op                          *op      = nullptr;
  bool                              has_edge_constrained_indices = false;
  std::uint64_t                     n_local = 0, n_owned = 0, n_ghost = 0, n_export = 0;
  cudaStream_t                      stream = nullptr, comm_stream = nullptr;
  cudaEvent_t                       ev_compute = nullptr, ev_comm = nullptr;
  ncclComm_t                        nccl   = nullptr;
  Number                           *d_export = nullptr, *d_import = nullptr; // packed owned entries out / in
  std::shared_ptr<const Utilities::MPI::Partitioner> partitioner;
  mutable std::map<const Number *, Mirror>           mirrors; // host vectors: device mirrors by host array
  mutable unsigned long                              epoch = 1;
  mutable MyTimerOutput             timer;
};

} // namespace glsb

#endif // deal.II available
