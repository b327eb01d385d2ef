/* glsb200.h -- C ABI of libglsb200.so: the B200 (sm_100a) implementation of the
 * matrix-free GLS/SUPG-PSPG Navier-Stokes operator of peterrum/dealii-ns-gls.
 *
 * Every entry point replaces one virtual of the reference's OperatorBase<Number>
 * (include/operator_base.h:13-73) as implemented by NavierStokesOperator<dim,Number>
 * (include/operator_ns.h:17-189, include/operator_ns.cc); the reference-side
 * binding (a NavierStokesOperatorB200 subclass that forwards to these calls) is
 * shown in INTEGRATION.md and dealii_ns_gls_b200/cpp/dealii_adapter.h.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.
 *   - pointers in glsb_desc are HOST pointers, read during glsb_create() and not
 *     retained; vectors passed to the compute calls are DEVICE pointers to
 *     n_owned + n_ghost values of the operator's number type (the layout of
 *     LinearAlgebra::distributed::Vector: owned block, then ghost block).
 *   - every compute call enqueues work on the caller's stream (a cudaStream_t
 *     passed as void*, NULL = default stream) and returns without synchronising
 *     unless stated otherwise.
 *   - return value 0 = ok; non-zero = error, text via glsb_last_error().
 *     (The reference signals errors with AssertThrow, e.g. operator_ns.cc:287.)
 *   - there is no CPU fallback: glsb_create() fails if no CUDA device is usable.
 */
#ifndef GLSB200_H
#define GLSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLSB_ABI_VERSION 3

typedef struct glsb_op glsb_op;

enum { GLSB_F64 = 0, GLSB_F32 = 1 };
enum { GLSB_GEOM_CARTESIAN = 0, GLSB_GEOM_GENERAL = 2 };
enum { GLSB_CELLS_ALL = 0, GLSB_CELLS_INTERIOR = 1, GLSB_CELLS_BOUNDARY = 2 };

/* A dof index with this bit set refers to constraint row (index & 0x7fffffff)
 * instead of a vector entry (how read_dof_values / distribute_local_to_global
 * resolve AffineConstraints, operator_ns.cc:819-828). */
#define GLSB_CONSTRAINED_BIT 0x80000000u

typedef struct glsb_desc
{
  int32_t abi_version; /* GLSB_ABI_VERSION */
  int32_t device;      /* CUDA device ordinal */

  /* discretisation: FESystem(FE_Q(degree), dim+1), QGauss(degree+1)
   * (main.cc:251, performance.cc:33-38) */
  int32_t dim;         /* 2 or 3 */
  int32_t degree;      /* 1..4 */
  int32_t number_type; /* GLSB_F64 (Krylov operator) or GLSB_F32 (multigrid levels, config.h:6-7) */

  /* operator flags, ctor arguments of NavierStokesOperator (operator_ns.h:24-41) */
  int32_t increment_form;
  int32_t consider_time_derivative; /* as passed; the library ANDs it with time_order > 0 (operator_ns.cc:97-98) */
  int32_t cell_wise_stabilization;
  int32_t time_order; /* TimeIntegratorData::get_order(): 0 = "none" */
  double  nu, c1, c2;
  double  theta; /* TimeIntegratorData::get_theta() */

  /* cells and dofs of this rank (MatrixFree / DoFHandler / Partitioner) */
  uint64_t        n_cells;
  uint64_t        n_owned;
  uint64_t        n_ghost;
  const uint32_t *dof_indices; /* [n_cells][(dim+1)*(degree+1)^dim], component-blocked,
                                  lexicographic (x fastest) like FEEvaluation; local vector
                                  indices, or GLSB_CONSTRAINED_BIT | row */

  /* constraints_homogeneous (MatrixFree constraint index 0, operator_ns.cc:107-108) as CSR rows */
  uint32_t        n_constraint_rows;
  const uint32_t *row_dof;   /* [n_rows] vector index of the constrained dof */
  const uint32_t *row_ptr;   /* [n_rows + 1] */
  const uint32_t *entry_col; /* master vector indices */
  const double   *entry_val; /* weights */
  /* matrix_free.get_constrained_dofs() (operator_ns.cc:123-124): rows where vmult is the identity */
  uint32_t        n_constrained_indices;
  const uint32_t *constrained_indices;

  /* geometry (MappingInfo): Cartesian -> inv_jac[n_cells][dim] (diagonal of J^-1),
   * jxw[n_cells] (det J); general -> inv_jac[n_cells][n_q][dim][dim] with
   * inv_jac[..][e][j] = (J^-1)_{e j}, jxw[n_cells][n_q] = det J * w_q */
  int32_t       geometry_type;
  const double *inv_jac;
  const double *jxw;
  /* cell->minimum_vertex_distance() and cell->measure() (operator_ns.cc:373-374, :399) */
  const double *cell_h_min;
  const double *cell_measure;

  /* ghost exchange lists of the vector partitioner (owned-local indices this rank
   * exports, concatenated over neighbours); may be empty */
  uint64_t        n_export;
  const uint32_t *export_indices;

  /* multigrid with local smoothing (GMG-LS, main.cc:569-732): owned indices on the refinement edge of
   * this level (edge_constrained_indices, operator_ns.cc:131-152); has_edge_constrained_indices is the
   * MPI::max over all ranks of "the list is not empty" (operator_ns.cc:149-151).  Empty for global
   * coarsening and for the fine-level operator. */
  uint32_t        n_edge_constrained_indices;
  const uint32_t *edge_constrained_indices;
  int32_t         has_edge_constrained_indices;

  /* boundary faces with outflow terms (do_vmult_boundary, operator_ns.cc:1195-1301; the ctor arguments
   * all_outflow_bcs_cut / all_outflow_bcs_nitsche, operator_ns.h:33-35, resolved to faces by the caller like
   * MatrixFree::loop resolves boundary ids): kind 1 = "cut" (v, beta min(0, U.n) u), 2 = Nitsche.  Face
   * quadrature: QGauss(degree + 1) in the tangential directions, ascending direction fastest; the arrays are
   * what MatrixFree's face MappingInfo holds (mapping_update_flags_boundary_faces, operator_ns.cc:113-117).
   * May be empty (all five BASELINE configs). */
  uint32_t        n_outflow_faces;
  const uint32_t *face_cell;            /* [n_faces] cell the face belongs to */
  const uint32_t *face_no;              /* [n_faces] 2 * direction + side (GeometryInfo face number) */
  const uint32_t *face_kind;            /* [n_faces] 1 or 2 */
  const double   *face_normal;          /* [n_faces][n_qf][dim] unit outer normal */
  const double   *face_jxw;             /* [n_faces][n_qf] */
  const double   *face_inv_jac;         /* [n_faces][n_qf][dim][dim], [e][j] = (J^-1)_{e j} */
  const double   *face_target_velocity; /* [n_faces][n_qf][dim] Nitsche target (operator_ns.cc:495-521) or NULL */
} glsb_desc;

/* ---- life cycle -------------------------------------------------------- */

/* replaces the constructor NavierStokesOperator::NavierStokesOperator (operator_ns.cc:68-153) */
/* Environment switches read by glsb_create:
 *   GLSB_DETERMINISTIC=1  bit-reproducible cell loops for single-rank operators: the cells are coloured over the
 *                         vector entries they scatter to and run colour by colour, so every entry is summed in a
 *                         fixed order instead of in the arrival order of the atomics (about 25 % slower; the
 *                         register-tiled vmult is reproducible on hanging-node meshes too, the generic kernels on
 *                         meshes without weighted rows). */
int  glsb_create(const glsb_desc *desc, glsb_op **out);
void glsb_destroy(glsb_op *op);
/* last error text of op (or of the last failed glsb_create when op == NULL) */
const char *glsb_last_error(const glsb_op *op);
/* invalidate_system (operator_ns.cc:227-232) */
int glsb_invalidate_system(glsb_op *op);

/* ---- operator application ---------------------------------------------- */

/* vmult (operator_ns.cc:684-732): dst = A(U) src, identity on constrained rows.
 * weight = time_integrator_data.get_primary_weight(), read at call time like
 * the reference does (operator_ns.cc:958, :1071).  Single-rank form: zeroes dst,
 * loops over all cells, copies the constrained rows. */
int glsb_vmult(glsb_op *op, void *dst, const void *src, double weight, void *stream);

/* vmult for callers whose vectors live in HOST memory (the reference's VectorType is a host
 * LinearAlgebra::distributed::Vector, config.h:9-10): src_host / dst_host are host pointers to n_owned
 * values (page-locked for full speed, see glsb_host_register); upload, cell kernels and download run as a
 * chunked pipeline over PCIe.  Ordered after the work on `stream`; `stream` waits for the last download.
 * Single-rank operators only (n_ghost == 0). */
int glsb_vmult_host(glsb_op *op, void *dst_host, const void *src_host, double weight, void *stream);
/* The same pipeline for partitioned operators (n_ghost > 0), split around the ghost exchange the host layer
 * drives: _begin zeroes d_dst and pipelines upload / interior cells / speculative download in chunks (d_src,
 * d_dst: device vectors of n_owned + n_ghost values owned by the caller; dst_host page-locked); the caller then
 * imports the ghosts of d_src, runs glsb_vmult_cells(GLSB_CELLS_BOUNDARY) and compress(add)s d_dst; _finish
 * re-sends the entries that changed after their download (later chunks, boundary cells, compress).  The
 * identity on constrained rows (operator_ns.cc:719-721) is applied chunk by chunk inside _begin. */
int glsb_vmult_host_begin(glsb_op *op, void *d_dst, void *d_src, void *dst_host, const void *src_host, double weight,
                          void *stream);
int glsb_vmult_host_finish(glsb_op *op, void *d_dst, void *dst_host, void *stream);
/* cudaHostRegister / cudaHostUnregister for a caller-owned host vector */
int glsb_host_register(void *ptr, uint64_t bytes);
int glsb_host_unregister(void *ptr);

/* The three pieces of cell_loop(..., zero_dst = true) for the host layer that
 * overlaps the ghost exchange with the interior cells (operator_ns.cc:703-708):
 *   glsb_vmult_begin  : zero dst (owned + ghost)
 *   glsb_vmult_cells  : which = GLSB_CELLS_INTERIOR (cells without ghost dofs),
 *                       GLSB_CELLS_BOUNDARY or GLSB_CELLS_ALL
 *   glsb_vmult_finish : dst[constrained] = src[constrained] (operator_ns.cc:719-721) */
int glsb_vmult_begin(glsb_op *op, void *dst, void *stream);
int glsb_vmult_cells(glsb_op *op, void *dst, const void *src, double weight, int which, void *stream);
int glsb_vmult_finish(glsb_op *op, void *dst, const void *src, void *stream);
/* chunk `part` of `n_parts` equal pieces of the cell class `which`: lets the host layer put the ghost
 * import behind one half of the interior cells and compress(add) behind the other */
int glsb_vmult_cells_part(glsb_op *op, void *dst, const void *src, double weight, int which, int part,
                          int n_parts, void *stream);
/* Edge-constrained dofs around vmult (operator_ns.cc:692-700, :724-731).  glsb_vmult does both itself; a
 * host layer that drives the pieces calls glsb_edge_begin before the ghost import (saves src at the edge
 * indices and zeroes it there -- the reference const_casts src the same way) and glsb_edge_finish after
 * glsb_vmult_finish (src restored, dst = saved src value at the edge indices). */
int glsb_edge_begin(glsb_op *op, void *src, void *stream);
int glsb_edge_finish(glsb_op *op, void *dst, void *src, void *stream);
/* vmult_interface_down (operator_ns.cc:734-752): the cell loop with zeroed dst plus the identity on
 * constrained rows, without the edge handling of vmult. */
int glsb_vmult_interface_down(glsb_op *op, void *dst, const void *src, double weight, void *stream);
/* vmult_interface_up (operator_ns.cc:754-787): dst = A * (src restricted to the edge indices); dst = 0 if
 * no rank has edge indices; no identity on constrained rows.  Single-rank form. */
int glsb_vmult_interface_up(glsb_op *op, void *dst, const void *src, double weight, void *stream);
/* cpy = 0 except cpy[edge] = src[edge] (operator_ns.cc:768-774), for a host layer that imports the ghosts
 * of cpy itself before glsb_vmult_cells */
int glsb_edge_extract(glsb_op *op, void *cpy, const void *src, void *stream);

/* leave n_sms multiprocessors free when launching the (persistent) cell kernels, so that
 * communication kernels (NCCL send/recv) can run next to them; 0 = use the whole GPU */
int glsb_set_sm_reserve(glsb_op *op, int n_sms);

/* evaluate_residual (operator_ns.cc:648-682): src must already carry the
 * inhomogeneous boundary values (constraints_inhomogeneous.distribute, :655-656)
 * and imported ghost values; dst = -(C^T F(src)), zero on constrained rows.
 * evaluate_rhs (:622-646) is the same call on distribute(0). */
int glsb_evaluate_residual(glsb_op *op, void *dst, const void *src_with_bc, double weight, void *stream);
int glsb_evaluate_residual_cells(glsb_op *op, void *dst, const void *src_with_bc, double weight,
                                 int which, void *stream);

/* ---- state -------------------------------------------------------------- */

/* set_linearization_point (operator_ns.cc:570-620) + compute_penalty_parameters
 * (:322-420): vec with imported ghost values; dt = get_current_dt(). */
int glsb_set_linearization_point(glsb_op *op, const void *vec, double dt, void *stream);

/* set_previous_solution (operator_ns.cc:234-320): history[i], i = 0..order, device
 * vectors with imported ghosts; weights = get_weights(). No-op if time_order == 0. */
int glsb_set_previous_solution(glsb_op *op, const void *const *history, const double *weights,
                               int order, void *stream);

/* compute_inverse_diagonal (operator_ns.cc:195-225): diag = (|d| > 1e-10 ? 1/d : 1) of
 * diag(C^T A C), 1 on constrained rows.  With ghosts, call glsb_diagonal_cells then
 * compress(add) then glsb_diagonal_finish. */
int glsb_compute_inverse_diagonal(glsb_op *op, void *diag, double weight, void *stream);
int glsb_diagonal_cells(glsb_op *op, void *diag, double weight, void *stream);
int glsb_diagonal_finish(glsb_op *op, void *diag, void *stream);

/* get_system_matrix (operator_ns.cc:1303-1434, used by the coarse-grid solvers, multigrid.cc:395-425): the
 * matrix of vmult, dense, row-major n x n doubles on the device (A[i * n + j]), obtained column by column
 * from the operator's own cell loop (MatrixFreeTools::compute_matrix does the same cell-wise): constrained
 * rows carry 1 on the diagonal.  Meant for the coarsest multigrid level only.  n = n_owned + n_ghost: a
 * partitioned operator returns the matrix of its own cells over its local dofs, which the host layer sums over
 * the ranks (identity rows of owned constrained dofs only). */
int glsb_get_system_matrix(glsb_op *op, double *A_dev, double weight, void *stream);

/* get_max_u (operator_ns.cc:530-568): max over the local quadrature points of |u|;
 * synchronises the stream; the MPI::max over ranks stays with the caller. */
int glsb_get_max_u(glsb_op *op, const void *vec, double *out_host, void *stream);

/* ---- relaxation smoother on the device -------------------------------
 * The multigrid smoother of the reference is deal.II's PreconditionRelaxation<OperatorBase<MGNumber>,
 * DiagonalMatrix> (include/multigrid.h:67-69): damped point-Jacobi x <- x + omega D^-1 (b - A x), 5 sweeps,
 * omega from a 20-step power iteration on D^-1 A with smoothing range 20 (multigrid.h:30-32,
 * multigrid.cc:290-304, :353-370).  deal.II runs it as op.vmult + separate host vector operations; here all
 * sweeps of a call are enqueued on the stream back to back (vmult kernel + one fused update kernel per
 * sweep), nothing returns to the host in between. */
/* PreconditionRelaxation::vmult: n_iterations sweeps from a ZERO initial guess (the first sweep is
 * dst = omega * inv_diag * src and needs no operator application) */
int glsb_relaxation_vmult(glsb_op *op, void *dst, const void *src, const void *inv_diag, double omega,
                          int n_iterations, double weight, void *stream);
/* PreconditionRelaxation::step: n_iterations sweeps starting from the current dst */
int glsb_relaxation_step(glsb_op *op, void *dst, const void *src, const void *inv_diag, double omega,
                         int n_iterations, double weight, void *stream);
/* one fused update x += omega * inv_diag * (b - t) for host layers that drive an exchange-aware vmult
 * themselves (t = A x on entry) */
int glsb_relaxation_update(glsb_op *op, void *x, const void *t, const void *b, const void *inv_diag, double omega,
                           void *stream);
/* PreconditionRelaxation::estimate_eigenvalues with EigenvalueAlgorithm::power_iteration and relaxation = 0:
 * start vector (i + first_local_index) % 11 minus its mean, zero on constrained dofs, normalised;
 * n_power_iterations of e <- D^-1 A e / |.|, lambda = e . D^-1 A e; ev_max = 1.2 lambda;
 * omega = 2 / (ev_max / smoothing_range + ev_max).  Synchronises the stream.  Single-rank operators. */
int glsb_estimate_relaxation(glsb_op *op, const void *inv_diag, int n_power_iterations, double smoothing_range,
                             double weight, uint64_t first_local_index, double *omega_out, double *ev_max_out,
                             void *stream);

/* ---- multigrid transfer on the device ---------------------------------
 * The reference moves vectors between the levels of its global-coarsening multigrid with deal.II's
 * MGTransferGlobalCoarsening over one MGTwoLevelTransfer per level pair (main.cc:540-563); the V-cycle
 * calls prolongate_and_add / restrict_and_add (multigrid.cc:534-548 -> deal.II Multigrid), and the level
 * operators get their linearization point and history through interpolate_to_mg (main.cc:789-790,
 * :825-827).  A glsb_transfer is one MGTwoLevelTransfer between a level and the next coarser one (same
 * FE_Q(degree)^(dim+1), every coarse cell refined once):
 *   prolongate_and_add : fine   += W P C_c coarse      (C_c resolves the coarse constraints, W = weights)
 *   restrict_and_add   : coarse += C_c^T P^T W fine
 *   interpolate        : coarse  = value of the fine function at the coarse support points (FE_Q restriction
 *                        matrices; used with a transfer built WITHOUT constraints, main.cc:540-549)
 * All vectors are device pointers of the transfer's number type. */
typedef struct glsb_transfer glsb_transfer;

typedef struct glsb_transfer_desc
{
  int32_t abi_version; /* GLSB_ABI_VERSION */
  int32_t device;
  int32_t dim, degree;
  int32_t number_type; /* GLSB_F32 for the level vectors of the reference (config.h:7) */
  uint64_t n_coarse_cells;
  uint64_t n_fine_dofs, n_coarse_dofs; /* local vector lengths */
  /* [n_coarse_cells][(dim+1)(degree+1)^dim] component-blocked lexicographic local indices into the coarse
   * vector, or GLSB_CONSTRAINED_BIT | row (constraints of the coarse level, main.cc:552-556) */
  const uint32_t *coarse_dof_indices;
  /* [n_coarse_cells][2^dim][(dim+1)(degree+1)^dim] plain indices into the fine vector; child number
   * cx + 2 cy (+ 4 cz) like GeometryInfo */
  const uint32_t *fine_dof_indices;
  uint32_t        n_constraint_rows; /* coarse constraint rows as CSR (may be 0) */
  const uint32_t *row_ptr;
  const uint32_t *entry_col;
  const double   *entry_val;
  /* [n_fine_dofs] deal.II's transfer weights: 1 / (number of fine cells touching the dof), 0 on constrained
   * fine dofs; NULL = all ones (discontinuous / tests) */
  const double *weights;
} glsb_transfer_desc;

int  glsb_transfer_create(const glsb_transfer_desc *desc, glsb_transfer **out);
void glsb_transfer_destroy(glsb_transfer *t);
const char *glsb_transfer_last_error(const glsb_transfer *t);
int glsb_transfer_prolongate_and_add(glsb_transfer *t, void *dst_fine, const void *src_coarse, void *stream);
int glsb_transfer_restrict_and_add(glsb_transfer *t, void *dst_coarse, const void *src_fine, void *stream);
int glsb_transfer_interpolate(glsb_transfer *t, void *dst_coarse, const void *src_fine, void *stream);

/* ---- device-resident Krylov vectors ----------------------------------
 * deal.II's SolverGMRES (solver_l.cc:46-74: restart 30, right preconditioning) orthogonalises every new
 * Krylov vector against up to 30 basis vectors with host loops over LinearAlgebra::distributed::Vector.
 * These are the same vector operations on device vectors, batched: one pass over the data for all k inner
 * products, one for the k-term update.  type = GLSB_F64 / GLSB_F32; results of reductions are double. */
/* out_dev[j] = sum_i V[j * stride + i] * w[i], j < k  (out_dev: k doubles on the device, overwritten;
 * deterministic two-stage reduction through one scratch buffer per device: calls on different streams of
 * the same device must not overlap) */
int glsb_vec_multi_dot(double *out_dev, const void *V, uint64_t stride, int k, const void *w, uint64_t n, int type,
                       void *stream);
/* w[i] += scale * sum_j coef_dev[j] * V[j * stride + i]  (coef_dev: k doubles on the device) */
int glsb_vec_multi_axpy(void *w, const void *V, uint64_t stride, int k, const double *coef_dev, double scale,
                        uint64_t n, int type, void *stream);
/* y = a x + b y  (b == 0: y is not read) */
int glsb_vec_axpby(void *y, double a, const void *x, double b, uint64_t n, int type, void *stream);
/* dst (dst_type) = src (src_type): the double <-> float copies of PreconditionMG::vmult / copy_to_mg /
 * copy_from_mg and MGCoarseGridApplyPreconditioner (multigrid.cc:6-149) */
int glsb_vec_convert(void *dst, int dst_type, const void *src, int src_type, uint64_t n, void *stream);
/* v[idx[i]] = 0: AffineConstraints::set_zero (main.cc:856) */
int glsb_vec_set_zero_indexed(void *v, const uint32_t *idx_dev, uint64_t n_idx, int type, void *stream);
/* y (type) = A x with a dense row-major double matrix on the device: the coarse-grid "direct" solver of
 * the V-cycle applied as the precomputed inverse (multigrid.cc:419-425, :512-529 use Trilinos on the host) */
int glsb_dense_apply(void *y, const double *A_dev, const void *x, uint32_t m, uint32_t n, int type, void *stream);

/* ---- ghost exchange helpers (update_ghost_values / compress(add)) ------- */

/* buf[i] = vec[export_indices[i]] */
int glsb_pack_export(glsb_op *op, void *buf, const void *vec, void *stream);
/* vec[export_indices[i]] += buf[i] */
int glsb_unpack_add(glsb_op *op, void *vec, const void *buf, void *stream);

/* ---- introspection (used by tests and the bench) ------------------------ */

uint64_t glsb_n_cells(const glsb_op *op);
uint64_t glsb_n_local(const glsb_op *op);     /* n_owned + n_ghost */
uint64_t glsb_n_interior_cells(const glsb_op *op);
/* copy a q-point table to a device buffer as [field][cell][q] in the caller's cell order;
 * name: "u_star_value", "u_star_gradient", "p_star_gradient", "u_time_derivative_old",
 * "delta_1", "delta_2", "delta_1_q", "delta_2_q" (operator_ns.h:114-132) */
int glsb_get_table(glsb_op *op, const char *name, void *out_dev, uint64_t out_count, void *stream);
/* kernel launches issued by this operator so far */
uint64_t glsb_launch_count(const glsb_op *op);
/* name of the kernel variant used for vmult ("generic" / "q2_regtile" ...) */
const char *glsb_vmult_variant(const glsb_op *op);
/* force a variant (testing): 0 = auto, 1 = generic */
int glsb_set_variant(glsb_op *op, int variant);

#ifdef __cplusplus
}
#endif
#endif /* GLSB200_H */
