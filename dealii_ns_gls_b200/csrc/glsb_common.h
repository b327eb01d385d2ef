// Internal declarations shared by the C-ABI layer and the kernel translation units.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace glsb
{
constexpr int MAX_N = 5; // degree 4

// 1-D tables, passed by value as a kernel parameter => they live in the constant bank
// and compile-time-indexed entries become immediate constant operands of DFMA/FFMA.
template <typename T, int n>
struct Shape
{
  T S[n * n];  // S[q*n + i]  = phi_i(x_q)          (FE_Q Lagrange basis at Gauss points)
  T D[n * n];  // D[q*n + q'] = collocation derivative on the Gauss points
  T G[n * n];  // G[q*n + i]  = phi_i'(x_q) (= D S)
  T w[n];      // 1-D Gauss weights on [0,1]
  // the same with the quadrature weights folded in (Cartesian cells: JxW = det J * w_qx w_qy w_qz is
  // separable, so the integration sweeps can carry the weights and the q-point loop needs none):
  T Sw[n * n]; // w_q S[q][i]
  T Gw[n * n]; // w_q G[q][i]
  T Dt[n * n]; // D[q'][q] w_q' / w_q
};

struct ShapeHost
{
  int    n;
  double S[MAX_N * MAX_N];
  double G[MAX_N * MAX_N];
  double D[MAX_N * MAX_N];
  double w[MAX_N];
  double xq[MAX_N];
  double nodes[MAX_N];
};
void compute_shape_host(int degree, ShapeHost &out);
// P[2][n][n], R[n][n], rchild[n]: two-level transfer matrices (glsb_basis.cpp)
void compute_transfer_host(int degree, double *P, double *R, int *rchild);
// Nf[2][n], Gf[2][n]: 1-D basis values / derivatives at xi = 0 and xi = 1 (boundary faces)
void compute_face_basis_host(int degree, double *Nf, double *Gf);

enum Branch : int
{
  BR_NEWTON      = 0, // do_vmult_cell<false>, increment_form (operator_ns.cc:1067-1182)
  BR_FIXED_POINT = 1, // do_vmult_cell<false>, !increment_form (operator_ns.cc:955-1066)
  BR_RESIDUAL    = 2  // do_vmult_cell<true> (same block, evaluate_residual terms on)
};

// Everything a cell kernel needs; T = Number. All pointers are device pointers.
template <typename T>
struct KParams
{
  // cells [cell_begin, cell_end) in internal order; ncp = padded cell stride of all SoA arrays
  uint32_t cell_begin, cell_end;
  uint32_t hole_begin, hole_end; // padding slots between the interior and the boundary range
  int      sm_reserve;           // multiprocessors to leave free (persistent kernels)
  int      packed;               // Q2 float kernel: two cells per lane (FFMA2)
  uint64_t ncp;
  const uint32_t *idx; // blocked [ncp/32][ndof + 1][32]: ndof = C*n_loc index rows (see idx_at()), then one
                       // row with a flag word per cell (non-zero: the cell has constrained dofs)
  uint32_t        ndof, nloc;
  const uint8_t  *cell_flags; // [ncp] bit 0: the cell has constrained dofs
  // constraint rows
  const uint32_t *row_dof, *row_ptr, *ecol;
  const T        *eval;
  // geometry
  int      geom;    // 0 Cartesian, 2 general
  const T *inv_jac; // Cartesian only: [dim][ncp] diagonal of J^-1
  const T *jxw;     // Cartesian only: [ncp] det J
  const double *h_min, *measure; // [ncp]
  // q-point data, one blocked array: element (field f, point q, cell) lives at
  //   ((((cell >> 5) * NL + q / QG) * FT + f) * QG + q % QG) * 32 + (cell & 31)
  // i.e. [batch of 32 cells][layer of QG points][field][point in layer][cell in batch], so that
  // the fields of one layer of one batch are one contiguous block (a single TMA bulk copy) and
  // 32 consecutive cells of a (field, point) row are contiguous (coalesced in every kernel).
  T  *Q;
  int FT, NL, QG;
  // field offsets (-1 = not stored): u_star_value, u_star_gradient, p_star_gradient,
  // u_time_derivative_old, delta_1_q, delta_2_q, J^-1 (general), JxW (general), u_old_gradient,
  // p_old_gradient (operator_ns.h:117-132)
  int fU, fH, fP, fO, fd1q, fd2q, fJ, fjxw, fGold, fgoldp;
  T  *d1c, *d2c; // [ncp] cell-wise delta_1, delta_2
  // scalars
  T      weight, nu, theta;
  double c1, c2, stau, nu_d;
  int    degree;
  int    ctd, cell_wise, has_o, theta_ne_1;
  int    sign_negative; // scatter -value (evaluate_residual "dst *= -1")
  // set_previous_solution
  const T *hist[4];
  T        hist_w[4];
  int      hist_n;
  // outputs / inputs
  const T *src;
  T       *dst;
  unsigned long long *max_bits; // get_max_u
  uint64_t unit_stride;         // get_system_matrix: elements between the columns of the output matrix
};

// dof index of local dof `dof` of cell `cell`: rows of 32 consecutive cells are contiguous (coalesced
// for every kernel) and the ndof + 1 rows of a 32-cell batch are one contiguous block (one bulk copy)
template <typename T>
__host__ __device__ inline uint64_t idx_at(const KParams<T> &p, uint32_t dof, uint32_t cell)
{
  // inside its row, component c's entry of a cell is rotated by 8c cells: the 4 component lanes of a
  // cell in the Q2 kernel then hit 4 different shared-memory banks when the block is staged there
  return ((uint64_t)(cell >> 5) * (p.ndof + 1) + dof) * 32 + (((cell & 31) + 8 * (dof / p.nloc)) & 31);
}

template <typename T>
__host__ __device__ inline bool cell_active(const KParams<T> &p, uint32_t cell)
{
  return cell < p.cell_end && !(cell >= p.hole_begin && cell < p.hole_end);
}

// position of (point q, cell) in the blocked q-point array.  Field f of (q, cell) lives at
//   base + f * fstride + ((lane32 + rot(f)) & 31),  lane32 = cell & 31,
// where rot(f) = 4 * (component row of the field) for u_star_value[c], u_star_gradient[c][.],
// p_star_gradient[c], u_time_derivative_old[c] and 0 otherwise: the Q2 kernel has the 4 component
// lanes of a cell read rows c = 0..2 of these fields in one request, and the rotation puts them
// in different shared-memory banks.
template <typename T>
struct QPos
{
  uint64_t base;
  uint32_t fstride, lane32;
};
template <typename T>
__host__ __device__ inline QPos<T> qpos(const KParams<T> &p, uint32_t q, uint32_t cell)
{
  QPos<T> r;
  const uint32_t layer = q / (uint32_t)p.QG, ql = q - layer * (uint32_t)p.QG;
  r.base    = ((((uint64_t)(cell >> 5) * p.NL + layer) * p.FT) * p.QG + ql) * 32;
  r.fstride = (uint32_t)p.QG * 32;
  r.lane32  = cell & 31;
  return r;
}
// offset of field f with component row `row` (0 for fields without one)
template <typename T>
__host__ __device__ inline uint64_t qoff(const QPos<T> &qp, int f, int row)
{
  return qp.base + (uint64_t)f * qp.fstride + ((qp.lane32 + 4u * (uint32_t)row) & 31u);
}

// columns of C_cell for cells with weighted constraint rows (compute_diagonal)
struct DiagColumns
{
  const uint32_t *cell;    // [n_list] internal cell index
  const uint32_t *col_ptr; // [n_list + 1] -> columns
  const uint32_t *col_dof; // [n_cols] vector index g
  const uint32_t *ent_ptr; // [n_cols + 1] -> entries
  const uint32_t *ent_loc; // local dof index (c*n_loc + l)
  const double   *ent_val; // weight
  uint32_t        n_list;
};

// dispatch entry points implemented per (dim, T) translation unit
template <int dim, typename T>
struct Kernels
{
  static int vmult(int n, int branch, const KParams<T> &p, const ShapeHost &sh, cudaStream_t s);
  static int linearization(int n, const KParams<T> &p, const ShapeHost &sh, cudaStream_t s);
  static int previous(int n, int with_gradients, const KParams<T> &p, const ShapeHost &sh, cudaStream_t s);
  static int diagonal(int n, int branch, const KParams<T> &p, const ShapeHost &sh, const uint8_t *skip_cell,
                      const DiagColumns &dc, cudaStream_t s);
  static int max_u(int n, const KParams<T> &p, const ShapeHost &sh, cudaStream_t s);
  // all columns of the system matrix in one launch (unit vectors as src, column-major T matrix at p.dst)
  static int matrix_columns(int n, int branch, uint32_t n_columns, const KParams<T> &p, const ShapeHost &sh,
                            cudaStream_t s);
  // register-tiled Q2 (dim 3, degree 2) Newton-branch vmult; returns -1 if not applicable
  static int vmult_q2(const KParams<T> &p, const ShapeHost &sh, int n_stage_fields, cudaStream_t s);
};

template <typename T, int n>
inline Shape<T, n> to_shape(const ShapeHost &h)
{
  Shape<T, n> s;
  for (int i = 0; i < n * n; ++i)
    {
      s.S[i] = (T)h.S[i];
      s.D[i] = (T)h.D[i];
      s.G[i] = (T)h.G[i];
    }
  for (int i = 0; i < n; ++i)
    s.w[i] = (T)h.w[i];
  for (int q = 0; q < n; ++q)
    for (int j = 0; j < n; ++j)
      {
        s.Sw[q * n + j] = (T)(h.w[q] * h.S[q * n + j]);
        s.Gw[q * n + j] = (T)(h.w[q] * h.G[q * n + j]);
        s.Dt[q * n + j] = (T)(h.D[q * n + j] * h.w[q] / h.w[j]);
      }
  return s;
}

// register-tiled vmult kernel (glsb_q2.cuh), one translation unit per (Number, n): glsb_inst_q2.cu.
// Return -1 if the variant does not apply (the caller falls back to the generic / column kernel).
namespace q2
{
template <typename T, int n>
int launch_degree(const KParams<T> &p, const ShapeHost &sh, int F, cudaStream_t s);
int launch_packed_q2(const KParams<float> &p, const ShapeHost &sh, int F, cudaStream_t s); // float, FFMA2
inline int launch_packed_q2(const KParams<double> &, const ShapeHost &, int, cudaStream_t) { return -1; }
// Q3 / Q4: 2 n^3 values per lane fit in float only
template <int n>
inline int launch_float_only(const KParams<float> &p, const ShapeHost &sh, int F, cudaStream_t s)
{
  return launch_degree<float, n>(p, sh, F, s);
}
template <int n>
inline int launch_float_only(const KParams<double> &, const ShapeHost &, int, cudaStream_t)
{
  return -1;
}
} // namespace q2

} // namespace glsb
