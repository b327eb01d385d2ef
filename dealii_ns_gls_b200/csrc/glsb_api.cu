// C ABI of libglsb200.so (include/glsb200.h): operator state, setup-time layout
// transformations and dispatch to the kernels.  No CPU fallback anywhere.
#include "../../include/glsb200.h"
#include "glsb_common.h"
#include "glsb_faces.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace glsb;

namespace
{
std::string g_create_error;

struct DevBuf
{
  void  *p     = nullptr;
  size_t bytes = 0;
  DevBuf()     = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release()
  {
    if (p)
      cudaFree(p);
    p     = nullptr;
    bytes = 0;
  }
  bool alloc(size_t b)
  {
    release();
    if (b == 0)
      return true;
    if (cudaMalloc(&p, b) != cudaSuccess)
      {
        p = nullptr;
        return false;
      }
    bytes = b;
    return true;
  }
  template <typename U>
  U *as() const
  {
    return static_cast<U *>(p);
  }
};

__global__ void k_transpose_idx(const uint32_t *__restrict__ raw, const uint32_t *__restrict__ perm,
                                const uint8_t *__restrict__ flags, uint32_t *__restrict__ out, uint32_t n_cells,
                                uint32_t ndof, uint32_t nloc, uint64_t ncp)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint64_t)(ndof + 1) * ncp)
    return;
  const uint32_t i = t % ncp, d = t / ncp;
  // blocked [ncp/32][ndof + 1][32]; perm is defined for every slot (padding repeats a cell)
  if (d == ndof)
    out[((uint64_t)(i >> 5) * (ndof + 1) + d) * 32 + (i & 31)] = flags[i];
  else
    out[((uint64_t)(i >> 5) * (ndof + 1) + d) * 32 + (((i & 31) + 8 * (d / nloc)) & 31)] =
      raw[(uint64_t)perm[i] * ndof + d];
}

// general geometry raw[(cell*nq + q)*inner + f] (double, caller's cell order) -> blocked q-point
// array field f0 + f (see KParams::Q)
template <typename T>
__global__ void k_fill_geom(const double *__restrict__ raw, const uint32_t *__restrict__ perm, T *__restrict__ Q,
                            uint32_t nq, uint32_t inner, uint64_t ncp, int FT, int NL, int QG, int f0)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint64_t)inner * nq * ncp)
    return;
  const uint32_t i = t % ncp;
  const uint64_t r = t / ncp;
  const uint32_t q = r % nq, f = r / nq;
  const uint32_t layer = q / QG, ql = q - layer * QG;
  const uint64_t o = ((((uint64_t)(i >> 5) * NL + layer) * FT + (f0 + f)) * QG + ql) * 32 + (i & 31);
  Q[o] = (T)raw[((uint64_t)perm[i] * nq + q) * inner + f];
}

template <typename T>
__global__ void k_copy_indexed(T *__restrict__ dst, const T *__restrict__ src, const uint32_t *__restrict__ idx,
                               uint32_t n)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n)
    dst[idx[t]] = src[idx[t]];
}

template <typename T>
__global__ void k_pack(T *__restrict__ buf, const T *__restrict__ vec, const uint32_t *__restrict__ idx, uint64_t n)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n)
    buf[t] = vec[idx[t]];
}

template <typename T>
__global__ void k_unpack_add(T *__restrict__ vec, const T *__restrict__ buf, const uint32_t *__restrict__ idx,
                             uint64_t n)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n)
    atomicAdd(vec + idx[t], buf[t]);
}

// ---- relaxation smoother (PreconditionRelaxation with a DiagonalMatrix, multigrid.h:67-69) ----
template <typename T>
__global__ void k_relax_first(T *__restrict__ x, const T *__restrict__ b, const T *__restrict__ d, T omega, uint64_t n)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    x[i] = omega * (d[i] * b[i]);
}
template <typename T>
__global__ void k_relax_update(T *__restrict__ x, const T *__restrict__ t, const T *__restrict__ b,
                               const T *__restrict__ d, T omega, uint64_t n)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    x[i] += omega * (d[i] * (b[i] - t[i]));
}
// power iteration pieces; sums[0..2] are double accumulators on the device
// block-level sum of v, one atomicAdd per block (one per WARP serialises ~1e6 atomics on one address at 3e7
// dofs: 4.6 ms per power iteration against 0.9 ms for the vmult)
__device__ __forceinline__ void block_atomic_sum(double v, double *__restrict__ target)
{
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads(); // red may still be read by the previous call
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32)
    {
      v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, o);
      if (threadIdx.x == 0 && v != 0)
        atomicAdd(target, v);
    }
}
template <typename T>
__global__ void k_pi_init(T *__restrict__ e, uint64_t first, uint64_t n, double *__restrict__ sums)
{
  double v = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    {
      const double x = (double)((i + first) % 11);
      e[i]           = (T)x;
      v += x;
    }
  block_atomic_sum(v, sums);
}
template <typename T>
__global__ void k_pi_shift(T *__restrict__ e, const double *__restrict__ sums, uint64_t n)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    e[i] -= (T)(sums[0] / (double)n); // vector.add(-mean_value)
}
// v1 = d * v2 (v2 may be null: v1 = e, for the initial normalisation); sums[1] += e . v1, sums[2] += v1 . v1
template <typename T>
__global__ void k_pi_apply(T *__restrict__ v1, const T *__restrict__ v2, const T *__restrict__ d,
                           const T *__restrict__ e, uint64_t n, double *__restrict__ sums)
{
  double s1 = 0, s2 = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    {
      const T x = v2 ? d[i] * v2[i] : e[i];
      if (v2)
        v1[i] = x;
      s1 += (double)e[i] * (double)x;
      s2 += (double)x * (double)x;
    }
  block_atomic_sum(s1, sums + 1);
  block_atomic_sum(s2, sums + 2);
}
// e = v1 / |v1|; records lambda = sums[1] into hist[k] and clears the accumulators
template <typename T>
__global__ void k_pi_normalise(T *__restrict__ e, const T *__restrict__ v1, uint64_t n,
                               const double *__restrict__ sums)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    e[i] = (T)((double)v1[i] / sqrt(sums[2]));
}
__global__ void k_pi_record(double *__restrict__ sums, double *__restrict__ hist, int k)
{
  if (k >= 0)
    hist[k] = fabs(sums[1]); // deal.II's power_iteration returns std::abs(eigenvalue_estimate)
  sums[0] = sums[1] = sums[2] = 0;
}

// edge-constrained dofs (GMG-LS): save src at the edge indices and zero it there / restore
template <typename T>
__global__ void k_edge_save_zero(T *__restrict__ src, T *__restrict__ saved, const uint32_t *__restrict__ idx, uint32_t n)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n)
    {
      saved[t]    = src[idx[t]];
      src[idx[t]] = T(0);
    }
}
template <typename T>
__global__ void k_edge_restore(T *__restrict__ dst, T *__restrict__ src, const T *__restrict__ saved,
                               const uint32_t *__restrict__ idx, uint32_t n)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n)
    {
      src[idx[t]] = saved[t];
      dst[idx[t]] = saved[t]; // the reference writes the saved SRC value into dst (operator_ns.cc:729-730)
    }
}

// glsb_vmult_host: last[d] = last chunk of cells that touches vector entry d
__global__ void k_last_touch(const uint32_t *__restrict__ idx, const uint8_t *__restrict__ chunk_of_batch,
                             const uint32_t *__restrict__ row_dof, const uint32_t *__restrict__ row_ptr,
                             const uint32_t *__restrict__ ecol, uint32_t *__restrict__ last, uint32_t n_slots,
                             uint32_t ndof)
{
  // n_slots = whole 32-slot batches.  A row of a batch is rotated by 8 * component, so the entries of a real cell
  // can sit in the columns of the batch's padding slots: every column of every batch is visited (padding slots
  // repeat a real cell of the same batch, i.e. of the same chunk, and mark nothing new)
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint64_t)ndof * n_slots)
    return;
  const uint32_t i = t % n_slots, j = t / n_slots;
  const uint32_t iv = idx[((uint64_t)(i >> 5) * (ndof + 1) + j) * 32 + (i & 31)];
  const uint32_t c  = chunk_of_batch[i >> 5];
  if (iv & GLSB_CONSTRAINED_BIT)
    {
      const uint32_t r = iv & ~GLSB_CONSTRAINED_BIT;
      atomicMax(last + row_dof[r], c);
      for (uint32_t e = row_ptr[r]; e < row_ptr[r + 1]; ++e)
        atomicMax(last + ecol[e], c);
    }
  else
    atomicMax(last + iv, c);
}
__global__ void k_mark_last(const uint32_t *__restrict__ idx, uint64_t n, uint32_t value, uint32_t *__restrict__ last)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n)
    atomicMax(last + idx[t], value);
}
// flag[d] = entry d is touched again after the chunk whose completion triggers its download
__global__ void k_late_flags(const uint32_t *__restrict__ last, const uint64_t *__restrict__ in_end, uint32_t nch,
                             uint8_t *__restrict__ flag, uint64_t n)
{
  const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n)
    return;
  uint32_t c = 0;
  while (c + 1 < nch && in_end[c] <= d)
    ++c;
  flag[d] = last[d] > c;
}
// compact the flags into an index list (order irrelevant); count may exceed capacity: the caller then keeps the scan
__global__ void k_compact_flags(const uint8_t *__restrict__ flag, uint64_t n, uint32_t *__restrict__ list,
                                uint32_t capacity, unsigned long long *__restrict__ count)
{
  const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d < n && flag[d])
    {
      const unsigned long long k = atomicAdd(count, 1ull);
      if (k < capacity)
        list[k] = (uint32_t)d;
    }
}
template <typename T>
__global__ void k_flush_listed(T *__restrict__ host_dst, const T *__restrict__ dev_dst, const uint32_t *__restrict__ list,
                               uint32_t m)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < m)
    host_dst[list[t]] = dev_dst[list[t]];
}
// re-send the flagged entries with stores to the (mapped, page-locked) host vector
template <typename T>
__global__ void k_flush_flagged(T *__restrict__ host_dst, const T *__restrict__ dev_dst,
                                const uint8_t *__restrict__ flag, uint64_t n)
{
  const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d < n && flag[d])
    host_dst[d] = dev_dst[d];
}

// diag finish: constrained rows -> 1, then x -> |x| > 1e-10 ? 1/x : 1 (operator_ns.cc:220-224)
template <typename T>
__global__ void k_set_indexed(T *__restrict__ v, const uint32_t *__restrict__ idx, uint32_t n, T val)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n)
    v[idx[t]] = val;
}
template <typename T>
__global__ void k_invert_guarded(T *__restrict__ v, uint64_t n)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n)
    {
      const T x = v[t];
      v[t]      = (fabs(x) > T(1.0e-10)) ? (T(1.0) / x) : T(1.0);
    }
}

// blocked q-point fields f0 .. f0+nf-1 (internal cell order) -> out [f][cell][q] (caller's order);
// per-cell arrays: nq = 1 and Q = the plain [ncp] array (FT = NL = QG = 1 addressing degenerates)
template <typename T>
__global__ void k_export_table(const T *__restrict__ Q, const uint32_t *__restrict__ perm, T *__restrict__ out,
                               uint32_t n_cells, uint32_t n_slots, uint32_t hole_begin, uint32_t hole_end,
                               uint32_t nq, uint32_t nf, int FT, int NL, int QG, int f0, int per_cell, int rowdiv)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint64_t)nf * nq * n_slots)
    return;
  const uint32_t i = t % n_slots;
  if (i >= hole_begin && i < hole_end)
    return;
  const uint64_t r = t / n_slots;
  const uint32_t q = r % nq, f = r / nq;
  uint64_t       o;
  if (per_cell)
    o = i;
  else
    {
      const uint32_t layer = q / QG, ql = q - layer * QG;
      const uint32_t row = rowdiv > 0 ? f / rowdiv : 0; // component row of the field (rotation, see qoff())
      o = ((((uint64_t)(i >> 5) * NL + layer) * FT + (f0 + f)) * QG + ql) * 32 + (((i & 31) + 4 * row) & 31);
    }
  out[((uint64_t)f * n_cells + perm[i]) * nq + q] = Q[o];
}
} // namespace

struct glsb_op
{
  int      dim = 0, degree = 0, n = 0, number_type = 0, C = 0, n_loc = 0, nq = 0;
  int      increment_form = 0, ctd = 0, cell_wise = 0, time_order = 0, geom = 0, device = 0;
  double   nu = 0, c1 = 0, c2 = 0, theta = 1;
  uint64_t n_cells = 0, n_owned = 0, n_ghost = 0, ncp = 0;
  uint32_t n_interior = 0, n_int_pad = 0, n_slots = 0;
  // deterministic mode (GLSB_DETERMINISTIC=1, single-rank operators): cells sorted by colour, color_begin[c] =
  // first slot of colour c; cells of one colour share no vector entry, the cell loops run colour by colour
  std::vector<uint32_t> color_begin;
  uint32_t              hole_override[2] = {0, 0};
  uint32_t n_rows = 0, n_constrained = 0;
  uint64_t n_export = 0;
  size_t   tsize = 8;

  DevBuf   relax_t, pi_e, pi_v1, pi_v2, pi_sums; // smoother scratch vectors
  DevBuf   mat_e, mat_col;                        // glsb_get_system_matrix scratch vectors
  // boundary faces with outflow terms (operator_ns.cc:1195-1301)
  uint32_t n_faces = 0;
  DevBuf   f_slot, f_no, f_kind, f_normal, f_jxw, f_inv_jac, f_target, f_beta, f_velocity;
  double   Nf[2 * MAX_N] = {}, Gf[2 * MAX_N] = {};
  uint32_t n_edge = 0;
  int      has_edge = 0;
  DevBuf   edge_idx, edge_saved, edge_cpy;
  DevBuf perm, idx, cell_flags, row_dof, row_ptr, ecol, eval, cidx, inv_jac, jxw, h_min, measure, export_idx;
  DevBuf Q, d1c, d2c, max_bits;
  int    FT = 0, NL = 1, QG = 1, F_stage = 0;
  int    packed = 0; // float Q2 operators: two cells per lane in the vmult kernel
  bool   regtile = false; // a register-tiled kernel exists for this (dim, degree, number)
  int    fU = -1, fH = -1, fP = -1, fO = -1, fd1q = -1, fd2q = -1, fJ = -1, fjxw = -1, fGold = -1, fgoldp = -1;
  DevBuf diag_skip, dc_cell, dc_col_ptr, dc_col_dof, dc_ent_ptr, dc_ent_loc, dc_ent_val;
  uint32_t dc_n_list = 0;

  // vmult on host vectors (glsb_vmult_host): cell chunks with the dof ranges they need / complete
  struct HostPipe
  {
    bool                     ready = false;
    cudaStream_t             s_in = nullptr, s_out = nullptr;
    cudaEvent_t              e_start = nullptr, e_done = nullptr;
    std::vector<cudaEvent_t> e_in, e_cells;
    std::vector<uint32_t>    cell_begin; // [n_chunks + 1] slots (multiples of 32)
    std::vector<uint64_t>    in_end;     // [n_chunks] src[0, in_end[c]) is needed by chunks 0..c
    std::vector<uint64_t>    out_end;    // [n_chunks] dst[0, out_end[c]) is final after chunks 0..c
    std::vector<uint32_t>    cidx_end;   // [n_chunks] constrained indices (sorted) below out_end[c]
    std::vector<uint32_t>    cidx_end_spec; // ... below in_end[c] (speculative download)
    DevBuf                   src, dst, late_flag; // late_flag[d]: entry d changes after its speculative download
    DevBuf                   late_list;           // the same as a compact index list (n_late entries), if it fits
    uint64_t                 n_late = 0;
  } hp, hpp; // hpp: the pipeline of partitioned operators (interior cells only, glsb_vmult_host_begin / _finish)
  std::vector<uint32_t> cidx_sorted;
  std::vector<uint32_t> slot_lo, slot_hi; // per 32-slot batch: smallest / largest vector index touched

  ShapeHost shape;
  bool      lin_valid = false, prev_valid = false;
  double    lin_dt = 0;
  int       variant_forced = 0, sm_reserve = 0;
  uint32_t  range_override[2] = {0, 0}; // glsb_vmult_host: explicit slot range for the next cell launch
  uint64_t  launches = 0;
  std::string err;
  std::string variant = "generic";
};

namespace
{
int fail(glsb_op *op, const std::string &msg)
{
  if (op)
    op->err = msg;
  else
    g_create_error = msg;
  return 1;
}

int cuda_fail(glsb_op *op, const char *what)
{
  cudaError_t e = cudaGetLastError();
  char        b[512];
  snprintf(b, sizeof b, "%s: %s", what, cudaGetErrorString(e));
  return fail(op, b);
}

template <typename T>
KParams<T> base_params(const glsb_op *op)
{
  KParams<T> p;
  memset(&p, 0, sizeof p);
  p.ncp        = op->ncp;
  p.idx        = op->idx.as<uint32_t>();
  p.ndof       = (uint32_t)(op->C * op->n_loc);
  p.nloc       = (uint32_t)op->n_loc;
  p.cell_flags = op->cell_flags.as<uint8_t>();
  p.row_dof    = op->row_dof.as<uint32_t>();
  p.row_ptr    = op->row_ptr.as<uint32_t>();
  p.ecol       = op->ecol.as<uint32_t>();
  p.eval       = op->eval.as<T>();
  p.geom       = op->geom;
  p.inv_jac    = op->inv_jac.as<T>();
  p.jxw        = op->jxw.as<T>();
  p.h_min      = op->h_min.as<double>();
  p.measure    = op->measure.as<double>();
  p.Q          = op->Q.as<T>();
  p.FT         = op->FT;
  p.NL         = op->NL;
  p.QG         = op->QG;
  p.fU         = op->fU;
  p.fH         = op->fH;
  p.fP         = op->fP;
  p.fO         = op->fO;
  p.fd1q       = op->fd1q;
  p.fd2q       = op->fd2q;
  p.fJ         = op->fJ;
  p.fjxw       = op->fjxw;
  p.fGold      = op->fGold;
  p.fgoldp     = op->fgoldp;
  p.d1c        = op->d1c.as<T>();
  p.d2c        = op->d2c.as<T>();
  p.nu         = (T)op->nu;
  p.theta      = (T)op->theta;
  p.c1         = op->c1;
  p.c2         = op->c2;
  p.nu_d       = op->nu;
  p.degree     = op->degree;
  p.ctd        = op->ctd;
  p.cell_wise  = op->cell_wise;
  p.has_o      = op->prev_valid && op->fO >= 0;
  p.theta_ne_1 = (op->theta != 1.0);
  p.max_bits   = op->max_bits.as<unsigned long long>();
  p.sm_reserve = op->sm_reserve;
  p.packed     = op->packed;
  return p;
}

template <typename T>
void cell_range(const glsb_op *op, int which, KParams<T> &p)
{
  p.cell_begin = 0;
  p.cell_end   = op->n_slots;
  p.hole_begin = op->n_interior;
  p.hole_end   = op->n_int_pad;
  if (which == GLSB_CELLS_INTERIOR)
    {
      p.cell_end   = op->n_interior;
      p.hole_begin = p.hole_end = 0;
    }
  else if (which == GLSB_CELLS_BOUNDARY)
    {
      p.cell_begin = op->n_int_pad;
      p.hole_begin = p.hole_end = 0;
    }
  if (op->range_override[1] > op->range_override[0])
    {
      p.cell_begin = op->range_override[0];
      p.cell_end   = op->range_override[1];
    }
  if (op->hole_override[1] > op->hole_override[0])
    {
      p.hole_begin = op->hole_override[0];
      p.hole_end   = op->hole_override[1];
    }
}

// deterministic mode: run `launch` once per colour.  A colour's slots [b, e) need not start on a 32-slot batch:
// the launch starts at the batch boundary below b and the slots in between (the previous colour's) are a hole.
template <typename F>
int for_each_color(glsb_op *op, F &&launch)
{
  const bool ranged = op->range_override[1] > op->range_override[0];
  if (op->color_begin.empty() || ranged)
    return launch();
  int rc = 0;
  for (size_t c = 0; c + 1 < op->color_begin.size() && rc == 0; ++c)
    {
      const uint32_t b = op->color_begin[c], e = op->color_begin[c + 1];
      if (e <= b)
        continue;
      op->range_override[0] = b & ~31u;
      op->range_override[1] = e;
      op->hole_override[0]  = b & ~31u;
      op->hole_override[1]  = b;
      rc                    = launch();
    }
  op->range_override[0] = op->range_override[1] = 0;
  op->hole_override[0] = op->hole_override[1] = 0;
  return rc;
}

#define GLSB_DISPATCH(op, CALL)                                   \
  do                                                              \
    {                                                             \
      if ((op)->number_type == GLSB_F64)                          \
        {                                                         \
          if ((op)->dim == 2)                                     \
            rc = CALL(2, double);                                 \
          else                                                    \
            rc = CALL(3, double);                                 \
        }                                                         \
      else                                                        \
        {                                                         \
          if ((op)->dim == 2)                                     \
            rc = CALL(2, float);                                  \
          else                                                    \
            rc = CALL(3, float);                                  \
        }                                                         \
    }                                                             \
  while (0)

template <typename T>
FaceParams<T> face_params(const glsb_op *op)
{
  FaceParams<T> f;
  memset(&f, 0, sizeof f);
  f.nf       = op->n_faces;
  f.n        = op->n;
  f.nloc     = op->n_loc;
  f.dim      = op->dim;
  f.nqf      = op->dim == 2 ? op->n : op->n * op->n;
  f.slot     = op->f_slot.as<uint32_t>();
  f.no       = op->f_no.as<uint32_t>();
  f.kind     = op->f_kind.as<uint32_t>();
  f.normal   = op->f_normal.as<T>();
  f.jxw      = op->f_jxw.as<T>();
  f.inv_jac  = op->f_inv_jac.as<T>();
  f.target   = op->f_target.as<T>();
  f.beta     = op->f_beta.as<T>();
  f.velocity = op->f_velocity.as<T>();
  f.nu       = (T)op->nu;
  for (int i = 0; i < op->n * op->n; ++i)
    {
      f.S[i] = (T)op->shape.S[i];
      f.G[i] = (T)op->shape.G[i];
    }
  for (int i = 0; i < 2 * op->n; ++i)
    {
      f.Nf[i] = (T)op->Nf[i];
      f.Gf[i] = (T)op->Gf[i];
    }
  return f;
}

// the boundary lambda of MatrixFree::loop (operator_ns.cc:710-717): which = 0 apply, 1 residual, 2 face_velocity,
// 3 diagonal
template <int dim, typename T>
int do_faces(glsb_op *op, const KParams<T> &p, int what, cudaStream_t s)
{
  if (op->n_faces == 0)
    return 0;
  const FaceParams<T> f  = face_params<T>(op);
  const size_t        sm = sizeof(T) * face_smem_elems<dim>(op->n);
  op->launches++;
  if (what == 0)
    k_faces_apply<T, dim, false><<<op->n_faces, FACE_THREADS, sm, s>>>(p, f);
  else if (what == 1)
    k_faces_apply<T, dim, true><<<op->n_faces, FACE_THREADS, sm, s>>>(p, f);
  else if (what == 2)
    k_faces_velocity<T, dim><<<op->n_faces, FACE_THREADS, sm, s>>>(p, f);
  else
    k_faces_diag<T, dim><<<op->n_faces, FACE_THREADS, 0, s>>>(p, f);
  return cudaGetLastError() != cudaSuccess;
}

template <int dim, typename T>
int do_cells(glsb_op *op, void *dst, const void *src, double weight, int which, int branch, cudaStream_t s,
             int part = 0, int n_parts = 1)
{
  // boundary faces ride with the last part of the launch that covers the cells next to the partition surface
  // (their contributions to ghost dofs must be in dst before compress(add))
  const bool colored    = !op->color_begin.empty() && op->hole_override[1] >= op->hole_override[0] &&
                       op->range_override[1] > op->range_override[0] && op->range_override[1] == op->color_begin.back();
  const bool with_faces = op->n_faces > 0 && which != GLSB_CELLS_INTERIOR && part == n_parts - 1 &&
                          (!(op->range_override[1] > op->range_override[0]) || colored);
  if (with_faces)
    {
      KParams<T> pf = base_params<T>(op);
      cell_range(op, GLSB_CELLS_ALL, pf);
      pf.src           = static_cast<const T *>(src);
      pf.dst           = static_cast<T *>(dst);
      pf.weight        = (T)weight;
      pf.sign_negative = (branch == BR_RESIDUAL);
      if (do_faces<dim, T>(op, pf, branch == BR_RESIDUAL ? 1 : 0, s))
        return 1;
    }
  if (!op->color_begin.empty() && which == GLSB_CELLS_ALL && n_parts == 1 &&
      !(op->range_override[1] > op->range_override[0]))
    return for_each_color(op, [&]() { return do_cells<dim, T>(op, dst, src, weight, which, branch, s, 0, 1); });
  KParams<T> p = base_params<T>(op);
  cell_range(op, which, p);
  if (n_parts > 1)
    {
      // equal chunks of whole 32-cell batches
      const uint32_t nb = (p.cell_end - p.cell_begin + 31) / 32;
      const uint32_t b0 = (uint32_t)((uint64_t)nb * part / n_parts), b1 = (uint32_t)((uint64_t)nb * (part + 1) / n_parts);
      const uint32_t base = p.cell_begin, end = p.cell_end;
      p.cell_begin = base + b0 * 32;
      p.cell_end   = (base + b1 * 32 < end) ? base + b1 * 32 : end;
    }
  p.src           = static_cast<const T *>(src);
  p.dst           = static_cast<T *>(dst);
  p.weight        = (T)weight;
  p.sign_negative = (branch == BR_RESIDUAL);
  if (p.cell_end <= p.cell_begin)
    return 0;
  op->launches++;
  if (dim == 3 && op->regtile && branch == BR_NEWTON && op->variant_forced != 1)
    {
      const int rc = Kernels<dim, T>::vmult_q2(p, op->shape, op->F_stage, s);
      if (rc >= 0)
        {
          static const char *names[] = {"", "", "q1_regtile_tma", "q2_regtile_tma", "q3_regtile_tma", "q4_regtile_tma"};
          op->variant = names[op->n];
          return rc;
        }
    }
  op->variant = (dim == 3 && op->n >= 4 && getenv("GLSB_NO_COL") == nullptr) ? "column" : "generic";
  return Kernels<dim, T>::vmult(op->n, branch, p, op->shape, s);
}

template <int dim, typename T>
int do_lin(glsb_op *op, const void *vec, double dt, cudaStream_t s)
{
  KParams<T> p = base_params<T>(op);
  cell_range(op, GLSB_CELLS_ALL, p);
  p.src  = static_cast<const T *>(vec);
  p.stau = (dt == 0.0) ? 0.0 : 1.0 / dt;
  op->launches++;
  if (do_faces<dim, T>(op, p, 2, s)) // face_velocity (operator_ns.cc:460-476)
    return 1;
  return Kernels<dim, T>::linearization(op->n, p, op->shape, s);
}

template <int dim, typename T>
int do_prev(glsb_op *op, const void *const *history, const double *weights, int order, cudaStream_t s)
{
  KParams<T> p = base_params<T>(op);
  cell_range(op, GLSB_CELLS_ALL, p);
  p.hist_n = order;
  for (int i = 0; i < order; ++i)
    {
      p.hist[i]   = static_cast<const T *>(history[i + 1]);
      p.hist_w[i] = (T)weights[i + 1];
    }
  op->launches++;
  int rc = Kernels<dim, T>::previous(op->n, 0, p, op->shape, s);
  if (rc == 0 && op->theta != 1.0)
    {
      p.hist_n    = 1;
      p.hist[0]   = static_cast<const T *>(history[1]);
      p.hist_w[0] = (T)1;
      op->launches++;
      rc = Kernels<dim, T>::previous(op->n, 1, p, op->shape, s);
    }
  return rc;
}

template <int dim, typename T>
int do_diag(glsb_op *op, void *diag, double weight, cudaStream_t s)
{
  KParams<T> p = base_params<T>(op);
  cell_range(op, GLSB_CELLS_ALL, p);
  p.dst    = static_cast<T *>(diag);
  p.weight = (T)weight;
  DiagColumns dc;
  dc.cell    = op->dc_cell.as<uint32_t>();
  dc.col_ptr = op->dc_col_ptr.as<uint32_t>();
  dc.col_dof = op->dc_col_dof.as<uint32_t>();
  dc.ent_ptr = op->dc_ent_ptr.as<uint32_t>();
  dc.ent_loc = op->dc_ent_loc.as<uint32_t>();
  dc.ent_val = op->dc_ent_val.as<double>();
  dc.n_list  = op->dc_n_list;
  op->launches += 1 + (dc.n_list > 0);
  if (do_faces<dim, T>(op, p, 3, s))
    return 1;
  if (!op->color_begin.empty() && !(op->range_override[1] > op->range_override[0]))
    {
      // colour by colour; the cells with weighted rows (dc list) are handled by the last launch only
      const size_t nc = op->color_begin.size() - 1;
      size_t       c  = 0;
      return for_each_color(op, [&]() {
        KParams<T> pc = p;
        cell_range(op, GLSB_CELLS_ALL, pc);
        DiagColumns dcc = dc;
        if (++c < nc)
          dcc.n_list = 0;
        return Kernels<dim, T>::diagonal(op->n, op->increment_form ? BR_NEWTON : BR_FIXED_POINT, pc, op->shape,
                                         op->diag_skip.as<uint8_t>(), dcc, s);
      });
    }
  return Kernels<dim, T>::diagonal(op->n, op->increment_form ? BR_NEWTON : BR_FIXED_POINT, p, op->shape,
                                   op->diag_skip.as<uint8_t>(), dc, s);
}

template <int dim, typename T>
int do_maxu(glsb_op *op, const void *vec, cudaStream_t s)
{
  KParams<T> p = base_params<T>(op);
  cell_range(op, GLSB_CELLS_ALL, p);
  p.src = static_cast<const T *>(vec);
  op->launches++;
  return Kernels<dim, T>::max_u(op->n, p, op->shape, s);
}

bool upload(DevBuf &b, const void *host, size_t bytes)
{
  if (!b.alloc(bytes))
    return false;
  if (bytes)
    return cudaMemcpy(b.p, host, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
  return true;
}

template <typename T>
bool upload_converted(DevBuf &b, const double *host, size_t count)
{
  std::vector<T> tmp(count);
  for (size_t i = 0; i < count; ++i)
    tmp[i] = (T)host[i];
  return upload(b, tmp.data(), count * sizeof(T));
}

bool ensure_tables(glsb_op *, bool, bool) { return true; } // the q-point array is allocated in glsb_create
} // namespace

namespace
{
template <typename T>
__global__ void k_unit_vector(T *__restrict__ e, uint64_t n, uint64_t j)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    e[i] = (i == j) ? T(1) : T(0);
}
// A (row-major double) = transpose of the column-major T matrix M; 1 on the diagonal of constrained rows
template <typename T>
__global__ void k_matrix_finish(double *__restrict__ A, const T *__restrict__ M, uint64_t n)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n * n)
    {
      const uint64_t i = t / n, j = t - i * n;
      A[t]             = (double)M[j * n + i];
    }
}
__global__ void k_matrix_identity_rows(double *__restrict__ A, const uint32_t *__restrict__ idx, uint32_t m, uint64_t n)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < m)
    A[(uint64_t)idx[t] * n + idx[t]] = 1.0;
}

template <int dim, typename T>
int do_matrix(glsb_op *op, void *M, double weight, cudaStream_t s)
{
  KParams<T> p = base_params<T>(op);
  cell_range(op, GLSB_CELLS_ALL, p);
  p.dst         = static_cast<T *>(M);
  p.weight      = (T)weight;
  p.unit_stride = op->n_owned + op->n_ghost;
  op->launches++;
  return Kernels<dim, T>::matrix_columns(op->n, op->increment_form ? BR_NEWTON : BR_FIXED_POINT,
                                         (uint32_t)(op->n_owned + op->n_ghost), p, op->shape, s);
}

template <typename T>
__global__ void k_store_column(double *__restrict__ A, const T *__restrict__ col, uint64_t n, uint64_t j)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    A[i * n + j] = (double)col[i];
}
} // namespace

extern "C" {

int glsb_create(const glsb_desc *d, glsb_op **out)
{
  if (!d || !out)
    return fail(nullptr, "glsb_create: null argument");
  *out = nullptr;
  if (d->abi_version != GLSB_ABI_VERSION)
    return fail(nullptr, "glsb_create: ABI version mismatch");
  if (d->dim != 2 && d->dim != 3)
    return fail(nullptr, "glsb_create: dim must be 2 or 3");
  if (d->degree < 1 || d->degree > 4)
    return fail(nullptr, "glsb_create: degree must be 1..4");
  if (d->number_type != GLSB_F64 && d->number_type != GLSB_F32)
    return fail(nullptr, "glsb_create: unknown number_type");
  if (d->geometry_type != GLSB_GEOM_CARTESIAN && d->geometry_type != GLSB_GEOM_GENERAL)
    return fail(nullptr, "glsb_create: unknown geometry_type");
  if (d->n_cells == 0 || d->n_cells > 0x7fffffffull)
    return fail(nullptr, "glsb_create: n_cells out of range");
  if (d->n_owned + d->n_ghost >= 0x80000000ull)
    return fail(nullptr, "glsb_create: local vector too long for 31-bit indices");
  if (d->consider_time_derivative && d->time_order > 0 && d->theta != 1.0)
    return fail(nullptr, "glsb_create: consider_time_derivative requires theta == 1 (operator_ns.cc:126-129)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, "glsb_create: no CUDA device available (there is no CPU fallback)");
  if (d->device < 0 || d->device >= ndev)
    return fail(nullptr, "glsb_create: bad device ordinal");
  if (cudaSetDevice(d->device) != cudaSuccess)
    return cuda_fail(nullptr, "cudaSetDevice");

  glsb_op *op          = new glsb_op;
  op->dim              = d->dim;
  op->degree           = d->degree;
  op->n                = d->degree + 1;
  op->C                = d->dim + 1;
  op->n_loc            = (d->dim == 2) ? op->n * op->n : op->n * op->n * op->n;
  op->nq               = op->n_loc;
  op->number_type      = d->number_type;
  op->tsize            = d->number_type == GLSB_F64 ? 8 : 4;
  op->increment_form   = d->increment_form;
  op->time_order       = d->time_order;
  op->ctd              = d->consider_time_derivative && d->time_order > 0; // operator_ns.cc:97-98
  op->cell_wise        = d->cell_wise_stabilization;
  op->nu               = d->nu;
  op->c1               = d->c1;
  op->c2               = d->c2;
  op->theta            = d->theta;
  op->geom             = d->geometry_type;
  op->device           = d->device;
  op->n_cells          = d->n_cells;
  op->n_owned          = d->n_owned;
  op->n_ghost          = d->n_ghost;
  op->n_rows           = d->n_constraint_rows;
  op->n_constrained    = d->n_constrained_indices;
  op->n_export         = d->n_export;
  compute_shape_host(d->degree, op->shape);

  const uint32_t ndof    = op->C * op->n_loc;
  const uint64_t n_local = d->n_owned + d->n_ghost;
  const uint32_t nc      = (uint32_t)d->n_cells;
  bool           ok      = true;
  std::string    why;

  // ---- validate indices; classify cells: interior first, then cells touching ghosts ----
  std::vector<uint8_t> row_ghost(d->n_constraint_rows, 0), row_weighted(d->n_constraint_rows, 0);
  for (uint32_t r = 0; r < d->n_constraint_rows && ok; ++r)
    {
      if (d->row_dof[r] >= n_local)
        {
          ok  = false;
          why = "constraint row_dof out of range";
        }
      for (uint32_t e = d->row_ptr[r]; e < d->row_ptr[r + 1]; ++e)
        {
          if (d->entry_col[e] >= n_local)
            {
              ok  = false;
              why = "constraint entry_col out of range";
              break;
            }
          if (d->entry_col[e] >= d->n_owned)
            row_ghost[r] = 1;
          row_weighted[r] = 1;
        }
      if (d->row_dof[r] >= d->n_owned)
        row_ghost[r] = 1;
    }
  std::vector<uint8_t>  is_boundary(nc, 0), has_weighted(nc, 0), has_constrained(nc, 0);
  std::vector<uint32_t> cell_lo(nc, 0xffffffffu), cell_hi(nc, 0);
  for (uint32_t k = 0; k < nc && ok; ++k)
    {
      const uint32_t *row = d->dof_indices + (uint64_t)k * ndof;
      uint32_t        lo = 0xffffffffu, hi = 0;
      for (uint32_t j = 0; j < ndof; ++j)
        {
          const uint32_t iv = row[j];
          if (iv & GLSB_CONSTRAINED_BIT)
            {
              const uint32_t r = iv & ~GLSB_CONSTRAINED_BIT;
              if (r >= d->n_constraint_rows)
                {
                  ok  = false;
                  why = "dof_indices refers to a constraint row that does not exist";
                  break;
                }
              has_constrained[k] = 1;
              is_boundary[k] |= row_ghost[r];
              has_weighted[k] |= row_weighted[r];
              lo = std::min(lo, d->row_dof[r]);
              hi = std::max(hi, d->row_dof[r]);
              for (uint32_t e = d->row_ptr[r]; e < d->row_ptr[r + 1]; ++e)
                {
                  lo = std::min(lo, d->entry_col[e]);
                  hi = std::max(hi, d->entry_col[e]);
                }
            }
          else
            {
              if (iv >= n_local)
                {
                  ok  = false;
                  why = "dof_indices out of range";
                  break;
                }
              if (iv >= d->n_owned)
                is_boundary[k] = 1;
              lo = std::min(lo, iv);
              hi = std::max(hi, iv);
            }
        }
      cell_lo[k] = lo;
      cell_hi[k] = hi;
    }
  if (!ok)
    {
      delete op;
      return fail(nullptr, "glsb_create: " + why);
    }
  // internal slots: [interior cells | padding up to a multiple of 32 | boundary cells | padding];
  // padding slots repeat a real cell (safe to read) and are never scattered
  std::vector<uint32_t> perm;
  perm.reserve(nc + 160);
  const bool deterministic = getenv("GLSB_DETERMINISTIC") != nullptr && atoi(getenv("GLSB_DETERMINISTIC")) != 0;
  if (deterministic)
    {
      // Bit-reproducible summation: greedy colouring of the cells over the vector entries they scatter to (masters
      // of weighted rows included; 8 colours on a structured hexahedral mesh), cells reordered colour by colour
      // (stable, so the Morton locality inside a colour survives).  Two cells of one colour never add to the same
      // entry, colours run as consecutive launches: every entry is summed in a fixed order whatever the timing.
      if (d->n_ghost != 0)
        {
          delete op;
          return fail(nullptr, "glsb_create: GLSB_DETERMINISTIC is for single-rank operators (n_ghost == 0)");
        }
      std::vector<uint32_t> mask(n_local, 0u), color(nc, 0u);
      uint32_t              n_colors = 0;
      for (uint32_t k = 0; k < nc; ++k)
        {
          const uint32_t *row  = d->dof_indices + (uint64_t)k * ndof;
          uint32_t        used = 0;
          auto            visit = [&](auto &&f) {
            for (uint32_t j = 0; j < ndof; ++j)
              {
                const uint32_t iv = row[j];
                if (iv & GLSB_CONSTRAINED_BIT)
                  {
                    const uint32_t r = iv & ~GLSB_CONSTRAINED_BIT;
                    for (uint32_t e = d->row_ptr[r]; e < d->row_ptr[r + 1]; ++e)
                      f(d->entry_col[e]);
                  }
                else
                  f(iv);
              }
          };
          visit([&](uint32_t i) { used |= mask[i]; });
          uint32_t c = 0;
          while (c < 32 && (used >> c) & 1u)
            ++c;
          if (c == 32)
            {
              delete op;
              return fail(nullptr, "glsb_create: GLSB_DETERMINISTIC needs more than 32 colours on this mesh");
            }
          color[k] = c;
          n_colors = std::max(n_colors, c + 1);
          visit([&](uint32_t i) { mask[i] |= 1u << c; });
        }
      op->color_begin.assign(n_colors + 1, 0);
      for (uint32_t c = 0; c < n_colors; ++c)
        {
          op->color_begin[c] = (uint32_t)perm.size();
          for (uint32_t k = 0; k < nc; ++k)
            if (color[k] == c)
              perm.push_back(k);
        }
      op->color_begin[n_colors] = (uint32_t)perm.size();
    }
  else
    for (uint32_t k = 0; k < nc; ++k)
      if (!is_boundary[k])
        perm.push_back(k);
  op->n_interior = (uint32_t)perm.size();
  while (perm.size() % 32 != 0)
    perm.push_back(perm.empty() ? 0 : perm.back());
  op->n_int_pad = (uint32_t)perm.size();
  for (uint32_t k = 0; k < nc; ++k)
    if (is_boundary[k])
      perm.push_back(k);
  op->n_slots = (uint32_t)perm.size();
  op->ncp     = ((uint64_t)op->n_slots + 127) / 128 * 128;
  while (perm.size() < op->ncp)
    perm.push_back(perm.back());
  auto slot_is_real = [&](uint64_t i) { return i < op->n_interior || (i >= op->n_int_pad && i < op->n_slots); };
  op->slot_lo.assign((op->n_slots + 31) / 32, 0xffffffffu);
  op->slot_hi.assign((op->n_slots + 31) / 32, 0);
  for (uint32_t i = 0; i < op->n_slots; ++i)
    if (slot_is_real(i))
      {
        op->slot_lo[i >> 5] = std::min(op->slot_lo[i >> 5], cell_lo[perm[i]]);
        op->slot_hi[i >> 5] = std::max(op->slot_hi[i >> 5], cell_hi[perm[i]]);
      }

  ok = ok && upload(op->perm, perm.data(), perm.size() * 4);
  {
    std::vector<uint8_t> fl(op->ncp);
    for (uint64_t i = 0; i < op->ncp; ++i)
      fl[i] = has_constrained[perm[i]];
    ok = ok && upload(op->cell_flags, fl.data(), fl.size());
  }

  // ---- dof indices: [cell][dof] -> [dof][ncp] in internal order -------------------------
  {
    DevBuf raw;
    ok = ok && upload(raw, d->dof_indices, (size_t)nc * ndof * 4);
    // one extra (zeroed) batch behind the last one: the packed float kernel stages batches in pairs
    ok = ok && op->idx.alloc((size_t)(ndof + 1) * (op->ncp + 32) * 4);
    ok = ok && cudaMemset(op->idx.p, 0, op->idx.bytes) == cudaSuccess;
    if (ok)
      {
        const uint64_t tot = (uint64_t)(ndof + 1) * op->ncp;
        k_transpose_idx<<<(unsigned)((tot + 255) / 256), 256>>>(raw.as<uint32_t>(), op->perm.as<uint32_t>(),
                                                                 op->cell_flags.as<uint8_t>(), op->idx.as<uint32_t>(), nc,
                                                                 ndof, (uint32_t)op->n_loc, op->ncp);
        ok = cudaDeviceSynchronize() == cudaSuccess;
      }
  }

  // ---- constraints ----------------------------------------------------------------------
  {
    const uint32_t              nr = d->n_constraint_rows;
    static const uint32_t       zero2[2] = {0, 0};
    const uint32_t              ne = nr ? d->row_ptr[nr] : 0;
    ok = ok && upload(op->row_dof, nr ? d->row_dof : zero2, (nr ? nr : 1) * 4);
    ok = ok && upload(op->row_ptr, nr ? d->row_ptr : zero2, (size_t)(nr + 1) * 4);
    ok = ok && upload(op->ecol, ne ? d->entry_col : zero2, (ne ? ne : 1) * 4);
    const double zd = 0;
    if (op->number_type == GLSB_F64)
      ok = ok && upload(op->eval, ne ? d->entry_val : &zd, (ne ? ne : 1) * 8);
    else
      ok = ok && upload_converted<float>(op->eval, ne ? d->entry_val : &zd, ne ? ne : 1);
    // sorted: glsb_vmult_host applies the identity on constrained rows range by range
    op->cidx_sorted.assign(d->constrained_indices, d->constrained_indices + d->n_constrained_indices);
    std::sort(op->cidx_sorted.begin(), op->cidx_sorted.end());
    ok = ok && upload(op->cidx, d->n_constrained_indices ? op->cidx_sorted.data() : zero2,
                      (size_t)(d->n_constrained_indices ? d->n_constrained_indices : 1) * 4);
    ok = ok && upload(op->export_idx, d->n_export ? d->export_indices : zero2,
                      (size_t)(d->n_export ? d->n_export : 1) * 4);
  }

  // ---- columns of C_cell for compute_diagonal on cells with weighted rows ---------------
  {
    std::vector<uint8_t>  skip(op->ncp, 0);
    std::vector<uint32_t> l_cell, col_ptr(1, 0), col_dof, ent_ptr(1, 0), ent_loc;
    std::vector<double>   ent_val;
    for (uint32_t i = 0; i < op->n_slots; ++i)
      {
        const uint32_t k = perm[i];
        if (!slot_is_real(i) || !has_weighted[k])
          continue;
        skip[i] = 1;
        l_cell.push_back(i);
        // gather (g, local, weight) triples, then group by g
        std::vector<std::pair<uint32_t, std::pair<uint32_t, double>>> tr;
        const uint32_t *row = d->dof_indices + (uint64_t)k * ndof;
        for (uint32_t j = 0; j < ndof; ++j)
          {
            const uint32_t iv = row[j];
            if (iv & GLSB_CONSTRAINED_BIT)
              {
                const uint32_t r = iv & ~GLSB_CONSTRAINED_BIT;
                for (uint32_t e = d->row_ptr[r]; e < d->row_ptr[r + 1]; ++e)
                  tr.push_back({d->entry_col[e], {j, d->entry_val[e]}});
              }
            else
              tr.push_back({iv, {j, 1.0}});
          }
        std::stable_sort(tr.begin(), tr.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
        for (size_t t = 0; t < tr.size(); ++t)
          {
            if (t == 0 || tr[t].first != tr[t - 1].first)
              {
                if (t)
                  ent_ptr.push_back((uint32_t)ent_loc.size());
                col_dof.push_back(tr[t].first);
              }
            ent_loc.push_back(tr[t].second.first);
            ent_val.push_back(tr[t].second.second);
          }
        if (!tr.empty())
          ent_ptr.push_back((uint32_t)ent_loc.size());
        col_ptr.push_back((uint32_t)col_dof.size());
      }
    op->dc_n_list = (uint32_t)l_cell.size();
    ok            = ok && upload(op->diag_skip, skip.data(), skip.size());
    if (op->dc_n_list)
      {
        ok = ok && upload(op->dc_cell, l_cell.data(), l_cell.size() * 4);
        ok = ok && upload(op->dc_col_ptr, col_ptr.data(), col_ptr.size() * 4);
        ok = ok && upload(op->dc_col_dof, col_dof.data(), col_dof.size() * 4);
        ok = ok && upload(op->dc_ent_ptr, ent_ptr.data(), ent_ptr.size() * 4);
        ok = ok && upload(op->dc_ent_loc, ent_loc.data(), ent_loc.size() * 4);
        ok = ok && upload(op->dc_ent_val, ent_val.data(), ent_val.size() * 8);
      }
  }

  // ---- q-point data layout: the fields the Newton-branch vmult streams come first ----------
  {
    const int dm = op->dim;
    int       f  = 0;
    auto take = [&](int nf) {
      const int o = f;
      f += nf;
      return o;
    };
    const bool general = op->geom == GLSB_GEOM_GENERAL;
    op->fU = take(dm);
    op->fH = take(dm * dm);
    op->fP = take(dm);
    if (op->ctd)
      op->fO = take(dm);
    if (!op->cell_wise)
      {
        op->fd1q = take(1);
        op->fd2q = take(1);
      }
    if (general)
      {
        op->fJ   = take(dm * dm);
        op->fjxw = take(1);
      }
    op->F_stage = f;
    if (!op->ctd && op->time_order > 0)
      op->fO = take(dm);
    if (op->cell_wise)
      {
        op->fd1q = take(1);
        op->fd2q = take(1);
      }
    if (op->theta != 1.0)
      {
        op->fGold  = take(dm * dm);
        op->fgoldp = take(dm);
      }
    op->FT = f;
    // degrees with a register-tiled kernel (glsb_q2.cuh): Q2, Q1, and Q3 in float
    op->regtile = op->dim == 3 && (op->n <= 3 || ((op->n == 4 || op->n == 5) && op->number_type == GLSB_F32));
    // Q4 float keeps the interpolated values in shared memory (glsb_q2.cuh, TSM): row stages only
    const bool tsm = op->n == 5 && op->number_type == GLSB_F32;
    if (op->regtile && op->n != 3)
      {
        // one stage = one quadrature layer (n^2 points) if 2-3 CTAs with a 2-deep ring fit, else one row
        const size_t layer = (size_t)op->F_stage * op->n * op->n * 32 * op->tsize;
        const size_t fixed = (size_t)4 * 2 * 180 * op->tsize + 2 * (4 * op->nq + 1) * 32 * 4 + 256;
        const int    ctas  = (op->number_type == GLSB_F32 && op->n < 4) ? 3 : 2;
        const bool   fits  = !tsm && ctas * (2 * layer + fixed + 1024) <= 228 * 1024;
        const char  *rows  = getenv("GLSB_Q2_ROWS");
        op->QG             = rows ? (atoi(rows) == 1 ? op->n : op->n * op->n) : (fits ? op->n * op->n : op->n);
        op->NL             = op->nq / op->QG;
      }
    else if (op->dim == 3 && op->n == 3)
      {
        // one block per ring stage of the Q2 kernel: a whole quadrature layer (9 points) if two CTAs with a
        // 2-deep ring of such stages fit into an SM's shared memory, else one (qz, qy) row of 3 points
        // (measured, config C in FP64: 73 % instead of 55 % of the HBM roofline with row stages)
        const char  *rows  = getenv("GLSB_Q2_ROWS");
        const size_t layer = (size_t)op->F_stage * 9 * 32 * op->tsize;
        const size_t fixed = (size_t)4 * 2 * 180 * op->tsize + 2 * 109 * 32 * 4 + 256;
        const bool   fits  = 2 * (2 * layer + fixed + 1024) <= 228 * 1024;
        // opt-in (GLSB_Q2_PACK=1): the packed float kernel, two cells per lane with FFMA2.  Measured on config P
        // in FP32: 48.4 GDoF/s against 47.4 for the scalar kernel (29 % fewer instructions per cell, but the
        // 64-bit register footprint allows 8 instead of 12 warps per SM), 25.3 against 28.5 on config C: not
        // the default.  It stages two batches per slot, hence row stages.
        op->packed         = op->number_type == GLSB_F32 && getenv("GLSB_Q2_PACK") != nullptr;
        op->QG             = rows ? (atoi(rows) == 1 ? 3 : 9) : ((fits && !op->packed) ? 9 : 3);
        op->NL           = 27 / op->QG;
      }
    else
      {
        op->NL = 1;
        op->QG = op->nq;
      }
    const size_t bytes = (size_t)op->FT * op->nq * (op->ncp + 32) * op->tsize; // + one batch, see idx
    ok = ok && op->Q.alloc(bytes) && op->d1c.alloc(op->ncp * op->tsize) && op->d2c.alloc(op->ncp * op->tsize);
    if (ok)
      ok = cudaMemset(op->Q.p, 0, bytes) == cudaSuccess;
  }

  // ---- geometry -------------------------------------------------------------------------
  {
    const int dm = op->dim;
    // per-cell scalars in internal order, padded
    std::vector<double> hm(op->ncp), ms(op->ncp);
    for (uint64_t i = 0; i < op->ncp; ++i)
      {
        const uint32_t k = perm[i];
        hm[i]            = d->cell_h_min[k];
        ms[i]            = d->cell_measure[k];
      }
    ok = ok && upload(op->h_min, hm.data(), hm.size() * 8) && upload(op->measure, ms.data(), ms.size() * 8);
    // ---- boundary faces with outflow terms ----
    if (ok && d->n_outflow_faces > 0)
      {
        const uint32_t nf = d->n_outflow_faces, nqf = (uint32_t)(op->dim == 2 ? op->n : op->n * op->n), dm = (uint32_t)op->dim;
        std::vector<uint32_t> slot_of_cell(nc, 0xffffffffu), fslot(nf);
        for (uint32_t i = 0; i < op->n_slots; ++i)
          if (!(i >= op->n_interior && i < op->n_int_pad) && slot_of_cell[perm[i]] == 0xffffffffu)
            slot_of_cell[perm[i]] = i;
        std::vector<double> beta(nf);
        for (uint32_t f = 0; f < nf && ok; ++f)
          {
            if (d->face_cell[f] >= nc || d->face_no[f] >= 2 * dm || (d->face_kind[f] != 1 && d->face_kind[f] != 2))
              ok = false;
            else
              {
                fslot[f] = slot_of_cell[d->face_cell[f]];
                // effective_beta_face = beta / h^(p+1), beta = 1, h after Lethe (operator_ns.cc:427-458)
                const double m = d->cell_measure[d->face_cell[f]];
                const double h = (op->dim == 2 ? std::sqrt(4.0 * m / M_PI) : std::pow(6.0 * m / M_PI, 1.0 / 3.0)) / op->degree;
                beta[f] = 1.0 / std::pow(h, (double)(op->degree + 1));
              }
          }
        if (!ok)
          {
            delete op;
            return fail(nullptr, "glsb_create: bad outflow face arrays");
          }
        // the face part of the inverse diagonal (k_faces_diag) handles plain and zero-constrained dofs only:
        // hanging-node (weighted) rows on an outflow-face cell would silently lose their face contribution to
        // diag(C^T A C) (MatrixFreeTools::compute_diagonal includes it, operator_ns.cc:203-218) -- refuse loudly
        for (uint32_t f = 0; f < nf; ++f)
          if (has_weighted[d->face_cell[f]])
            {
              delete op;
              return fail(nullptr, "glsb_create: weighted (hanging-node) constraint rows on a cell with an outflow "
                                   "face are not supported by the face part of compute_inverse_diagonal");
            }
        op->n_faces = nf;
        compute_face_basis_host(op->degree, op->Nf, op->Gf);
        std::vector<double> zeros((size_t)nf * nqf * dm, 0.0);
        const double *tgt = d->face_target_velocity ? d->face_target_velocity : zeros.data();
        ok = upload(op->f_slot, fslot.data(), nf * 4) && upload(op->f_no, d->face_no, nf * 4) &&
             upload(op->f_kind, d->face_kind, nf * 4);
        if (op->number_type == GLSB_F64)
          ok = ok && upload_converted<double>(op->f_normal, d->face_normal, (size_t)nf * nqf * dm) &&
               upload_converted<double>(op->f_jxw, d->face_jxw, (size_t)nf * nqf) &&
               upload_converted<double>(op->f_inv_jac, d->face_inv_jac, (size_t)nf * nqf * dm * dm) &&
               upload_converted<double>(op->f_target, tgt, (size_t)nf * nqf * dm) &&
               upload_converted<double>(op->f_beta, beta.data(), nf) &&
               upload_converted<double>(op->f_velocity, zeros.data(), (size_t)nf * nqf * dm);
        else
          ok = ok && upload_converted<float>(op->f_normal, d->face_normal, (size_t)nf * nqf * dm) &&
               upload_converted<float>(op->f_jxw, d->face_jxw, (size_t)nf * nqf) &&
               upload_converted<float>(op->f_inv_jac, d->face_inv_jac, (size_t)nf * nqf * dm * dm) &&
               upload_converted<float>(op->f_target, tgt, (size_t)nf * nqf * dm) &&
               upload_converted<float>(op->f_beta, beta.data(), nf) &&
               upload_converted<float>(op->f_velocity, zeros.data(), (size_t)nf * nqf * dm);
      }
    if (op->geom == GLSB_GEOM_CARTESIAN)
      {
        std::vector<double> ij((size_t)dm * op->ncp), dj(op->ncp);
        for (uint64_t i = 0; i < op->ncp; ++i)
          {
            const uint32_t k = perm[i];
            for (int e = 0; e < dm; ++e)
              ij[e * op->ncp + i] = d->inv_jac[(uint64_t)k * dm + e];
            dj[i] = d->jxw[k];
          }
        if (op->number_type == GLSB_F64)
          ok = ok && upload(op->inv_jac, ij.data(), ij.size() * 8) && upload(op->jxw, dj.data(), dj.size() * 8);
        else
          ok = ok && upload_converted<float>(op->inv_jac, ij.data(), ij.size()) &&
               upload_converted<float>(op->jxw, dj.data(), dj.size());
      }
    else
      {
        const uint32_t inner = dm * dm;
        DevBuf         raw;
        const uint64_t tot1 = (uint64_t)inner * op->nq * op->ncp, tot2 = (uint64_t)op->nq * op->ncp;
        ok = ok && upload(raw, d->inv_jac, (size_t)nc * op->nq * inner * 8);
        if (ok)
          {
            if (op->number_type == GLSB_F64)
              k_fill_geom<double><<<(unsigned)((tot1 + 255) / 256), 256>>>(raw.as<double>(), op->perm.as<uint32_t>(),
                                                                           op->Q.as<double>(), op->nq, inner, op->ncp,
                                                                           op->FT, op->NL, op->QG, op->fJ);
            else
              k_fill_geom<float><<<(unsigned)((tot1 + 255) / 256), 256>>>(raw.as<double>(), op->perm.as<uint32_t>(),
                                                                          op->Q.as<float>(), op->nq, inner, op->ncp,
                                                                          op->FT, op->NL, op->QG, op->fJ);
            ok = cudaDeviceSynchronize() == cudaSuccess;
          }
        ok = ok && upload(raw, d->jxw, (size_t)nc * op->nq * 8);
        if (ok)
          {
            if (op->number_type == GLSB_F64)
              k_fill_geom<double><<<(unsigned)((tot2 + 255) / 256), 256>>>(raw.as<double>(), op->perm.as<uint32_t>(),
                                                                           op->Q.as<double>(), op->nq, 1, op->ncp,
                                                                           op->FT, op->NL, op->QG, op->fjxw);
            else
              k_fill_geom<float><<<(unsigned)((tot2 + 255) / 256), 256>>>(raw.as<double>(), op->perm.as<uint32_t>(),
                                                                          op->Q.as<float>(), op->nq, 1, op->ncp,
                                                                          op->FT, op->NL, op->QG, op->fjxw);
            ok = cudaDeviceSynchronize() == cudaSuccess;
          }
      }
  }
  ok = ok && op->max_bits.alloc(8);
  op->n_edge   = d->n_edge_constrained_indices;
  op->has_edge = d->has_edge_constrained_indices || d->n_edge_constrained_indices > 0;
  if (op->n_edge)
    {
      for (uint32_t i = 0; i < op->n_edge && ok; ++i)
        if (d->edge_constrained_indices[i] >= d->n_owned)
          {
            ok = false;
            cudaGetLastError();
            delete op;
            return fail(nullptr, "glsb_create: edge_constrained_indices must be owned indices");
          }
      ok = ok && upload(op->edge_idx, d->edge_constrained_indices, (size_t)op->n_edge * 4) &&
           op->edge_saved.alloc((size_t)op->n_edge * op->tsize);
    }

  if (!ok)
    {
      cudaError_t e = cudaGetLastError();
      std::string m = std::string("glsb_create: device setup failed: ") + cudaGetErrorString(e);
      delete op;
      return fail(nullptr, m);
    }
  *out = op;
  return 0;
}

void glsb_destroy(glsb_op *op)
{
  if (op)
    {
      cudaSetDevice(op->device);
      for (glsb_op::HostPipe *h : {&op->hpp})
        {
          for (cudaEvent_t e : h->e_in)
            cudaEventDestroy(e);
          for (cudaEvent_t e : h->e_cells)
            cudaEventDestroy(e);
          if (h->e_start)
            cudaEventDestroy(h->e_start);
          if (h->e_done)
            cudaEventDestroy(h->e_done);
          if (h->s_in)
            cudaStreamDestroy(h->s_in);
          if (h->s_out)
            cudaStreamDestroy(h->s_out);
        }
      for (cudaEvent_t e : op->hp.e_in)
        cudaEventDestroy(e);
      for (cudaEvent_t e : op->hp.e_cells)
        cudaEventDestroy(e);
      if (op->hp.e_start)
        cudaEventDestroy(op->hp.e_start);
      if (op->hp.e_done)
        cudaEventDestroy(op->hp.e_done);
      if (op->hp.s_in)
        cudaStreamDestroy(op->hp.s_in);
      if (op->hp.s_out)
        cudaStreamDestroy(op->hp.s_out);
      delete op;
    }
}

const char *glsb_last_error(const glsb_op *op) { return op ? op->err.c_str() : g_create_error.c_str(); }

int glsb_invalidate_system(glsb_op *op)
{
  if (!op)
    return 1;
  return 0; // the assembled system matrix stays with the retained CPU operator (coarse level)
}

int glsb_vmult_begin(glsb_op *op, void *dst, void *stream)
{
  if (!op || !dst)
    return fail(op, "glsb_vmult_begin: null argument");
  if (cudaMemsetAsync(dst, 0, (op->n_owned + op->n_ghost) * op->tsize, (cudaStream_t)stream) != cudaSuccess)
    return cuda_fail(op, "glsb_vmult_begin: memset");
  return 0; // a memset node, not one of our kernels: not counted in launches
}

int glsb_vmult_cells(glsb_op *op, void *dst, const void *src, double weight, int which, void *stream)
{
  if (!op || !dst || !src)
    return fail(op, "glsb_vmult_cells: null argument");
  if (!op->lin_valid)
    return fail(op, "glsb_vmult: set_linearization_point has not been called");
  if (op->increment_form && op->ctd && !op->prev_valid)
    return fail(op, "glsb_vmult: set_previous_solution has not been called");
  const int branch = op->increment_form ? BR_NEWTON : BR_FIXED_POINT;
  int       rc     = 0;
#define CALL(D, T) do_cells<D, T>(op, dst, src, weight, which, branch, (cudaStream_t)stream)
  GLSB_DISPATCH(op, CALL);
#undef CALL
  if (rc)
    return cuda_fail(op, "glsb_vmult_cells: launch");
  return 0;
}

int glsb_vmult_cells_part(glsb_op *op, void *dst, const void *src, double weight, int which, int part, int n_parts,
                          void *stream)
{
  if (!op || !dst || !src)
    return fail(op, "glsb_vmult_cells_part: null argument");
  if (n_parts < 1 || part < 0 || part >= n_parts)
    return fail(op, "glsb_vmult_cells_part: bad part");
  if (!op->lin_valid)
    return fail(op, "glsb_vmult: set_linearization_point has not been called");
  if (op->increment_form && op->ctd && !op->prev_valid)
    return fail(op, "glsb_vmult: set_previous_solution has not been called");
  const int branch = op->increment_form ? BR_NEWTON : BR_FIXED_POINT;
  int       rc     = 0;
#define CALL(D, T) do_cells<D, T>(op, dst, src, weight, which, branch, (cudaStream_t)stream, part, n_parts)
  GLSB_DISPATCH(op, CALL);
#undef CALL
  if (rc)
    return cuda_fail(op, "glsb_vmult_cells_part: launch");
  return 0;
}

int glsb_set_sm_reserve(glsb_op *op, int n_sms)
{
  if (!op || n_sms < 0)
    return 1;
  op->sm_reserve = n_sms;
  return 0;
}

int glsb_vmult_finish(glsb_op *op, void *dst, const void *src, void *stream)
{
  if (!op || !dst || !src)
    return fail(op, "glsb_vmult_finish: null argument");
  if (op->n_constrained == 0)
    return 0;
  const unsigned g = (op->n_constrained + 255) / 256;
  if (op->number_type == GLSB_F64)
    k_copy_indexed<double><<<g, 256, 0, (cudaStream_t)stream>>>((double *)dst, (const double *)src,
                                                                op->cidx.as<uint32_t>(), op->n_constrained);
  else
    k_copy_indexed<float><<<g, 256, 0, (cudaStream_t)stream>>>((float *)dst, (const float *)src,
                                                               op->cidx.as<uint32_t>(), op->n_constrained);
  op->launches++;
  if (cudaGetLastError() != cudaSuccess)
    return cuda_fail(op, "glsb_vmult_finish");
  return 0;
}

int glsb_edge_begin(glsb_op *op, void *src, void *stream)
{
  if (!op || !src)
    return fail(op, "glsb_edge_begin: null argument");
  if (op->n_edge == 0)
    return 0;
  const unsigned g = (op->n_edge + 255) / 256;
  if (op->number_type == GLSB_F64)
    k_edge_save_zero<double><<<g, 256, 0, (cudaStream_t)stream>>>((double *)src, op->edge_saved.as<double>(),
                                                                  op->edge_idx.as<uint32_t>(), op->n_edge);
  else
    k_edge_save_zero<float><<<g, 256, 0, (cudaStream_t)stream>>>((float *)src, op->edge_saved.as<float>(),
                                                                 op->edge_idx.as<uint32_t>(), op->n_edge);
  op->launches++;
  return cudaGetLastError() != cudaSuccess ? cuda_fail(op, "glsb_edge_begin") : 0;
}

int glsb_edge_finish(glsb_op *op, void *dst, void *src, void *stream)
{
  if (!op || !dst || !src)
    return fail(op, "glsb_edge_finish: null argument");
  if (op->n_edge == 0)
    return 0;
  const unsigned g = (op->n_edge + 255) / 256;
  if (op->number_type == GLSB_F64)
    k_edge_restore<double><<<g, 256, 0, (cudaStream_t)stream>>>((double *)dst, (double *)src,
                                                                op->edge_saved.as<double>(),
                                                                op->edge_idx.as<uint32_t>(), op->n_edge);
  else
    k_edge_restore<float><<<g, 256, 0, (cudaStream_t)stream>>>((float *)dst, (float *)src, op->edge_saved.as<float>(),
                                                               op->edge_idx.as<uint32_t>(), op->n_edge);
  op->launches++;
  return cudaGetLastError() != cudaSuccess ? cuda_fail(op, "glsb_edge_finish") : 0;
}

int glsb_vmult_interface_down(glsb_op *op, void *dst, const void *src, double weight, void *stream)
{
  int rc = glsb_vmult_begin(op, dst, stream);
  if (rc == 0)
    rc = glsb_vmult_cells(op, dst, src, weight, GLSB_CELLS_ALL, stream);
  if (rc == 0)
    rc = glsb_vmult_finish(op, dst, src, stream);
  return rc;
}

int glsb_edge_extract(glsb_op *op, void *cpy, const void *src, void *stream)
{
  if (!op || !cpy || !src)
    return fail(op, "glsb_edge_extract: null argument");
  if (cudaMemsetAsync(cpy, 0, (op->n_owned + op->n_ghost) * op->tsize, (cudaStream_t)stream) != cudaSuccess)
    return cuda_fail(op, "glsb_edge_extract: memset");
  if (op->n_edge == 0)
    return 0;
  const unsigned g = (op->n_edge + 255) / 256;
  if (op->number_type == GLSB_F64)
    k_copy_indexed<double><<<g, 256, 0, (cudaStream_t)stream>>>((double *)cpy, (const double *)src,
                                                                op->edge_idx.as<uint32_t>(), op->n_edge);
  else
    k_copy_indexed<float><<<g, 256, 0, (cudaStream_t)stream>>>((float *)cpy, (const float *)src,
                                                               op->edge_idx.as<uint32_t>(), op->n_edge);
  op->launches++;
  return cudaGetLastError() != cudaSuccess ? cuda_fail(op, "glsb_edge_extract") : 0;
}

int glsb_vmult_interface_up(glsb_op *op, void *dst, const void *src, double weight, void *stream)
{
  if (!op || !dst || !src)
    return fail(op, "glsb_vmult_interface_up: null argument");
  if (op->n_ghost != 0)
    return fail(op, "glsb_vmult_interface_up: with ghost entries use glsb_edge_extract + ghost import + "
                    "glsb_vmult_cells + compress");
  int rc = glsb_vmult_begin(op, dst, stream); // dst = 0
  if (rc || !op->has_edge)
    return rc;
  const size_t bytes = (op->n_owned + op->n_ghost) * op->tsize;
  if (op->edge_cpy.bytes != bytes && !op->edge_cpy.alloc(bytes))
    return cuda_fail(op, "glsb_vmult_interface_up: scratch vector");
  rc = glsb_edge_extract(op, op->edge_cpy.p, src, stream);
  if (rc == 0)
    rc = glsb_vmult_cells(op, dst, op->edge_cpy.p, weight, GLSB_CELLS_ALL, stream);
  return rc; // no identity on the constrained rows here (operator_ns.cc:776-786)
}

int glsb_vmult(glsb_op *op, void *dst, const void *src, double weight, void *stream)
{
  // the reference mutates src at the edge indices during the loop and restores it (operator_ns.cc:692-700)
  int rc = glsb_edge_begin(op, const_cast<void *>(src), stream);
  if (rc == 0)
    rc = glsb_vmult_begin(op, dst, stream);
  if (rc == 0)
    rc = glsb_vmult_cells(op, dst, src, weight, GLSB_CELLS_ALL, stream);
  if (rc == 0)
    rc = glsb_vmult_finish(op, dst, src, stream);
  if (rc == 0)
    rc = glsb_edge_finish(op, dst, const_cast<void *>(src), stream);
  return rc;
}


int glsb_get_system_matrix(glsb_op *op, double *A_dev, double weight, void *stream)
{
  if (!op || !A_dev)
    return 1;
  // partitioned operators: the matrix of THIS rank's cell loop over its local dofs [owned | ghost]; the host
  // layer sums the ranks' matrices into the global one (multigrid.py, MGCoarseGridDirect)
  const uint64_t n = op->n_owned + op->n_ghost;
  if (op->n_ghost != 0 && (op->n_faces != 0 || op->n_edge != 0))
    return fail(op, "glsb_get_system_matrix: partitioned operators with outflow faces or edge indices are not supported");
  if (!op->lin_valid)
    return fail(op, "glsb_get_system_matrix: set_linearization_point has not been called");
  if (op->n_faces == 0 && op->n_edge == 0 && n <= 65535)
    {
      // all columns in one launch: blockIdx.y = column, the unit vectors are never materialised
      cudaStream_t s0 = static_cast<cudaStream_t>(stream);
      if (op->mat_col.bytes < n * n * op->tsize && !op->mat_col.alloc(n * n * op->tsize))
        return fail(op, "glsb_get_system_matrix: out of device memory");
      cudaMemsetAsync(op->mat_col.p, 0, n * n * op->tsize, s0);
      int rc = 0;
#define CALL(D, T) do_matrix<D, T>(op, op->mat_col.p, weight, s0)
      GLSB_DISPATCH(op, CALL);
#undef CALL
      if (rc)
        return cuda_fail(op, "glsb_get_system_matrix: launch");
      const unsigned g = (unsigned)((n * n + 255) / 256);
      if (op->number_type == GLSB_F64)
        k_matrix_finish<double><<<g, 256, 0, s0>>>(A_dev, op->mat_col.as<double>(), n);
      else
        k_matrix_finish<float><<<g, 256, 0, s0>>>(A_dev, op->mat_col.as<float>(), n);
      if (op->n_constrained)
        k_matrix_identity_rows<<<(op->n_constrained + 255) / 256, 256, 0, s0>>>(A_dev, op->cidx.as<uint32_t>(),
                                                                               op->n_constrained, n);
      op->launches += 2;
      return cudaGetLastError() != cudaSuccess ? cuda_fail(op, "glsb_get_system_matrix") : 0;
    }
  // operators with outflow faces or edge indices: column by column through glsb_vmult
  if ((op->mat_e.bytes < n * op->tsize && !op->mat_e.alloc(n * op->tsize)) ||
      (op->mat_col.bytes < n * op->tsize && !op->mat_col.alloc(n * op->tsize)))
    return fail(op, "glsb_get_system_matrix: out of device memory");
  cudaStream_t   s      = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  for (uint64_t j = 0; j < n; ++j)
    {
      if (op->number_type == GLSB_F64)
        k_unit_vector<double><<<blocks, 256, 0, s>>>(op->mat_e.as<double>(), n, j);
      else
        k_unit_vector<float><<<blocks, 256, 0, s>>>(op->mat_e.as<float>(), n, j);
      const int rc = glsb_vmult(op, op->mat_col.p, op->mat_e.p, weight, stream);
      if (rc)
        return rc;
      if (op->number_type == GLSB_F64)
        k_store_column<double><<<blocks, 256, 0, s>>>(A_dev, op->mat_col.as<double>(), n, j);
      else
        k_store_column<float><<<blocks, 256, 0, s>>>(A_dev, op->mat_col.as<float>(), n, j);
    }
  return cudaGetLastError() != cudaSuccess;
}


// vmult with HOST vectors.  The cells are cut into chunks (internal order); chunk c needs src[0, in_end[c]), so
// the upload of src and the cell kernels are pipelined on two streams.  The download of dst is pipelined too:
//   * speculative mode (dst_host is page-locked and device-accessible): dst[in_end[c-1], in_end[c]) is sent as
//     soon as chunk c is done; the few entries of that range that a LATER chunk still adds to (chunk-surface
//     nodes, found once on the device: k_last_touch / k_late_flags) are re-sent at the end by a kernel that
//     stores straight into the host vector.  Works for any numbering, e.g. deal.II's first-touch numbering along
//     the Morton curve, where low-numbered surface nodes are touched by cells far down the cell order;
//   * conservative mode (pageable dst_host): dst[0, out_end[c]) is sent once no later chunk touches it.
// PCIe is full duplex, so the total is ~ max(upload, download) instead of upload + kernels + download.
// late_flag -> late_list (n_late entries); leaves late_list empty when more than 1/8 of the entries are late
static void compact_late(glsb_op::HostPipe &hp, uint64_t n)
{
  const uint32_t cap = (uint32_t)std::min<uint64_t>(n / 8 + 1024, 0x7fffffffull);
  DevBuf         cnt;
  if (!hp.late_list.alloc((size_t)cap * 4) || !cnt.alloc(8))
    {
      hp.late_list.release();
      return;
    }
  cudaMemset(cnt.p, 0, 8);
  k_compact_flags<<<(unsigned)((n + 255) / 256), 256>>>(hp.late_flag.as<uint8_t>(), n, hp.late_list.as<uint32_t>(), cap,
                                                        cnt.as<unsigned long long>());
  unsigned long long m = 0;
  if (cudaMemcpy(&m, cnt.p, 8, cudaMemcpyDeviceToHost) != cudaSuccess || m > cap)
    {
      hp.late_list.release();
      return;
    }
  hp.n_late = m;
}

static int host_pipe_setup(glsb_op *op)
{
  glsb_op::HostPipe &hp = op->hp;
  if (hp.ready)
    return 0;
  const uint64_t n_local = op->n_owned + op->n_ghost;
  const uint32_t nb      = (op->n_slots + 31) / 32;
  const char    *env     = getenv("GLSB_HOST_CHUNKS");
  uint32_t       nch     = env ? (uint32_t)atoi(env) : 16;
  if (nch < 1)
    nch = 1;
  if (nch > 64)
    nch = 64;
  if (nb < 8 * nch)
    nch = 1;
  hp.cell_begin.resize(nch + 1);
  for (uint32_t c = 0; c <= nch; ++c)
    hp.cell_begin[c] = (uint32_t)((uint64_t)nb * c / nch) * 32;
  hp.cell_begin[nch] = op->n_slots;
  std::vector<uint64_t> lo(nch, n_local), hi(nch, 0);
  std::vector<uint8_t>  chunk_of_batch(nb, 0);
  for (uint32_t c = 0; c < nch; ++c)
    for (uint32_t b = hp.cell_begin[c] / 32; b < (hp.cell_begin[c + 1] + 31) / 32 && b < nb; ++b)
      {
        chunk_of_batch[b] = (uint8_t)c;
        if (op->slot_lo[b] != 0xffffffffu)
          {
            lo[c] = std::min<uint64_t>(lo[c], op->slot_lo[b]);
            hi[c] = std::max<uint64_t>(hi[c], (uint64_t)op->slot_hi[b] + 1);
          }
      }
  hp.in_end.resize(nch);
  hp.out_end.resize(nch);
  hp.cidx_end.resize(nch);
  hp.cidx_end_spec.resize(nch);
  uint64_t run = 0;
  for (uint32_t c = 0; c < nch; ++c)
    {
      run          = std::max(run, hi[c]);
      hp.in_end[c] = (c + 1 == nch) ? n_local : run;
    }
  run = n_local;
  for (uint32_t c = nch; c-- > 0;)
    {
      hp.out_end[c] = (c + 1 == nch) ? n_local : run; // nothing of chunks > c starts below `run`
      run           = std::min(run, lo[c]);
    }
  auto below = [&](uint64_t x) {
    return (uint32_t)(std::lower_bound(op->cidx_sorted.begin(), op->cidx_sorted.end(),
                                       (uint32_t)std::min<uint64_t>(x, 0xffffffffull)) -
                      op->cidx_sorted.begin());
  };
  for (uint32_t c = 0; c < nch; ++c)
    {
      hp.cidx_end[c]      = below(hp.out_end[c]);
      hp.cidx_end_spec[c] = below(hp.in_end[c]);
    }
  hp.cidx_end[nch - 1] = hp.cidx_end_spec[nch - 1] = (uint32_t)op->cidx_sorted.size();
  if (!hp.src.alloc(n_local * op->tsize) || !hp.dst.alloc(n_local * op->tsize))
    return 1;
  // entries that change after their speculative download
  {
    DevBuf last, cob, d_in_end;
    if (!last.alloc(n_local * 4) || !hp.late_flag.alloc(n_local) || !upload(cob, chunk_of_batch.data(), nb) ||
        !upload(d_in_end, hp.in_end.data(), nch * 8))
      return 1;
    cudaMemset(last.p, 0, n_local * 4);
    const uint32_t ndof = (uint32_t)(op->C * op->n_loc);
    const uint32_t slots_all = (op->n_slots + 31) / 32 * 32; // whole batches, see k_last_touch
    const uint64_t tot       = (uint64_t)ndof * slots_all;
    k_last_touch<<<(unsigned)((tot + 255) / 256), 256>>>(op->idx.as<uint32_t>(), cob.as<uint8_t>(),
                                                         op->row_dof.as<uint32_t>(), op->row_ptr.as<uint32_t>(),
                                                         op->ecol.as<uint32_t>(), last.as<uint32_t>(), slots_all, ndof);
    k_late_flags<<<(unsigned)((n_local + 255) / 256), 256>>>(last.as<uint32_t>(), d_in_end.as<uint64_t>(), nch,
                                                             hp.late_flag.as<uint8_t>(), n_local);
    if (cudaDeviceSynchronize() != cudaSuccess)
      return 1;
    compact_late(hp, n_local);
  }
  if (cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking) != cudaSuccess)
    return 1;
  hp.e_in.resize(nch);
  hp.e_cells.resize(nch);
  for (uint32_t c = 0; c < nch; ++c)
    if (cudaEventCreateWithFlags(&hp.e_in[c], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&hp.e_cells[c], cudaEventDisableTiming) != cudaSuccess)
      return 1;
  if (cudaEventCreateWithFlags(&hp.e_start, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&hp.e_done, cudaEventDisableTiming) != cudaSuccess)
    return 1;
  hp.ready = true;
  return 0;
}

int glsb_vmult_host(glsb_op *op, void *dst_host, const void *src_host, double weight, void *stream)
{
  if (op && op->n_faces > 0)
    return fail(op, "glsb_vmult_host: operators with outflow faces use glsb_vmult on device vectors");
  if (!op || !dst_host || !src_host)
    return fail(op, "glsb_vmult_host: null argument");
  if (op->n_ghost != 0)
    return fail(op, "glsb_vmult_host: operators with ghost entries exchange on the device; use glsb_vmult");
  if (!op->lin_valid)
    return fail(op, "glsb_vmult: set_linearization_point has not been called");
  if (op->increment_form && op->ctd && !op->prev_valid)
    return fail(op, "glsb_vmult: set_previous_solution has not been called");
  if (host_pipe_setup(op))
    return cuda_fail(op, "glsb_vmult_host: setup");
  glsb_op::HostPipe &hp     = op->hp;
  cudaStream_t       s      = (cudaStream_t)stream;
  const uint32_t     nch    = (uint32_t)hp.in_end.size();
  const size_t       ts     = op->tsize;
  const uint64_t     n      = op->n_owned + op->n_ghost;
  const int          branch = op->increment_form ? BR_NEWTON : BR_FIXED_POINT;
  char              *d_src = hp.src.as<char>(), *d_dst = hp.dst.as<char>();
  // speculative download needs a device-visible (page-locked, mapped) destination
  void *dst_mapped = nullptr;
  {
    cudaPointerAttributes at;
    static const bool     off = getenv("GLSB_HOST_NO_SPEC") != nullptr;
    if (!off && nch > 1 && cudaPointerGetAttributes(&at, dst_host) == cudaSuccess && at.type == cudaMemoryTypeHost &&
        at.devicePointer != nullptr)
      dst_mapped = at.devicePointer;
    cudaGetLastError();
  }
  const std::vector<uint64_t> &out_end  = dst_mapped ? hp.in_end : hp.out_end;
  const std::vector<uint32_t> &cidx_end = dst_mapped ? hp.cidx_end_spec : hp.cidx_end;
  // everything below is ordered after the work already enqueued on the caller's stream
  cudaEventRecord(hp.e_start, s);
  cudaStreamWaitEvent(hp.s_in, hp.e_start, 0);
  cudaStreamWaitEvent(hp.s_out, hp.e_start, 0);
  if (cudaMemsetAsync(d_dst, 0, (size_t)n * ts, s) != cudaSuccess)
    return cuda_fail(op, "glsb_vmult_host: memset");
  uint64_t in_done = 0, out_done = 0;
  uint32_t c_done = 0;
  for (uint32_t c = 0; c < nch; ++c)
    {
      if (hp.in_end[c] > in_done)
        {
          if (cudaMemcpyAsync(d_src + in_done * ts, (const char *)src_host + in_done * ts, (hp.in_end[c] - in_done) * ts,
                              cudaMemcpyHostToDevice, hp.s_in) != cudaSuccess)
            return cuda_fail(op, "glsb_vmult_host: upload");
          in_done = hp.in_end[c];
        }
      cudaEventRecord(hp.e_in[c], hp.s_in);
      cudaStreamWaitEvent(s, hp.e_in[c], 0);
      int rc = 0;
      {
        // cells [cell_begin[c], cell_begin[c + 1]) of the internal order (the padding hole is skipped)
        op->range_override[0] = hp.cell_begin[c];
        op->range_override[1] = hp.cell_begin[c + 1];
#define CALL(D, T) do_cells<D, T>(op, d_dst, d_src, weight, GLSB_CELLS_ALL, branch, s)
        GLSB_DISPATCH(op, CALL);
#undef CALL
        op->range_override[0] = op->range_override[1] = 0;
      }
      if (rc)
        return cuda_fail(op, "glsb_vmult_host: launch");
      // identity on the constrained rows inside the range sent after this chunk (operator_ns.cc:719-721)
      if (cidx_end[c] > c_done)
        {
          const uint32_t m = cidx_end[c] - c_done;
          if (op->number_type == GLSB_F64)
            k_copy_indexed<double><<<(m + 255) / 256, 256, 0, s>>>((double *)d_dst, (const double *)d_src,
                                                                 op->cidx.as<uint32_t>() + c_done, m);
          else
            k_copy_indexed<float><<<(m + 255) / 256, 256, 0, s>>>((float *)d_dst, (const float *)d_src,
                                                                op->cidx.as<uint32_t>() + c_done, m);
          op->launches++;
          c_done = cidx_end[c];
        }
      cudaEventRecord(hp.e_cells[c], s);
      if (out_end[c] > out_done)
        {
          cudaStreamWaitEvent(hp.s_out, hp.e_cells[c], 0);
          if (cudaMemcpyAsync((char *)dst_host + out_done * ts, d_dst + out_done * ts, (out_end[c] - out_done) * ts,
                              cudaMemcpyDeviceToHost, hp.s_out) != cudaSuccess)
            return cuda_fail(op, "glsb_vmult_host: download");
          out_done = out_end[c];
        }
    }
  cudaEventRecord(hp.e_done, hp.s_out);
  cudaStreamWaitEvent(s, hp.e_done, 0);
  if (dst_mapped)
    {
      // after the last range has landed: re-send what later chunks changed
      if (hp.late_list.p)
        {
          const uint32_t m = (uint32_t)hp.n_late;
          if (m && op->number_type == GLSB_F64)
            k_flush_listed<double><<<(m + 255) / 256, 256, 0, s>>>((double *)dst_mapped, (const double *)d_dst,
                                                                 hp.late_list.as<uint32_t>(), m);
          else if (m)
            k_flush_listed<float><<<(m + 255) / 256, 256, 0, s>>>((float *)dst_mapped, (const float *)d_dst,
                                                                hp.late_list.as<uint32_t>(), m);
        }
      else if (op->number_type == GLSB_F64)
        k_flush_flagged<double><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((double *)dst_mapped, (const double *)d_dst,
                                                                          hp.late_flag.as<uint8_t>(), n);
      else
        k_flush_flagged<float><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((float *)dst_mapped, (const float *)d_dst,
                                                                         hp.late_flag.as<uint8_t>(), n);
      op->launches++;
    }
  if (cudaGetLastError() != cudaSuccess)
    return cuda_fail(op, "glsb_vmult_host");
  return 0;
}

// ---- the same pipeline for partitioned operators (n_ghost > 0) -------------------------------------------
// The interior cells (no ghost dofs) are chunked exactly like above; the cells at the partition surface and the
// two halves of the ghost exchange stay with the host layer between _begin and _finish:
//   glsb_vmult_host_begin : zero d_dst; per chunk: upload src, interior cells of the chunk, speculative download
//   host layer            : update_ghost_values(d_src) -> glsb_vmult_cells(BOUNDARY) -> compress(add)(d_dst)
//   glsb_vmult_host_finish: constrained rows, last download, re-send of the entries that changed after their
//                           download (touched by a later chunk, by the boundary cells or by compress(add))
static int host_pipe_setup_part(glsb_op *op)
{
  glsb_op::HostPipe &hp = op->hpp;
  if (hp.ready)
    return 0;
  const uint64_t n_local = op->n_owned + op->n_ghost;
  const uint32_t nb_all = (op->n_slots + 31) / 32, nb = op->n_int_pad / 32;
  const char    *env = getenv("GLSB_HOST_CHUNKS");
  uint32_t       nch = env ? (uint32_t)atoi(env) : 16;
  nch                = std::max(1u, std::min(64u, nch));
  if (nb < 8 * nch)
    nch = 1;
  hp.cell_begin.resize(nch + 1);
  for (uint32_t c = 0; c <= nch; ++c)
    hp.cell_begin[c] = (uint32_t)((uint64_t)nb * c / nch) * 32;
  hp.cell_begin[nch] = op->n_interior;
  std::vector<uint64_t> hi(nch, 0);
  std::vector<uint8_t>  chunk_of_batch(nb_all, (uint8_t)nch); // batches of boundary cells: after every chunk
  for (uint32_t c = 0; c < nch; ++c)
    for (uint32_t b = hp.cell_begin[c] / 32; b < (hp.cell_begin[c + 1] + 31) / 32 && b < nb; ++b)
      {
        chunk_of_batch[b] = (uint8_t)c;
        if (op->slot_lo[b] != 0xffffffffu)
          hi[c] = std::max<uint64_t>(hi[c], (uint64_t)op->slot_hi[b] + 1);
      }
  hp.in_end.resize(nch);
  hp.cidx_end_spec.resize(nch);
  uint64_t run = 0;
  for (uint32_t c = 0; c < nch; ++c)
    {
      run          = std::min<uint64_t>(std::max(run, hi[c]), op->n_owned);
      hp.in_end[c] = (c + 1 == nch) ? op->n_owned : run;
    }
  for (uint32_t c = 0; c < nch; ++c)
    hp.cidx_end_spec[c] = (uint32_t)(std::lower_bound(op->cidx_sorted.begin(), op->cidx_sorted.end(),
                                                      (uint32_t)std::min<uint64_t>(hp.in_end[c], 0xffffffffull)) -
                                     op->cidx_sorted.begin());
  {
    DevBuf last, cob, d_in_end;
    if (!last.alloc(n_local * 4) || !hp.late_flag.alloc(n_local) || !upload(cob, chunk_of_batch.data(), nb_all) ||
        !upload(d_in_end, hp.in_end.data(), nch * 8))
      return 1;
    cudaMemset(last.p, 0, n_local * 4);
    const uint32_t ndof = (uint32_t)(op->C * op->n_loc);
    const uint32_t slots_all = (op->n_slots + 31) / 32 * 32; // whole batches, see k_last_touch
    const uint64_t tot       = (uint64_t)ndof * slots_all;
    k_last_touch<<<(unsigned)((tot + 255) / 256), 256>>>(op->idx.as<uint32_t>(), cob.as<uint8_t>(),
                                                         op->row_dof.as<uint32_t>(), op->row_ptr.as<uint32_t>(),
                                                         op->ecol.as<uint32_t>(), last.as<uint32_t>(), slots_all, ndof);
    if (op->n_export) // compress(add) changes the exported entries last
      k_mark_last<<<(unsigned)((op->n_export + 255) / 256), 256>>>(op->export_idx.as<uint32_t>(), op->n_export, nch + 1,
                                                                   last.as<uint32_t>());
    k_late_flags<<<(unsigned)((op->n_owned + 255) / 256), 256>>>(last.as<uint32_t>(), d_in_end.as<uint64_t>(), nch,
                                                                 hp.late_flag.as<uint8_t>(), op->n_owned);
    if (cudaDeviceSynchronize() != cudaSuccess)
      return 1;
    compact_late(hp, op->n_owned);
  }
  if (cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking) != cudaSuccess)
    return 1;
  hp.e_in.resize(nch);
  hp.e_cells.resize(nch);
  for (uint32_t c = 0; c < nch; ++c)
    if (cudaEventCreateWithFlags(&hp.e_in[c], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&hp.e_cells[c], cudaEventDisableTiming) != cudaSuccess)
      return 1;
  if (cudaEventCreateWithFlags(&hp.e_start, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&hp.e_done, cudaEventDisableTiming) != cudaSuccess)
    return 1;
  hp.ready = true;
  return 0;
}

int glsb_vmult_host_begin(glsb_op *op, void *d_dst_v, void *d_src_v, void *dst_host, const void *src_host,
                          double weight, void *stream)
{
  if (!op || !d_dst_v || !d_src_v || !dst_host || !src_host)
    return fail(op, "glsb_vmult_host_begin: null argument");
  if (op->n_faces > 0 || op->n_edge > 0)
    return fail(op, "glsb_vmult_host_begin: operators with outflow faces or edge indices use the device-vector calls");
  if (!op->lin_valid)
    return fail(op, "glsb_vmult: set_linearization_point has not been called");
  if (op->increment_form && op->ctd && !op->prev_valid)
    return fail(op, "glsb_vmult: set_previous_solution has not been called");
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, dst_host) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer)
    {
      cudaGetLastError();
      return fail(op, "glsb_vmult_host_begin: dst_host must be page-locked (glsb_host_register)");
    }
  if (host_pipe_setup_part(op))
    return cuda_fail(op, "glsb_vmult_host_begin: setup");
  glsb_op::HostPipe &hp     = op->hpp;
  cudaStream_t       s      = (cudaStream_t)stream;
  const uint32_t     nch    = (uint32_t)hp.in_end.size();
  const size_t       ts     = op->tsize;
  const int          branch = op->increment_form ? BR_NEWTON : BR_FIXED_POINT;
  char              *d_src = (char *)d_src_v, *d_dst = (char *)d_dst_v;
  cudaEventRecord(hp.e_start, s);
  cudaStreamWaitEvent(hp.s_in, hp.e_start, 0);
  cudaStreamWaitEvent(hp.s_out, hp.e_start, 0);
  if (cudaMemsetAsync(d_dst, 0, (size_t)(op->n_owned + op->n_ghost) * ts, s) != cudaSuccess)
    return cuda_fail(op, "glsb_vmult_host_begin: memset");
  uint64_t in_done = 0, out_done = 0;
  uint32_t c_done = 0;
  for (uint32_t c = 0; c < nch; ++c)
    {
      if (hp.in_end[c] > in_done)
        {
          if (cudaMemcpyAsync(d_src + in_done * ts, (const char *)src_host + in_done * ts, (hp.in_end[c] - in_done) * ts,
                              cudaMemcpyHostToDevice, hp.s_in) != cudaSuccess)
            return cuda_fail(op, "glsb_vmult_host_begin: upload");
          in_done = hp.in_end[c];
        }
      cudaEventRecord(hp.e_in[c], hp.s_in);
      cudaStreamWaitEvent(s, hp.e_in[c], 0);
      int rc = 0;
      if (hp.cell_begin[c + 1] > hp.cell_begin[c])
        {
          op->range_override[0] = hp.cell_begin[c];
          op->range_override[1] = hp.cell_begin[c + 1];
#define CALL(D, T) do_cells<D, T>(op, d_dst, d_src, weight, GLSB_CELLS_ALL, branch, s)
          GLSB_DISPATCH(op, CALL);
#undef CALL
          op->range_override[0] = op->range_override[1] = 0;
        }
      if (rc)
        return cuda_fail(op, "glsb_vmult_host_begin: launch");
      if (hp.cidx_end_spec[c] > c_done) // identity on the constrained rows inside the range sent now
        {
          const uint32_t m = hp.cidx_end_spec[c] - c_done;
          if (op->number_type == GLSB_F64)
            k_copy_indexed<double><<<(m + 255) / 256, 256, 0, s>>>((double *)d_dst, (const double *)d_src,
                                                                 op->cidx.as<uint32_t>() + c_done, m);
          else
            k_copy_indexed<float><<<(m + 255) / 256, 256, 0, s>>>((float *)d_dst, (const float *)d_src,
                                                                op->cidx.as<uint32_t>() + c_done, m);
          op->launches++;
          c_done = hp.cidx_end_spec[c];
        }
      cudaEventRecord(hp.e_cells[c], s);
      if (hp.in_end[c] > out_done)
        {
          cudaStreamWaitEvent(hp.s_out, hp.e_cells[c], 0);
          if (cudaMemcpyAsync((char *)dst_host + out_done * ts, d_dst + out_done * ts, (hp.in_end[c] - out_done) * ts,
                              cudaMemcpyDeviceToHost, hp.s_out) != cudaSuccess)
            return cuda_fail(op, "glsb_vmult_host_begin: download");
          out_done = hp.in_end[c];
        }
    }
  if (cudaGetLastError() != cudaSuccess)
    return cuda_fail(op, "glsb_vmult_host_begin");
  return 0;
}

int glsb_vmult_host_finish(glsb_op *op, void *d_dst_v, void *dst_host, void *stream)
{
  if (!op || !d_dst_v || !dst_host || !op->hpp.ready)
    return fail(op, "glsb_vmult_host_finish: call glsb_vmult_host_begin first");
  glsb_op::HostPipe &hp = op->hpp;
  cudaStream_t       s  = (cudaStream_t)stream;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, dst_host) != cudaSuccess || !at.devicePointer)
    return fail(op, "glsb_vmult_host_finish: dst_host must be page-locked");
  // the downloads of _begin have to land before the late entries are re-sent
  cudaEventRecord(hp.e_done, hp.s_out);
  cudaStreamWaitEvent(s, hp.e_done, 0);
  const uint64_t n = op->n_owned;
  if (hp.late_list.p)
    {
      const uint32_t m = (uint32_t)hp.n_late;
      if (m && op->number_type == GLSB_F64)
        k_flush_listed<double><<<(m + 255) / 256, 256, 0, s>>>((double *)at.devicePointer, (const double *)d_dst_v,
                                                             hp.late_list.as<uint32_t>(), m);
      else if (m)
        k_flush_listed<float><<<(m + 255) / 256, 256, 0, s>>>((float *)at.devicePointer, (const float *)d_dst_v,
                                                            hp.late_list.as<uint32_t>(), m);
    }
  else if (op->number_type == GLSB_F64)
    k_flush_flagged<double><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((double *)at.devicePointer, (const double *)d_dst_v,
                                                                      hp.late_flag.as<uint8_t>(), n);
  else
    k_flush_flagged<float><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((float *)at.devicePointer, (const float *)d_dst_v,
                                                                     hp.late_flag.as<uint8_t>(), n);
  op->launches++;
  if (cudaGetLastError() != cudaSuccess)
    return cuda_fail(op, "glsb_vmult_host_finish");
  return 0;
}

/* page-lock a host vector so that glsb_vmult_host's copies are asynchronous DMA transfers */
int glsb_host_register(void *ptr, uint64_t bytes)
{
  return cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault) == cudaSuccess ? 0 : 1;
}
int glsb_host_unregister(void *ptr) { return cudaHostUnregister(ptr) == cudaSuccess ? 0 : 1; }

// ---- relaxation smoother -----------------------------------------------------------------------
extern "C++" {
template <typename T>
static int relax_update(glsb_op *op, void *x, const void *t, const void *b, const void *d, double omega, cudaStream_t s)
{
  const uint64_t n = op->n_owned; // ghost entries take no part in vector operations
  if (n == 0)
    return 0;
  if (t)
    k_relax_update<T><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((T *)x, (const T *)t, (const T *)b, (const T *)d,
                                                                  (T)omega, n);
  else
    k_relax_first<T><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((T *)x, (const T *)b, (const T *)d, (T)omega, n);
  op->launches++;
  return cudaGetLastError() != cudaSuccess;
}
} // extern "C++"

static int relax_sweeps(glsb_op *op, void *dst, const void *src, const void *inv_diag, double omega, int n_it,
                        double weight, bool from_zero, void *stream)
{
  if (!op || !dst || !src || !inv_diag)
    return fail(op, "glsb_relaxation: null argument");
  if (n_it < 0)
    return fail(op, "glsb_relaxation: negative n_iterations");
  if (op->n_ghost != 0)
    return fail(op, "glsb_relaxation: with ghost entries drive the sweeps from the host layer "
                    "(exchange-aware vmult + glsb_relaxation_update)");
  cudaStream_t s     = (cudaStream_t)stream;
  const size_t bytes = (op->n_owned + op->n_ghost) * op->tsize;
  if (op->relax_t.bytes != bytes && !op->relax_t.alloc(bytes))
    return cuda_fail(op, "glsb_relaxation: scratch vector");
  const bool f64 = op->number_type == GLSB_F64;
  int        it  = 0;
  if (from_zero)
    {
      if (n_it == 0)
        return cudaMemsetAsync(dst, 0, bytes, s) != cudaSuccess ? cuda_fail(op, "glsb_relaxation: memset") : 0;
      if (f64 ? relax_update<double>(op, dst, nullptr, src, inv_diag, omega, s) :
                relax_update<float>(op, dst, nullptr, src, inv_diag, omega, s))
        return cuda_fail(op, "glsb_relaxation: first sweep");
      it = 1;
    }
  for (; it < n_it; ++it)
    {
      int rc = glsb_vmult(op, op->relax_t.p, dst, weight, stream);
      if (rc)
        return rc;
      if (f64 ? relax_update<double>(op, dst, op->relax_t.p, src, inv_diag, omega, s) :
                relax_update<float>(op, dst, op->relax_t.p, src, inv_diag, omega, s))
        return cuda_fail(op, "glsb_relaxation: update");
    }
  return 0;
}

int glsb_relaxation_vmult(glsb_op *op, void *dst, const void *src, const void *inv_diag, double omega,
                          int n_iterations, double weight, void *stream)
{
  return relax_sweeps(op, dst, src, inv_diag, omega, n_iterations, weight, true, stream);
}

int glsb_relaxation_step(glsb_op *op, void *dst, const void *src, const void *inv_diag, double omega, int n_iterations,
                         double weight, void *stream)
{
  return relax_sweeps(op, dst, src, inv_diag, omega, n_iterations, weight, false, stream);
}

int glsb_relaxation_update(glsb_op *op, void *x, const void *t, const void *b, const void *inv_diag, double omega,
                           void *stream)
{
  if (!op || !x || !t || !b || !inv_diag)
    return fail(op, "glsb_relaxation_update: null argument");
  const int rc = op->number_type == GLSB_F64 ? relax_update<double>(op, x, t, b, inv_diag, omega, (cudaStream_t)stream) :
                                               relax_update<float>(op, x, t, b, inv_diag, omega, (cudaStream_t)stream);
  return rc ? cuda_fail(op, "glsb_relaxation_update") : 0;
}

extern "C++" {
template <typename T>
static int estimate_relaxation(glsb_op *op, const void *inv_diag, int n_it, double weight, uint64_t first,
                               double *lambda, cudaStream_t s)
{
  const uint64_t n     = op->n_owned;
  const size_t   bytes = (op->n_owned + op->n_ghost) * op->tsize;
  const unsigned g     = (unsigned)((n + 255) / 256);
  if ((op->pi_e.bytes != bytes && !op->pi_e.alloc(bytes)) || (op->pi_v1.bytes != bytes && !op->pi_v1.alloc(bytes)) ||
      (op->pi_v2.bytes != bytes && !op->pi_v2.alloc(bytes)) || !op->pi_sums.alloc((size_t)(3 + n_it) * 8))
    return 1;
  T      *e = op->pi_e.as<T>(), *v1 = op->pi_v1.as<T>(), *v2 = op->pi_v2.as<T>();
  double *sums = op->pi_sums.as<double>(), *hist = sums + 3;
  cudaMemsetAsync(e, 0, bytes, s);
  cudaMemsetAsync(sums, 0, (size_t)(3 + n_it) * 8, s);
  // set_initial_guess: (global index) % 11, mean-free; constraints.set_zero; eigenvector /= l2_norm
  const unsigned gr = g < 148u * 16u ? g : 148u * 16u; // reductions: grid-stride, one atomic per block
  k_pi_init<T><<<gr, 256, 0, s>>>(e, first, n, sums);
  k_pi_shift<T><<<g, 256, 0, s>>>(e, sums, n);
  if (op->n_constrained)
    k_set_indexed<T><<<(op->n_constrained + 255) / 256, 256, 0, s>>>(e, op->cidx.as<uint32_t>(), op->n_constrained, T(0));
  k_pi_record<<<1, 1, 0, s>>>(sums, hist, -1);
  k_pi_apply<T><<<gr, 256, 0, s>>>(v1, (const T *)nullptr, (const T *)nullptr, e, n, sums);
  k_pi_normalise<T><<<g, 256, 0, s>>>(e, e, n, sums);
  k_pi_record<<<1, 1, 0, s>>>(sums, hist, -1);
  op->launches += 6;
  for (int k = 0; k < n_it; ++k)
    {
      if (glsb_vmult(op, v2, e, weight, s))
        return 1;
      k_pi_apply<T><<<gr, 256, 0, s>>>(v1, v2, (const T *)inv_diag, e, n, sums); // vector1 = D^-1 A e; e . vector1
      k_pi_normalise<T><<<g, 256, 0, s>>>(e, v1, n, sums);                      // vector1 /= |vector1|; swap
      k_pi_record<<<1, 1, 0, s>>>(sums, hist, k);
      op->launches += 3;
    }
  if (cudaGetLastError() != cudaSuccess)
    return 1;
  *lambda = 0;
  if (n_it > 0 && cudaMemcpyAsync(lambda, hist + n_it - 1, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess)
    return 1;
  return cudaStreamSynchronize(s) != cudaSuccess;
}
} // extern "C++"

int glsb_estimate_relaxation(glsb_op *op, const void *inv_diag, int n_power_iterations, double smoothing_range,
                             double weight, uint64_t first_local_index, double *omega_out, double *ev_max_out,
                             void *stream)
{
  if (!op || !inv_diag || !omega_out)
    return fail(op, "glsb_estimate_relaxation: null argument");
  if (op->n_ghost != 0)
    return fail(op, "glsb_estimate_relaxation: single-rank operators only");
  if (n_power_iterations < 1)
    return fail(op, "glsb_estimate_relaxation: n_power_iterations must be positive");
  double    lambda = 0;
  const int rc     = op->number_type == GLSB_F64 ?
                       estimate_relaxation<double>(op, inv_diag, n_power_iterations, weight, first_local_index, &lambda,
                                               (cudaStream_t)stream) :
                       estimate_relaxation<float>(op, inv_diag, n_power_iterations, weight, first_local_index, &lambda,
                                              (cudaStream_t)stream);
  if (rc)
    return cuda_fail(op, "glsb_estimate_relaxation");
  // deal.II: max_eigenvalue_estimate = 1.2 * power-iteration value (safety factor);
  // alpha = max / smoothing_range (smoothing_range > 1), relaxation = 2 / (alpha + max)
  const double ev_max = 1.2 * lambda;
  const double alpha  = smoothing_range > 1.0 ? ev_max / smoothing_range : 0.9 * ev_max;
  *omega_out          = 2.0 / (alpha + ev_max);
  if (ev_max_out)
    *ev_max_out = ev_max;
  return 0;
}

int glsb_evaluate_residual_cells(glsb_op *op, void *dst, const void *src, double weight, int which, void *stream)
{
  if (!op || !dst || !src)
    return fail(op, "glsb_evaluate_residual: null argument");
  if (!op->lin_valid)
    return fail(op, "glsb_evaluate_residual: set_linearization_point has not been called");
  int rc = 0;
#define CALL(D, T) do_cells<D, T>(op, dst, src, weight, which, BR_RESIDUAL, (cudaStream_t)stream)
  GLSB_DISPATCH(op, CALL);
#undef CALL
  if (rc)
    return cuda_fail(op, "glsb_evaluate_residual: launch");
  return 0;
}

int glsb_evaluate_residual(glsb_op *op, void *dst, const void *src, double weight, void *stream)
{
  int rc = glsb_vmult_begin(op, dst, stream);
  if (rc == 0)
    rc = glsb_evaluate_residual_cells(op, dst, src, weight, GLSB_CELLS_ALL, stream);
  // constrained rows receive nothing from the scatter, so set_zero (operator_ns.cc:678) is implied
  return rc;
}

int glsb_set_linearization_point(glsb_op *op, const void *vec, double dt, void *stream)
{
  if (!op || !vec)
    return fail(op, "glsb_set_linearization_point: null argument");
  if (!ensure_tables(op, true, false))
    return cuda_fail(op, "glsb_set_linearization_point: table allocation");
  int rc = 0;
#define CALL(D, T) do_lin<D, T>(op, vec, dt, (cudaStream_t)stream)
  GLSB_DISPATCH(op, CALL);
#undef CALL
  if (rc)
    return cuda_fail(op, "glsb_set_linearization_point: launch");
  op->lin_valid = true;
  op->lin_dt    = dt;
  return 0;
}

int glsb_set_previous_solution(glsb_op *op, const void *const *history, const double *weights, int order,
                               void *stream)
{
  if (!op)
    return 1;
  if (op->time_order == 0) // operator_ns.cc:242-243
    return 0;
  if (!history || !weights)
    return fail(op, "glsb_set_previous_solution: null argument");
  if (order != op->time_order || order > 3)
    return fail(op, "glsb_set_previous_solution: order does not match the operator's time_order");
  if (!ensure_tables(op, false, true))
    return cuda_fail(op, "glsb_set_previous_solution: table allocation");
  int rc = 0;
#define CALL(D, T) do_prev<D, T>(op, history, weights, order, (cudaStream_t)stream)
  GLSB_DISPATCH(op, CALL);
#undef CALL
  if (rc)
    return cuda_fail(op, "glsb_set_previous_solution: launch");
  op->prev_valid = true;
  return 0;
}

int glsb_diagonal_cells(glsb_op *op, void *diag, double weight, void *stream)
{
  if (!op || !diag)
    return fail(op, "glsb_diagonal_cells: null argument");
  if (!op->lin_valid)
    return fail(op, "glsb_compute_inverse_diagonal: set_linearization_point has not been called");
  int rc = glsb_vmult_begin(op, diag, stream);
  if (rc)
    return rc;
#define CALL(D, T) do_diag<D, T>(op, diag, weight, (cudaStream_t)stream)
  GLSB_DISPATCH(op, CALL);
#undef CALL
  if (rc)
    return cuda_fail(op, "glsb_diagonal_cells: launch");
  return 0;
}

int glsb_diagonal_finish(glsb_op *op, void *diag, void *stream)
{
  if (!op || !diag)
    return fail(op, "glsb_diagonal_finish: null argument");
  cudaStream_t   s = (cudaStream_t)stream;
  const uint64_t n = op->n_owned; // ghosts carry nothing after compress
  if (op->number_type == GLSB_F64)
    {
      if (op->n_constrained)
        k_set_indexed<double><<<(op->n_constrained + 255) / 256, 256, 0, s>>>(
          (double *)diag, op->cidx.as<uint32_t>(), op->n_constrained, 1.0);
      // refinement-edge dofs of a GMG-LS level: diagonal = 0, which the guard turns into 1 (operator_ns.cc:219-224)
      if (op->n_edge)
        k_set_indexed<double><<<(op->n_edge + 255) / 256, 256, 0, s>>>((double *)diag, op->edge_idx.as<uint32_t>(),
                                                                       op->n_edge, 0.0);
      k_invert_guarded<double><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((double *)diag, n);
    }
  else
    {
      if (op->n_constrained)
        k_set_indexed<float><<<(op->n_constrained + 255) / 256, 256, 0, s>>>(
          (float *)diag, op->cidx.as<uint32_t>(), op->n_constrained, 1.0f);
      if (op->n_edge)
        k_set_indexed<float><<<(op->n_edge + 255) / 256, 256, 0, s>>>((float *)diag, op->edge_idx.as<uint32_t>(),
                                                                      op->n_edge, 0.0f);
      k_invert_guarded<float><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((float *)diag, n);
    }
  op->launches += 1 + (op->n_constrained > 0) + (op->n_edge > 0);
  if (cudaGetLastError() != cudaSuccess)
    return cuda_fail(op, "glsb_diagonal_finish");
  return 0;
}

int glsb_compute_inverse_diagonal(glsb_op *op, void *diag, double weight, void *stream)
{
  int rc = glsb_diagonal_cells(op, diag, weight, stream);
  if (rc == 0)
    rc = glsb_diagonal_finish(op, diag, stream);
  return rc;
}

int glsb_get_max_u(glsb_op *op, const void *vec, double *out_host, void *stream)
{
  if (!op || !vec || !out_host)
    return fail(op, "glsb_get_max_u: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(op->max_bits.p, 0, 8, s) != cudaSuccess)
    return cuda_fail(op, "glsb_get_max_u: memset");
  int rc = 0;
#define CALL(D, T) do_maxu<D, T>(op, vec, s)
  GLSB_DISPATCH(op, CALL);
#undef CALL
  if (rc)
    return cuda_fail(op, "glsb_get_max_u: launch");
  unsigned long long bits = 0;
  if (cudaMemcpyAsync(&bits, op->max_bits.p, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess)
    return cuda_fail(op, "glsb_get_max_u: readback");
  memcpy(out_host, &bits, 8);
  return 0;
}

int glsb_pack_export(glsb_op *op, void *buf, const void *vec, void *stream)
{
  if (!op || (op->n_export && (!buf || !vec)))
    return fail(op, "glsb_pack_export: null argument");
  if (op->n_export == 0)
    return 0;
  const unsigned g = (unsigned)((op->n_export + 255) / 256);
  if (op->number_type == GLSB_F64)
    k_pack<double><<<g, 256, 0, (cudaStream_t)stream>>>((double *)buf, (const double *)vec,
                                                        op->export_idx.as<uint32_t>(), op->n_export);
  else
    k_pack<float><<<g, 256, 0, (cudaStream_t)stream>>>((float *)buf, (const float *)vec,
                                                       op->export_idx.as<uint32_t>(), op->n_export);
  op->launches++;
  if (cudaGetLastError() != cudaSuccess)
    return cuda_fail(op, "glsb_pack_export");
  return 0;
}

int glsb_unpack_add(glsb_op *op, void *vec, const void *buf, void *stream)
{
  if (!op || (op->n_export && (!buf || !vec)))
    return fail(op, "glsb_unpack_add: null argument");
  if (op->n_export == 0)
    return 0;
  const unsigned g = (unsigned)((op->n_export + 255) / 256);
  if (op->number_type == GLSB_F64)
    k_unpack_add<double><<<g, 256, 0, (cudaStream_t)stream>>>((double *)vec, (const double *)buf,
                                                              op->export_idx.as<uint32_t>(), op->n_export);
  else
    k_unpack_add<float><<<g, 256, 0, (cudaStream_t)stream>>>((float *)vec, (const float *)buf,
                                                             op->export_idx.as<uint32_t>(), op->n_export);
  op->launches++;
  if (cudaGetLastError() != cudaSuccess)
    return cuda_fail(op, "glsb_unpack_add");
  return 0;
}

uint64_t glsb_n_cells(const glsb_op *op) { return op ? op->n_cells : 0; }
uint64_t glsb_n_local(const glsb_op *op) { return op ? op->n_owned + op->n_ghost : 0; }
uint64_t glsb_n_interior_cells(const glsb_op *op) { return op ? op->n_interior : 0; }
uint64_t glsb_launch_count(const glsb_op *op) { return op ? op->launches : 0; }
const char *glsb_vmult_variant(const glsb_op *op) { return op ? op->variant.c_str() : ""; }

int glsb_set_variant(glsb_op *op, int variant)
{
  if (!op)
    return 1;
  op->variant_forced = variant;
  return 0;
}

int glsb_get_table(glsb_op *op, const char *name, void *out, uint64_t out_count, void *stream)
{
  if (!op || !name || !out)
    return fail(op, "glsb_get_table: null argument");
  uint32_t      nf = 0, nq = op->nq;
  const int     d  = op->dim;
  int           f0 = -1, per_cell = 0, rowdiv = 0;
  const void   *base = op->Q.p;
  const std::string s(name);
  if (s == "u_star_value")
    f0 = op->fU, nf = d, rowdiv = 1;
  else if (s == "u_star_gradient")
    f0 = op->fH, nf = d * d, rowdiv = d;
  else if (s == "p_star_gradient")
    f0 = op->fP, nf = d, rowdiv = 1;
  else if (s == "u_time_derivative_old")
    f0 = op->fO, nf = d, rowdiv = 1;
  else if (s == "u_old_gradient")
    f0 = op->fGold, nf = d * d;
  else if (s == "p_old_gradient")
    f0 = op->fgoldp, nf = d;
  else if (s == "delta_1_q")
    f0 = op->fd1q, nf = 1;
  else if (s == "delta_2_q")
    f0 = op->fd2q, nf = 1;
  else if (s == "delta_1")
    f0 = 0, nf = 1, nq = 1, per_cell = 1, base = op->d1c.p;
  else if (s == "delta_2")
    f0 = 0, nf = 1, nq = 1, per_cell = 1, base = op->d2c.p;
  else
    return fail(op, "glsb_get_table: unknown table " + s);
  const bool is_prev = (s == "u_time_derivative_old" || s == "u_old_gradient" || s == "p_old_gradient");
  if (f0 < 0 || (is_prev && !op->prev_valid) || (!is_prev && !op->lin_valid))
    return fail(op, "glsb_get_table: table " + s + " has not been computed");
  if (out_count != (uint64_t)nf * nq * op->n_cells)
    return fail(op, "glsb_get_table: wrong output size for " + s);
  const uint64_t tot = (uint64_t)nf * nq * op->n_slots;
  const unsigned g   = (unsigned)((tot + 255) / 256);
  if (op->number_type == GLSB_F64)
    k_export_table<double><<<g, 256, 0, (cudaStream_t)stream>>>(
      (const double *)base, op->perm.as<uint32_t>(), (double *)out, (uint32_t)op->n_cells, op->n_slots,
      op->n_interior, op->n_int_pad, nq, nf, op->FT, op->NL, op->QG, f0, per_cell, rowdiv);
  else
    k_export_table<float><<<g, 256, 0, (cudaStream_t)stream>>>(
      (const float *)base, op->perm.as<uint32_t>(), (float *)out, (uint32_t)op->n_cells, op->n_slots,
      op->n_interior, op->n_int_pad, nq, nf, op->FT, op->NL, op->QG, f0, per_cell, rowdiv);
  if (cudaGetLastError() != cudaSuccess)
    return cuda_fail(op, "glsb_get_table");
  return 0;
}

} // extern "C"
