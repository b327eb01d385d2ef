// Register-tiled Q2 kernel (dim 3, degree 2): the fast path of vmult, Newton branch
// (do_vmult_cell<false> with increment_form, operator_ns.cc:1067-1182).
//
// Mapping: one thread per (cell, component); the 4 component threads of a cell are 4 adjacent
// lanes, a warp holds 8 cells, a CTA (4 warps) a batch of 32 consecutive cells.
//   * each thread keeps the 27 dof values of its component in registers and runs the whole sum
//     factorisation (x, y sweeps, then z per quadrature layer) in registers: no shared-memory
//     traffic and no CTA barrier for evaluate / integrate; 1-D matrices come from the constant
//     bank (kernel parameter) as immediate operands.
//   * the quadrature-point physics couples the components: the 4 lanes of a cell exchange
//     u, p, grad u, grad p and the SUPG residual through a small per-warp shared-memory scratch
//     (__syncwarp only, bank-conflict-free by a 33-element row stride); the transposed accesses
//     (lane c needs d_c u_j from lane j) are plain address arithmetic there, where a shuffle
//     version needs per-lane register selects that ptxas turns into divergent branches.
//   * the q-point tables (U, grad U, grad P [, du/dt_old, delta_q, J^-T, JxW]) are stored
//     [batch][row of 3 q-points][field][3][32 cells]; the block a CTA needs for one row (qz, qy) of
//     quadrature points is contiguous and is streamed HBM -> shared memory with ONE cp.async.bulk
//     (TMA engine, mbarrier complete_tx) per row into an nst-deep ring kept nst-1 rows ahead of
//     the arithmetic.
//   * the dof indices (+ one flag word per cell) of a batch are one contiguous block, staged one
//     batch ahead by a bulk copy; with them every thread gathers the 27 source values of its NEXT
//     batch with cp.async (LDGSTS) into its own shared-memory column while it computes the current
//     one (the 4 components of a node are adjacent lanes, so a node-major numbering gives full
//     32 B sectors); scatter with RED.ADD.F64 atomics.
#pragma once
#include "glsb_kernels.cuh"
#include <cstdlib>
#include <cstring>

namespace glsb
{
namespace q2
{
#ifndef GLSB_Q2_F64_CTAS
#define GLSB_Q2_F64_CTAS 2
#endif
#ifndef GLSB_Q2_GAH
#define GLSB_Q2_GAH 0 // gather-ahead staging of the next batch's source values (measured: no gain)
#endif
#ifndef GLSB_Q2_XCH
#define GLSB_Q2_XCH 0 // 1: block layout of the component exchange (one 16-byte quad per lane, broadcast reads)
#endif
#ifndef GLSB_Q2_F32_CTAS
#define GLSB_Q2_F32_CTAS 3 // resident CTAs per SM the float instantiation is compiled for
#endif
constexpr int CELLS = 32; // cells per CTA batch
constexpr int TPB   = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_LOOP:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra WAIT_DONE;\n"
               "bra WAIT_LOOP;\n"
               "WAIT_DONE:\n"
               "}" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
// bulk asynchronous copy global -> shared through the TMA engine, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                 smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// per-thread asynchronous copy global -> shared (LDGSTS), 4 or 8 bytes
template <int BYTES>
__device__ __forceinline__ void cp_async(void *dst, const void *src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- register value types -------------------------------------------------------------------------------
// V = T (double or float): one cell per lane.  V = F2 with T = float: TWO cells per lane (the same position
// of two consecutive 32-cell batches) in one 64-bit register, arithmetic with the packed FP32 instructions of
// sm_100 (add/mul/fma.f32x2 -> FADD2/FMUL2/FFMA2; ptxas contracts mul + add pairs into FFMA2).  The float
// level operators of the multigrid (config.h:7) then need half the arithmetic and exchange instructions per
// cell.
struct F2
{
  unsigned long long v;
};
__host__ __device__ __forceinline__ F2 f2_pack(float lo, float hi)
{
  F2 r;
#ifdef __CUDA_ARCH__
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
#else
  uint32_t a, b;
  memcpy(&a, &lo, 4);
  memcpy(&b, &hi, 4);
  r.v = (unsigned long long)a | ((unsigned long long)b << 32);
#endif
  return r;
}
__device__ __forceinline__ float f2_lo(F2 a) { return __uint_as_float((unsigned)(a.v & 0xffffffffull)); }
__device__ __forceinline__ float f2_hi(F2 a) { return __uint_as_float((unsigned)(a.v >> 32)); }
__device__ __forceinline__ F2 operator+(F2 a, F2 b)
{
  F2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ F2 operator-(F2 a, F2 b)
{
  F2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ F2 operator*(F2 a, F2 b)
{
  F2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ F2 &operator+=(F2 &a, F2 b) { return a = a + b; }
__device__ __forceinline__ F2 &operator*=(F2 &a, F2 b) { return a = a * b; }

template <typename V, typename T>
struct VOps // V == T
{
  static constexpr int VW = 1;
  static __host__ __device__ __forceinline__ V bcast(T s) { return s; }
  static __device__ __forceinline__ V          zero() { return V(0); }
  // element `i` of the first batch's block; the second batch's block is `sb` elements further
  static __device__ __forceinline__ V load(const T *base, int i, int) { return base[i]; }
};
template <>
struct VOps<F2, float>
{
  static constexpr int VW = 2;
  static __host__ __device__ __forceinline__ F2 bcast(float s) { return f2_pack(s, s); }
  static __device__ __forceinline__ F2          zero()
  {
    F2 r;
    r.v = 0ull;
    return r;
  }
  static __device__ __forceinline__ F2 load(const float *base, int i, int sb) { return f2_pack(base[i], base[i + sb]); }
};

template <typename T>
struct PackedOf
{
  using type = T;
};
template <>
struct PackedOf<float>
{
  using type = F2;
};
// 1-D tables with every entry duplicated into both halves of the packed type
template <typename T>
static Shape<typename PackedOf<T>::type, 3> to_packed_shape(const Shape<T, 3> &s)
{
  using V = typename PackedOf<T>::type;
  Shape<V, 3> r;
  for (int i = 0; i < 9; ++i)
    {
      r.S[i]  = VOps<V, T>::bcast(s.S[i]);
      r.D[i]  = VOps<V, T>::bcast(s.D[i]);
      r.G[i]  = VOps<V, T>::bcast(s.G[i]);
      r.Sw[i] = VOps<V, T>::bcast(s.Sw[i]);
      r.Gw[i] = VOps<V, T>::bcast(s.Gw[i]);
      r.Dt[i] = VOps<V, T>::bcast(s.Dt[i]);
    }
  for (int i = 0; i < 3; ++i)
    r.w[i] = VOps<V, T>::bcast(s.w[i]);
  return r;
}

// the value of cell v (0 or 1) held by a register value
template <typename T>
__device__ __forceinline__ T lane_value(T a, int)
{
  return a;
}
template <typename T>
__device__ __forceinline__ T lane_value(F2 a, int v)
{
  return v == 0 ? f2_lo(a) : f2_hi(a);
}

// c ? a : b as a data select (a ternary on the F2 struct compiles to a branch)
template <typename T>
__device__ __forceinline__ T vsel(bool c, T a, T b)
{
  return c ? a : b;
}
__device__ __forceinline__ F2 vsel(bool c, F2 a, F2 b)
{
  F2 r;
  r.v = c ? a.v : b.v;
  return r;
}
template <typename T>
__device__ __forceinline__ T sel3(int c, T a0, T a1, T a2)
{
  return vsel(c == 0, a0, vsel(c == 1, a1, a2));
}

// The kernel is templated on n = degree + 1 (n = 3: Q2, the tuned case; n = 2: Q1; n = 4: Q3 for the float
// level operators -- FP64 Q3 needs 2 x 64 values per lane and does not fit the register file).
// a ring stage holds ROWS rows (qz, qy) of n quadrature points: ROWS = 1 (n^2 stages per batch) or n (one
// stage per quadrature layer); the q-point array is laid out accordingly (KParams::NL = n^2 / ROWS, QG = n ROWS)
template <typename T, int ROWS, int n>
__host__ __device__ constexpr size_t stage_elems(int F)
{
  return (size_t)F * n * ROWS * CELLS;
}

template <int n>
__host__ __device__ constexpr int idx_elems() // index block of one batch: 4 n^3 dof indices + one flag word per cell
{
  return (4 * n * n * n + 1) * CELLS;
}
constexpr int MAX_NST = 8;
// resident CTAs per SM a variant is compiled for: FP64 and the 64-bit packed type 2 (255 registers), float 3
// (170 registers), float Q3 2 (2 x 64 values per lane)
template <int value_bytes, int n>
__host__ __device__ constexpr int target_ctas()
{
  return value_bytes == 4 ? (n >= 4 ? 2 : GLSB_Q2_F32_CTAS) : GLSB_Q2_F64_CTAS;
}
// entry q of a strided constant-bank row, q a runtime (CTA-uniform) index: selects instead of indexing
template <int n, typename V>
__device__ __forceinline__ V selq(int q, const V *base, int off, int stride)
{
  V r = base[off];
#pragma unroll
  for (int a = 1; a < n; ++a)
    r = vsel(q == a, base[off + a * stride], r);
  return r;
}
// exchange scratch of a warp: rows value, d_0, d_1, d_2, y.  A row holds 16 pairs (components 0/1 or 2/3
// of a cell, 16 bytes); lane (cell k, component c), h = c >> 1, owns element c & 1 of pair
// ((k + 4 h) & 7) + 8 h.  With that (measured with ncu, shared-memory wavefronts per warp instruction):
//   * the 64-bit stores of a half-warp hit 16 different bank pairs (2 wavefronts, the minimum);
//   * what the 4 lanes of a cell read in common (u, p, diagonal of grad u, y) are broadcast reads of
//     contiguous or 16-byte-strided words: 1 wavefront per 64-bit, 2 per 128-bit load;
//   * the column reads (lane c reads row 1 + c) touch 3 rows per quarter-warp: rows are 36 elements apart,
//     i.e. staggered by two 16-byte bank groups, so these 128-bit loads take 4 wavefronts, not 12.
constexpr int XROW  = 36;
constexpr int XSLOT = 5 * XROW;
template <typename T>
struct Pair;
template <>
struct alignas(16) Pair<double>
{
  double a, b;
};
template <>
struct alignas(8) Pair<float>
{
  float a, b;
};
template <>
struct alignas(16) Pair<F2>
{
  F2 a, b;
};

// value, d_0, d_1, d_2 of one (cell, component) as one 16-byte element (4-byte values only)
template <typename T>
struct alignas(16) Quad
{
  T a, b, c, d;
};

// issue the bulk copy of stage j (row j % 9 of quadrature points of this CTA's batch j / 9) into ring slot
// `slot`; called by one thread.  There is no producer warp: the slot is refilled by whichever warp is the
// last to release it (an arrival counter per slot), so no warp ever waits for another one's progress.
// A CTA works on units of VW consecutive batches (VW = 2 for the packed float kernel): a stage / an index
// slot then holds the VW blocks one after the other.
template <typename T, int ROWS, int VW, int n>
__device__ __forceinline__ void issue_stage(const KParams<T> &p, int F, T *tab, uint64_t *full, uint32_t j,
                                            uint32_t slot)
{
  constexpr uint32_t SPB   = n * n / ROWS; // stages per batch
  const uint32_t     bytes = (uint32_t)(stage_elems<T, ROWS, n>(F) * sizeof(T));
  const uint32_t     unit = blockIdx.x + (j / SPB) * gridDim.x, row = j % SPB;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(&full[slot], VW * bytes);
#pragma unroll
  for (int v = 0; v < VW; ++v)
    {
      const uint64_t batch = (uint64_t)(p.cell_begin >> 5) + (uint64_t)unit * VW + v;
      const T       *src   = p.Q + ((batch * SPB + row) * p.FT) * (n * ROWS * CELLS);
      bulk_g2s(tab + ((size_t)slot * VW + v) * stage_elems<T, ROWS, n>(F), src, bytes, &full[slot]);
    }
}

// TSM ("t in shared memory"): for Q4 in float, whose 2 x 125 values per lane do not fit the register file, the
// interpolated dof values live in a private shared-memory column of the lane during the layer loop
// (conflict-free: consecutive lanes, consecutive words) and only the accumulators stay in registers; the dof
// indices are then read from global memory instead of a staged block (no room for the ring).  Measured: Q4 float
// 22.1 against 19.3 GDoF/s for the generic kernel.  The same scheme for Q3 in double (64 + 32 accumulator and
// layer values in 64-bit registers) spills ~50 doubles and measured SLOWER than the generic kernel (14.3 against
// 15.8 GDoF/s), so FP64 Q3 / Q4 stay with the generic kernel.
template <typename T, int n>
__host__ __device__ constexpr bool use_tsm()
{
  return n == 5 && sizeof(T) == 4;
}

template <typename T, typename V, bool GENERAL, bool CTD, bool CELLWISE, int ROWS, int n>
__global__ void __launch_bounds__(TPB, (target_ctas<(int)sizeof(V), n>()))
  k_vmult_q2_newton(const KParams<T> p, const Shape<V, n> sh, const int F, const int nst, const int gah_rt)
{
  const bool gah = GLSB_Q2_GAH && gah_rt;
  constexpr bool TSM = use_tsm<T, n>();
  using VO          = VOps<V, T>;
  constexpr int VW  = VO::VW;
  constexpr int N2 = n * n, N3 = n * n * n;
  constexpr int IDX_ELEMS = idx_elems<n>();
  constexpr int ISL = VW * IDX_ELEMS;                  // index ring slot: the blocks of the unit's VW batches
  const int     SB  = (int)stage_elems<T, ROWS, n>(F); // table stage: offset of the second batch's block
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T        *tab   = reinterpret_cast<T *>(smem_raw);
  V        *xch   = reinterpret_cast<V *>(tab + nst * VW * stage_elems<T, ROWS, n>(F)); // [warp][2][XSLOT]
  // gather-ahead staging (gah): [n^3][TPB] source values of this CTA's NEXT batch, one private column per lane
  V        *gsm   = xch + (TPB / 32) * 2 * XSLOT + (GLSB_Q2_GAH ? threadIdx.x : 0);
  uint32_t *ibuf  = reinterpret_cast<uint32_t *>(xch + (TPB / 32) * 2 * XSLOT + (gah ? N3 * TPB : 0)); // [2][VW][4 n^3 + 1][32]
  V        *tsm   = reinterpret_cast<V *>(ibuf) + threadIdx.x;                        // TSM: [n^3][TPB] instead of ibuf
  uint64_t *full  = TSM ? reinterpret_cast<uint64_t *>(reinterpret_cast<V *>(ibuf) + N3 * TPB) :
                          reinterpret_cast<uint64_t *>(ibuf + 2 * ISL);
  uint64_t *ifull = full + MAX_NST;
  uint32_t *cnt   = reinterpret_cast<uint32_t *>(ifull + 2); // [MAX_NST + 2] release counters (tables, indices)

  const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int      c = lane & 3, col = 8 * warp + (lane >> 2);
  const bool     is_p = (c == 3);
  const int      cv   = is_p ? 0 : c; // table row used by this lane (pressure lane: any valid row)
  // exchange rows: the pairs (components 0/1 and 2/3) of this lane's cell, and the element this lane owns
  const int      xp0 = 2 * (lane >> 2), xp1 = 2 * (8 + (((lane >> 2) + 4) & 7));
  const int      xpos = ((c >> 1) ? xp1 : xp0) + (c & 1);
  V             *xw   = xch + warp * 2 * XSLOT;
#if GLSB_Q2_XCH
  // block layout of the exchange (per slot, 160 of the XSLOT values): entry e(k, j) = 4 k + ((j + (k >> 1)) & 3) of
  // cell k, component j holds (value, d_0, d_1, d_2) -- 16-byte quads for 4-byte values, two planes of 16-byte
  // pairs for 8-byte values.  The rotation makes the stores of a quarter-warp AND the broadcast reads of one
  // component by the 8 cells of a warp hit every bank once; SUPG residuals of the velocity rows follow at 128.
  constexpr bool XW8 = sizeof(V) == 8;
  const int      xk = lane >> 2, xrot = xk >> 1;
  const int      xe_own = 4 * xk + ((c + xrot) & 3);
#endif
  // column of this lane's cell in a 32-cell table row, for fields of component row 0, 1, 2 and cv
  const int      colr[4] = {col, (col + 4) & 31, (col + 8) & 31, (col + 4 * cv) & 31};
  // this lane's entries of an index block: dof (c, j) at ixo + j * CELLS, the cell's flag word at flo
  const int      ixo = (c * N3) * CELLS + ((col + 8 * c) & 31), flo = 4 * N3 * CELLS + col;

  if (threadIdx.x == 0)
    {
      for (int s = 0; s < nst; ++s)
        mbar_init(&full[s], 1);
      for (int s = 0; s < 2; ++s)
        mbar_init(&ifull[s], 1);
      for (int s = 0; s < MAX_NST + 2; ++s)
        cnt[s] = 0;
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  __syncthreads();

  const uint32_t n_batches = ((p.cell_end - p.cell_begin + CELLS - 1) / CELLS + VW - 1) / VW; // units
  const uint32_t my_n      = (blockIdx.x < n_batches) ? (n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t n_stages  = my_n * (N2 / ROWS);
  if (my_n == 0)
    return;

  // dof indices of batch bi of this CTA -> index ring slot bi & 1 (one bulk copy of 13.6 KB); one thread
  auto issue_idx = [&](uint32_t bi) {
    const uint32_t b     = bi & 1;
    const uint64_t batch = (uint64_t)(p.cell_begin >> 5) + (uint64_t)(blockIdx.x + bi * gridDim.x) * VW;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&ifull[b], ISL * 4);
    bulk_g2s(ibuf + b * ISL, p.idx + batch * IDX_ELEMS, ISL * 4, &ifull[b]); // consecutive batches are contiguous
  };
  // prologue: indices of the first two batches, all ring slots
  if (threadIdx.x == 0)
    {
      if (!TSM)
        {
          issue_idx(0);
          if (my_n > 1)
            issue_idx(1);
        }
      for (uint32_t j = 0; j < (uint32_t)nst && j < n_stages; ++j)
        issue_stage<T, ROWS, VW, n>(p, F, tab, full, j, j);
    }

  // gather-ahead: the source values of batch bi of this CTA go global -> shared with per-thread cp.async
  // (LDGSTS, no registers held while they are in flight) during the arithmetic of batch bi - 1; every lane
  // copies and later reads only its own column, so no barrier is involved.  Batches with constrained dofs
  // (warp-uniform flag) are gathered directly as before.
  auto gather_ahead = [&](uint32_t bi) -> void {
    const uint32_t *ib = ibuf + (bi & 1) * ISL;
    mbar_wait(&ifull[bi & 1], (bi >> 1) & 1);
    if (__any_sync(0xffffffffu, ib[flo] != 0 || (VW == 2 && ib[IDX_ELEMS + flo] != 0)))
      return;
    const uint32_t *ix = ib + ixo;
#pragma unroll
    for (int j = 0; j < N3; ++j)
      {
        if (VW == 1)
          cp_async<(int)sizeof(T)>(gsm + j * TPB, p.src + ix[j * CELLS]);
        else
          {
            T *g2 = reinterpret_cast<T *>(gsm + j * TPB);
            cp_async<(int)sizeof(T)>(g2, p.src + ix[j * CELLS]);
            cp_async<(int)sizeof(T)>(g2 + 1, p.src + ix[j * CELLS + IDX_ELEMS]);
          }
      }
    cp_async_commit();
  };
  if (!TSM && gah)
    gather_ahead(0);

  const V  w = VO::bcast(p.weight), nu = VO::bcast(p.nu);
  uint32_t it = 0, slot = 0, par = 0; // stage counter, its ring slot and phase parity
  for (uint32_t bi = 0; bi < my_n; ++bi)
    {
      const uint32_t unit   = blockIdx.x + bi * gridDim.x;
      // this lane's cell(s): position `col` of batch unit * VW (and of the next batch for the packed kernel)
      const uint32_t cellrA = p.cell_begin + (unit * VW) * CELLS + col, cellrB = cellrA + CELLS;
      const bool     actA = cell_active(p, cellrA), actB = (VW == 2) && cell_active(p, cellrB);
      const uint32_t cellA = cellrA < p.cell_end ? cellrA : p.cell_end - 1;
      const uint32_t cellB = cellrB < p.cell_end ? cellrB : p.cell_end - 1;
      // this unit's index block(s) (dof indices + one flag word per cell); cells with constrained dofs are
      // rare: one warp-uniform test instead of one per dof
      const uint32_t *iblk = TSM ? p.idx + ((uint64_t)(p.cell_begin >> 5) + unit) * IDX_ELEMS : ibuf + (bi & 1) * ISL;
      const uint32_t *ixs  = iblk + ixo;
      if (!TSM)
        mbar_wait(&ifull[bi & 1], (bi >> 1) & 1);
      const bool slow = __any_sync(0xffffffffu, iblk[flo] != 0 || (VW == 2 && iblk[IDX_ELEMS + flo] != 0));

      // ---- gather (read_dof_values) ------------------------------------------------------
      V t[N3];
      if (!slow && !TSM && gah)
        {
          cp_async_wait_all();
#pragma unroll
          for (int j = 0; j < N3; ++j)
            t[j] = gsm[j * TPB];
        }
      else if (!slow)
        {
#pragma unroll
          for (int j = 0; j < N3; ++j)
            {
              if (VW == 1)
                t[j] = VO::load(p.src, ixs[j * CELLS], 0);
              else
                t[j] = VO::load(p.src, ixs[j * CELLS], (int)ixs[j * CELLS + IDX_ELEMS] - (int)ixs[j * CELLS]);
            }
        }
      else
        {
#pragma unroll
          for (int j = 0; j < N3; ++j)
            {
              const T a = gather_resolved(p, p.src, ixs[j * CELLS]);
              if (VW == 1)
                t[j] = VO::load(&a, 0, 0);
              else
                {
                  const T ab[2] = {a, gather_resolved(p, p.src, ixs[j * CELLS + IDX_ELEMS])};
                  t[j]          = VO::load(ab, 0, 1);
                }
            }
        }
      const int cAB = (int)cellB - (int)cellA;
      V ij0 = VO::zero(), ij1 = VO::zero(), ij2 = VO::zero(), cdet = VO::zero();
      if (!GENERAL)
        {
          ij0  = VO::load(p.inv_jac, cellA, cAB);
          ij1  = VO::load(p.inv_jac + p.ncp, cellA, cAB);
          ij2  = VO::load(p.inv_jac + 2 * p.ncp, cellA, cAB);
          cdet = VO::load(p.jxw, cellA, cAB);
        }
      V d1c = VO::zero(), d2c = VO::zero();
      if (CELLWISE)
        {
          d1c = VO::load(p.d1c, cellA, cAB);
          d2c = VO::load(p.d2c, cellA, cAB);
        }

      // ---- interpolate to the quadrature points in x and y (registers only) --------------
#pragma unroll
      for (int l = 0; l < N2; ++l)
        {
          V in[n];
#pragma unroll
          for (int i = 0; i < n; ++i)
            in[i] = t[n * l + i];
#pragma unroll
          for (int q = 0; q < n; ++q)
            {
              V s = sh.S[q * n] * in[0];
#pragma unroll
              for (int i = 1; i < n; ++i)
                s += sh.S[q * n + i] * in[i];
              t[n * l + q] = s;
            }
        }
#pragma unroll
      for (int k = 0; k < n; ++k)
#pragma unroll
        for (int i = 0; i < n; ++i)
          {
            V in[n];
#pragma unroll
            for (int j = 0; j < n; ++j)
              in[j] = t[i + n * j + N2 * k];
#pragma unroll
            for (int q = 0; q < n; ++q)
              {
                V s = sh.S[q * n] * in[0];
#pragma unroll
                for (int j = 1; j < n; ++j)
                  s += sh.S[q * n + j] * in[j];
                t[i + n * q + N2 * k] = s;
              }
          }

      if (TSM)
        {
#pragma unroll
          for (int j = 0; j < N3; ++j)
            tsm[j * TPB] = t[j];
        }
#define GLSB_T(j) (TSM ? tsm[(j)*TPB] : t[(j)])
      V acc[N3];
#pragma unroll
      for (int j = 0; j < N3; ++j)
        acc[j] = VO::zero();

      // ---- quadrature layers --------------------------------------------------------------
#pragma unroll 1
      for (int qz = 0; qz < n; ++qz)
        {
          // runtime (CTA-uniform) layer index: select from the constant bank instead of indexing it
          V sz[n], gz[n], tz[n], hz[n];
          // test side: Cartesian cells carry the quadrature weights in the sweep matrices (Shape::Sw/Gw/Dt),
          // general cells get them with JxW from the table
          const V wz = selq<n, V>(qz, sh.w, 0, 1);
#pragma unroll
          for (int i = 0; i < n; ++i)
            {
              sz[i] = selq<n, V>(qz, sh.S, i, n);
              gz[i] = selq<n, V>(qz, sh.G, i, n);
              tz[i] = GENERAL ? sz[i] : sz[i] * wz;
              hz[i] = GENERAL ? gz[i] : gz[i] * wz;
            }
          V       vl[N2], wl[N2];
#pragma unroll
          for (int a = 0; a < N2; ++a)
            {
              V s = sz[0] * GLSB_T(a);
#pragma unroll
              for (int i = 1; i < n; ++i)
                s += sz[i] * GLSB_T(a + N2 * i);
              vl[a] = s;
              wl[a] = VO::zero();
            }
          const T *tb = nullptr; // first batch's block of the stage; the second one is SB elements further
#pragma unroll
          for (int qy = 0; qy < n; ++qy)
            {
              if (qy % ROWS == 0)
                {
                  mbar_wait(&full[slot], par);
                  tb = tab + (size_t)slot * VW * SB;
                }
              constexpr int QPS = n * ROWS;
              const int     qlo = (qy % ROWS) * n; // first point of this row inside the stage
#define GLSB_TAB(f, x) VO::load(tb, ((f)*QPS + qlo + (x)) * CELLS + col, SB)          /* fields without a component row */
#define GLSB_TABR(f, x, r) VO::load(tb, ((f)*QPS + qlo + (x)) * CELLS + colr[r], SB) /* rotated by 4 * row, see qoff() */
#pragma unroll
              for (int qx = 0; qx < n; ++qx)
                {
                  const int a   = n * qy + qx;
                  const V   val = vl[a];
                  V rx = sh.D[qx * n] * vl[n * qy], ry = sh.D[qy * n] * vl[qx], rz = gz[0] * GLSB_T(a);
#pragma unroll
                  for (int i = 1; i < n; ++i)
                    {
                      rx += sh.D[qx * n + i] * vl[n * qy + i];
                      ry += sh.D[qy * n + i] * vl[qx + n * i];
                      rz += gz[i] * GLSB_T(a + N2 * i);
                    }
                  // geometry: physical gradient of this lane's component
                  V g0, g1, g2, jq = VO::zero();
                  V J00, J01, J02, J10, J11, J12, J20, J21, J22;
                  if (GENERAL)
                    {
                      J00 = GLSB_TAB(p.fJ + 0, qx), J01 = GLSB_TAB(p.fJ + 1, qx), J02 = GLSB_TAB(p.fJ + 2, qx);
                      J10 = GLSB_TAB(p.fJ + 3, qx), J11 = GLSB_TAB(p.fJ + 4, qx), J12 = GLSB_TAB(p.fJ + 5, qx);
                      J20 = GLSB_TAB(p.fJ + 6, qx), J21 = GLSB_TAB(p.fJ + 7, qx), J22 = GLSB_TAB(p.fJ + 8, qx);
                      jq  = GLSB_TAB(p.fjxw, qx);
                      g0  = J00 * rx + J10 * ry + J20 * rz;
                      g1  = J01 * rx + J11 * ry + J21 * rz;
                      g2  = J02 * rx + J12 * ry + J22 * rz;
                    }
                  else
                    {
                      g0 = rx * ij0, g1 = ry * ij1, g2 = rz * ij2;
                    }
                  // ---- exchange round 1: publish value and gradient of this component --------------
                  V *xs = xw + (a & 1) * XSLOT;
#if GLSB_Q2_XCH
                  if (!XW8)
                    *reinterpret_cast<Quad<V> *>(xs + 4 * xe_own) = Quad<V>{val, g0, g1, g2};
                  else
                    {
                      *reinterpret_cast<Pair<V> *>(xs + 2 * xe_own)      = Pair<V>{val, g0};
                      *reinterpret_cast<Pair<V> *>(xs + 64 + 2 * xe_own) = Pair<V>{g1, g2};
                    }
#else
                  xs[xpos]            = val;
                  xs[XROW + xpos]     = g0;
                  xs[2 * XROW + xpos] = g1;
                  xs[3 * XROW + xpos] = g2;
#endif
                  // tables (field offsets of the prefix are fixed: U 0..2, grad U 3..11, grad P 12..14)
                  const V U0 = GLSB_TABR(0, qx, 0), U1 = GLSB_TABR(1, qx, 1), U2 = GLSB_TABR(2, qx, 2);
                  const V H0 = GLSB_TABR(3 + 3 * cv, qx, 3), H1 = GLSB_TABR(4 + 3 * cv, qx, 3),
                          H2 = GLSB_TABR(5 + 3 * cv, qx, 3);
                  const V Pc = GLSB_TABR(12 + cv, qx, 3);
                  const V d1 = CELLWISE ? d1c : GLSB_TAB(p.fd1q, qx);
                  const V d2 = CELLWISE ? d2c : GLSB_TAB(p.fd2q, qx);
                  __syncwarp();
#if GLSB_Q2_XCH
                  // the 4 lanes of a cell read the whole 4 x 4 block (broadcast reads), then pick their column
                  V qv[4], q0[4], q1[4], q2[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    {
                      const int e = 4 * xk + ((j + xrot) & 3);
                      if (!XW8)
                        {
                          const Quad<V> q = *reinterpret_cast<const Quad<V> *>(xs + 4 * e);
                          qv[j] = q.a, q0[j] = q.b, q1[j] = q.c, q2[j] = q.d;
                        }
                      else
                        {
                          const Pair<V> pa = *reinterpret_cast<const Pair<V> *>(xs + 2 * e),
                                        pb = *reinterpret_cast<const Pair<V> *>(xs + 64 + 2 * e);
                          qv[j] = pa.a, q0[j] = pa.b, q1[j] = pb.a, q2[j] = pb.b;
                        }
                    }
                  const V u0 = qv[0], u1 = qv[1], u2 = qv[2], pp = qv[3];
                  const V div = q0[0] + q1[1] + q2[2];
                  // column c of grad u and d_c p
                  const V Gc0 = sel3(cv, q0[0], q1[0], q2[0]), Gc1 = sel3(cv, q0[1], q1[1], q2[1]),
                          Gc2 = sel3(cv, q0[2], q1[2], q2[2]), gpc = sel3(cv, q0[3], q1[3], q2[3]);
#else
                  const Pair<V> u01 = *reinterpret_cast<const Pair<V> *>(xs + xp0),
                                u2p = *reinterpret_cast<const Pair<V> *>(xs + xp1);
                  const V u0 = u01.a, u1 = u01.b, u2 = u2p.a, pp = u2p.b;
                  const V div = xs[XROW + xp0] + xs[2 * XROW + xp0 + 1] + xs[3 * XROW + xp1];
                  // column c of grad u and d_c p: row (1 + c), the 4 elements of this cell
                  const V      *xc  = xs + (1 + cv) * XROW;
                  const Pair<V> G01 = *reinterpret_cast<const Pair<V> *>(xc + xp0),
                                G2p = *reinterpret_cast<const Pair<V> *>(xc + xp1);
                  const V Gc0 = G01.a, Gc1 = G01.b, Gc2 = G2p.a, gpc = G2p.b;
#endif
                  const V  td  = val * w;
                  const V  sgu = g0 * U0 + g1 * U1 + g2 * U2; // U . grad u_c
                  const V  ugs = H0 * u0 + H1 * u1 + H2 * u2; // u . grad U_c
                  V        y   = sgu + ugs;
                  if (CTD)
                    y = td + y;
                  // velocity row c: SUPG residual of the increment, delta_1 (y + d_c p)
                  const V r0  = d1 * (y + gpc);
                  // ---- exchange round 2: the pressure row is (grad q, residual_0): it needs r0 of the three
                  // velocity rows (operator_ns.cc:1166-1172), published as they are
#if GLSB_Q2_XCH
                  if (!XW8)
                    xs[128 + 4 * xk + c] = r0;
                  else
                    xs[128 + (c >> 1) * 16 + 2 * xk + (c & 1)] = r0;
#else
                  xs[4 * XROW + xpos] = r0;
#endif
                  const V sgs = H0 * U0 + H1 * U1 + H2 * U2; // U . grad U_c
                  V       rb  = Pc + sgs;
                  if (CTD)
                    rb = (GLSB_TABR(cv, qx, 3) * w + GLSB_TABR(p.fO + cv, qx, 3)) + rb;
                  const V rr1  = d1 * rb;
                  const V diag = d2 * div - pp;
                  V       vo   = td + sgu + ugs;
                  V       o0   = nu * (g0 + Gc0) + U0 * r0 + u0 * rr1 + vsel(c == 0, diag, VO::zero());
                  V       o1   = nu * (g1 + Gc1) + U1 * r0 + u1 * rr1 + vsel(c == 1, diag, VO::zero());
                  V       o2   = nu * (g2 + Gc2) + U2 * r0 + u2 * rr1 + vsel(c == 2, diag, VO::zero());
                  __syncwarp();
                  // pressure row: (q, div u) and (grad q, residual_0)
#if GLSB_Q2_XCH
                  V rp0, rp1, rp2;
                  if (!XW8)
                    {
                      const Quad<V> r = *reinterpret_cast<const Quad<V> *>(xs + 128 + 4 * xk);
                      rp0 = r.a, rp1 = r.b, rp2 = r.c;
                    }
                  else
                    {
                      const Pair<V> r = *reinterpret_cast<const Pair<V> *>(xs + 128 + 2 * xk);
                      rp0 = r.a, rp1 = r.b, rp2 = xs[144 + 2 * xk];
                    }
                  vo = vsel(is_p, div, vo);
                  o0 = vsel(is_p, rp0, o0);
                  o1 = vsel(is_p, rp1, o1);
                  o2 = vsel(is_p, rp2, o2);
#else
                  const Pair<V> r01 = *reinterpret_cast<const Pair<V> *>(xs + 4 * XROW + xp0);
                  vo = vsel(is_p, div, vo);
                  o0 = vsel(is_p, r01.a, o0);
                  o1 = vsel(is_p, r01.b, o1);
                  o2 = vsel(is_p, xs[4 * XROW + xp1], o2);
#endif
                  // submit_value / submit_gradient: times JxW, back to the reference cell
                  V ox, oy, oz;
                  if (GENERAL)
                    {
                      // at the 255-register limit (FP64): JxW and one row of J^-1 are read again here
                      // instead of being kept in 8 registers across the physics
                      constexpr bool RELOAD = sizeof(V) == 8 && VW == 1;
                      const V        jw = RELOAD ? GLSB_TAB(p.fjxw, qx) : jq;
                      vo *= jw;
                      if (RELOAD)
                        ox = (GLSB_TAB(p.fJ + 0, qx) * o0 + GLSB_TAB(p.fJ + 1, qx) * o1 + GLSB_TAB(p.fJ + 2, qx) * o2) * jw;
                      else
                        ox = (J00 * o0 + J01 * o1 + J02 * o2) * jw;
                      oy = (J10 * o0 + J11 * o1 + J12 * o2) * jw;
                      oz = (J20 * o0 + J21 * o1 + J22 * o2) * jw;
                    }
                  else
                    {
                      // weights live in Dt / Sw / Gw, det J is applied once per dof before the scatter
                      ox = o0 * ij0, oy = o1 * ij1, oz = o2 * ij2;
                    }
                  // integrate: collocation derivative transposed in x and y inside the layer, z into acc
                  wl[a] += vo;
#pragma unroll
                  for (int i = 0; i < n; ++i)
                    {
                      wl[i + n * qy] += (GENERAL ? sh.D[qx * n + i] : sh.Dt[qx * n + i]) * ox;
                      wl[qx + n * i] += (GENERAL ? sh.D[qy * n + i] : sh.Dt[qy * n + i]) * oy;
                    }
#pragma unroll
                  for (int i = 0; i < n; ++i)
                    acc[a + N2 * i] += hz[i] * oz;
                }
#undef GLSB_TAB
#undef GLSB_TABR
              // release the ring slot; the last warp to do so refills it with the stage nst stages ahead
              if ((qy + 1) % ROWS == 0)
                {
                  __syncwarp();
                  if (lane == 0)
                    {
                      // Every value this warp read from the slot has been CONSUMED by arithmetic above (the loads
                      // have returned), and __syncwarp orders the lanes; the refilling thread issues
                      // fence.proxy.async before the bulk copy.  An explicit __threadfence_block() pair around this
                      // counter was measured: the MEMBAR waits for the lane's outstanding global atomics and loads,
                      // 31.7 -> 23.5 GDoF/s (profiles/r02_experiments.txt) -- not kept.
                      if (atomicAdd(&cnt[slot], 1u) == TPB / 32 - 1)
                        {
                          cnt[slot] = 0;
                          if (it + nst < n_stages)
                            issue_stage<T, ROWS, VW, n>(p, F, tab, full, it + nst, slot);
                        }
                    }
                  ++it;
                  if (++slot == (uint32_t)nst)
                    {
                      slot = 0;
                      par ^= 1;
                    }
                }
            }
#pragma unroll
          for (int a = 0; a < N2; ++a)
#pragma unroll
            for (int i = 0; i < n; ++i)
              acc[a + N2 * i] += tz[i] * wl[a];
          // the next batch's index block (requested at the end of the previous batch) has had a layer's time
          // to arrive: start its gather now, it completes behind the remaining layers
          if (!TSM && gah && qz == 0 && bi + 1 < my_n)
            gather_ahead(bi + 1);
        }

      // ---- test with the basis in y and x (transposed sweeps) -----------------------------
#pragma unroll
      for (int k = 0; k < n; ++k)
#pragma unroll
        for (int i = 0; i < n; ++i)
          {
            V in[n];
#pragma unroll
            for (int j = 0; j < n; ++j)
              in[j] = acc[i + n * j + N2 * k];
#pragma unroll
            for (int q = 0; q < n; ++q)
              {
                V s = (GENERAL ? sh.S[q] : sh.Sw[q]) * in[0];
#pragma unroll
                for (int j = 1; j < n; ++j)
                  s += (GENERAL ? sh.S[n * j + q] : sh.Sw[n * j + q]) * in[j];
                acc[i + n * q + N2 * k] = s;
              }
          }
#pragma unroll
      for (int l = 0; l < N2; ++l)
        {
          V in[n];
#pragma unroll
          for (int i = 0; i < n; ++i)
            in[i] = acc[n * l + i];
#pragma unroll
          for (int q = 0; q < n; ++q)
            {
              V s = (GENERAL ? sh.S[q] : sh.Sw[q]) * in[0];
#pragma unroll
              for (int i = 1; i < n; ++i)
                s += (GENERAL ? sh.S[n * i + q] : sh.Sw[n * i + q]) * in[i];
              acc[n * l + q] = s;
            }
        }
      if (!GENERAL)
        {
#pragma unroll
          for (int j = 0; j < N3; ++j)
            acc[j] *= cdet;
        }

      // ---- scatter (distribute_local_to_global) ------------------------------------------
#pragma unroll
      for (int v = 0; v < VW; ++v)
        if (v == 0 ? actA : actB)
          {
            const uint32_t *ix = ixs + v * IDX_ELEMS;
            T               r[N3];
#pragma unroll
            for (int j = 0; j < N3; ++j)
              r[j] = lane_value<T>(acc[j], v);
            if (!slow)
              {
#pragma unroll
                for (int j = 0; j < N3; ++j)
                  atomic_add(p.dst + ix[j * CELLS], r[j]);
              }
            else
              {
#pragma unroll
                for (int j = 0; j < N3; ++j)
                  scatter_resolved(p, p.dst, ix[j * CELLS], r[j]);
              }
          }
#undef GLSB_T
      // release the index ring slot; the last warp refills it with the block two batches ahead
      __syncwarp();
      if (!TSM && lane == 0)
        {
          if (atomicAdd(&cnt[MAX_NST + (bi & 1)], 1u) == TPB / 32 - 1) // see the table ring above
            {
              cnt[MAX_NST + (bi & 1)] = 0;
              if (bi + 2 < my_n)
                issue_idx(bi + 2);
            }
        }
    }
}

// gather-ahead staging of the source values (GLSB_Q2_GAHEAD=0/1 overrides the default)
template <typename T, int n>
inline bool use_gather_ahead()
{
  static const int env = getenv("GLSB_Q2_GAHEAD") ? atoi(getenv("GLSB_Q2_GAHEAD")) : -1;
  if (use_tsm<T, n>() || !GLSB_Q2_GAH)
    return false;
  return env >= 0 ? env != 0 : false;
}

template <typename T, int ROWS, int VW, int n>
size_t smem_bytes(int F, int nst)
{
  const size_t idx_or_t = use_tsm<T, n>() ? (size_t)n * n * n * TPB * sizeof(T) : (size_t)2 * VW * idx_elems<n>() * 4;
  const size_t gsm      = use_gather_ahead<T, n>() ? (size_t)n * n * n * TPB * (VW * sizeof(T)) : 0;
  return nst * VW * stage_elems<T, ROWS, n>(F) * sizeof(T) + (size_t)(TPB / 32) * 2 * XSLOT * (VW * sizeof(T)) + gsm +
         idx_or_t + (MAX_NST + 2) * (sizeof(uint64_t) + sizeof(uint32_t));
}

// ring depth: 2 stages of a whole layer; with row stages as deep as the target occupancy allows (<= 4)
template <typename T, int ROWS, int VW, int n>
int ring_depth(int F)
{
  const int ctas = target_ctas<(int)(VW * sizeof(T)), n>();
  int       nst  = 2;
  while (ROWS == 1 && nst < 4 && ctas * (smem_bytes<T, ROWS, VW, n>(F, nst + 1) + 1024) <= 228 * 1024)
    ++nst;
  return nst;
}

template <typename T, typename V, bool GENERAL, bool CTD, bool CELLWISE, int ROWS, int n>
static int launch_rows(const KParams<T> &p, const Shape<V, n> &S, int F, cudaStream_t s)
{
  constexpr int VW = VOps<V, T>::VW;
  static int   n_sm = 0;
  static const int env_nst = getenv("GLSB_Q2_NST") ? atoi(getenv("GLSB_Q2_NST")) : 0;
  const int    nst  = env_nst >= 2 && env_nst <= MAX_NST ? env_nst : ring_depth<T, ROWS, VW, n>(F);
  const size_t smem = smem_bytes<T, ROWS, VW, n>(F, nst);
  auto         kern = k_vmult_q2_newton<T, V, GENERAL, CTD, CELLWISE, ROWS, n>;
  if (smem > 227 * 1024)
    return -1;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return 1;
  int bps = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, TPB, smem) != cudaSuccess || bps < 1)
    return -1;
  if (n_sm == 0)
    {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
  const uint32_t n_units = ((p.cell_end - p.cell_begin + CELLS - 1) / CELLS + VW - 1) / VW;
  const int      use_sm  = (p.sm_reserve > 0 && p.sm_reserve < n_sm) ? n_sm - p.sm_reserve : n_sm;
  const uint32_t grid    = n_units < (uint32_t)(use_sm * bps) ? n_units : (uint32_t)(use_sm * bps);
  kern<<<grid, TPB, smem, s>>>(p, S, F, nst, use_gather_ahead<T, n>() ? 1 : 0);
  return cudaGetLastError() != cudaSuccess;
}

template <typename T, typename V, bool GENERAL, bool CTD, bool CELLWISE, int n>
static int launch(const KParams<T> &p, const Shape<V, n> &S, int F, cudaStream_t s)
{
  if (p.QG == n * n)
    return launch_rows<T, V, GENERAL, CTD, CELLWISE, n, n>(p, S, F, s);
  return launch_rows<T, V, GENERAL, CTD, CELLWISE, 1, n>(p, S, F, s);
}

template <typename T, typename V, bool GENERAL, int n = 3>
static int launch_flags(const KParams<T> &p, const Shape<V, n> &S, int F, cudaStream_t s)
{
  if (p.ctd)
    return p.cell_wise ? launch<T, V, GENERAL, true, true, n>(p, S, F, s) :
                         launch<T, V, GENERAL, true, false, n>(p, S, F, s);
  return p.cell_wise ? launch<T, V, GENERAL, false, true, n>(p, S, F, s) :
                       launch<T, V, GENERAL, false, false, n>(p, S, F, s);
}

} // namespace q2
} // namespace glsb
