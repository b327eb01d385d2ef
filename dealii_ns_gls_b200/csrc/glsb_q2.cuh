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

namespace glsb
{
namespace q2
{
#ifndef GLSB_Q2_F64_CTAS
#define GLSB_Q2_F64_CTAS 2
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

template <typename T>
__device__ __forceinline__ T sel3(int c, T a0, T a1, T a2)
{
  return c == 0 ? a0 : (c == 1 ? a1 : a2);
}

// a ring stage holds ROWS rows (qz, qy) of 3 quadrature points: ROWS = 1 (9 stages per batch) or 3 (one
// stage per quadrature layer); the q-point array is laid out accordingly (KParams::NL = 9 / ROWS, QG = 3 ROWS)
template <typename T, int ROWS>
__host__ __device__ constexpr size_t stage_elems(int F)
{
  return (size_t)F * 3 * ROWS * CELLS;
}

constexpr int IDX_ROWS  = 109;              // 108 dof indices + one flag word per cell
constexpr int IDX_ELEMS = IDX_ROWS * CELLS; // index block of one batch
constexpr int MAX_NST   = 8;
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

// issue the bulk copy of stage j (row j % 9 of quadrature points of this CTA's batch j / 9) into ring slot
// `slot`; called by one thread.  There is no producer warp: the slot is refilled by whichever warp is the
// last to release it (an arrival counter per slot), so no warp ever waits for another one's progress.
template <typename T, int ROWS>
__device__ __forceinline__ void issue_stage(const KParams<T> &p, int F, T *tab, uint64_t *full, uint32_t j,
                                            uint32_t slot)
{
  constexpr uint32_t SPB   = 9 / ROWS; // stages per batch
  const uint32_t     bytes = (uint32_t)(stage_elems<T, ROWS>(F) * sizeof(T));
  const uint32_t     batch = blockIdx.x + (j / SPB) * gridDim.x, row = j % SPB;
  const T *src = p.Q + (((uint64_t)((p.cell_begin >> 5) + batch) * SPB + row) * p.FT) * (3 * ROWS * CELLS);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(&full[slot], bytes);
  bulk_g2s(tab + (size_t)slot * stage_elems<T, ROWS>(F), src, bytes, &full[slot]);
}

template <typename T, bool GENERAL, bool CTD, bool CELLWISE, int ROWS>
__global__ void __launch_bounds__(TPB, (sizeof(T) == 4 ? GLSB_Q2_F32_CTAS : GLSB_Q2_F64_CTAS))
  k_vmult_q2_newton(const KParams<T> p, const Shape<T, 3> sh, const int F, const int nst)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T        *tab   = reinterpret_cast<T *>(smem_raw);
  T        *xch   = tab + nst * stage_elems<T, ROWS>(F);              // [warp][2][XSLOT]
  uint32_t *ibuf  = reinterpret_cast<uint32_t *>(xch + (TPB / 32) * 2 * XSLOT); // [2][109][32] dof indices, flags
  uint64_t *full  = reinterpret_cast<uint64_t *>(ibuf + 2 * IDX_ELEMS);
  uint64_t *ifull = full + MAX_NST;
  uint32_t *cnt   = reinterpret_cast<uint32_t *>(ifull + 2); // [MAX_NST + 2] release counters (tables, indices)

  const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int      c = lane & 3, col = 8 * warp + (lane >> 2);
  const bool     is_p = (c == 3);
  const int      cv   = is_p ? 0 : c; // table row used by this lane (pressure lane: any valid row)
  // exchange rows: the pairs (components 0/1 and 2/3) of this lane's cell, and the element this lane owns
  const int      xp0 = 2 * (lane >> 2), xp1 = 2 * (8 + (((lane >> 2) + 4) & 7));
  const int      xpos = ((c >> 1) ? xp1 : xp0) + (c & 1);
  T             *xw   = xch + warp * 2 * XSLOT;
  // column of this lane's cell in a 32-cell table row, for fields of component row 0, 1, 2 and cv
  const int      colr[4] = {col, (col + 4) & 31, (col + 8) & 31, (col + 4 * cv) & 31};
  // this lane's entries of an index block: dof (c, j) at ixo + j * CELLS, the cell's flag word at flo
  const int      ixo = (c * 27) * CELLS + ((col + 8 * c) & 31), flo = 108 * CELLS + col;

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

  const uint32_t n_batches = (p.cell_end - p.cell_begin + CELLS - 1) / CELLS;
  const uint32_t my_n      = (blockIdx.x < n_batches) ? (n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t n_stages  = my_n * (9 / ROWS);
  if (my_n == 0)
    return;

  // dof indices of batch bi of this CTA -> index ring slot bi & 1 (one bulk copy of 13.6 KB); one thread
  auto issue_idx = [&](uint32_t bi) {
    const uint32_t b     = bi & 1;
    const uint32_t batch = (p.cell_begin >> 5) + blockIdx.x + bi * gridDim.x;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&ifull[b], IDX_ELEMS * 4);
    bulk_g2s(ibuf + b * IDX_ELEMS, p.idx + (uint64_t)batch * IDX_ELEMS, IDX_ELEMS * 4, &ifull[b]);
  };
  // prologue: indices of the first two batches, all ring slots
  if (threadIdx.x == 0)
    {
      issue_idx(0);
      if (my_n > 1)
        issue_idx(1);
      for (uint32_t j = 0; j < (uint32_t)nst && j < n_stages; ++j)
        issue_stage<T, ROWS>(p, F, tab, full, j, j);
    }

  const T  w = p.weight, nu = p.nu;
  uint32_t it = 0, slot = 0, par = 0; // stage counter, its ring slot and phase parity
  for (uint32_t bi = 0; bi < my_n; ++bi)
    {
      const uint32_t batch  = blockIdx.x + bi * gridDim.x;
      const uint32_t cell0  = p.cell_begin + batch * CELLS;
      const uint32_t cellr  = cell0 + col;
      const bool     active = cell_active(p, cellr);
      const uint32_t cell   = cellr < p.cell_end ? cellr : p.cell_end - 1;
      const uint32_t *ixs   = ibuf + (bi & 1) * IDX_ELEMS + ixo;
      // this batch's index block (dof indices + one flag word per cell); cells with constrained dofs are
      // rare: one warp-uniform test instead of one per dof
      mbar_wait(&ifull[bi & 1], (bi >> 1) & 1);
      const bool slow = __any_sync(0xffffffffu, ibuf[(bi & 1) * IDX_ELEMS + flo] != 0);

      // ---- gather (read_dof_values) ------------------------------------------------------
      T t[27];
      if (!slow)
        {
#pragma unroll
          for (int j = 0; j < 27; ++j)
            t[j] = p.src[ixs[j * CELLS]];
        }
      else
        {
#pragma unroll
          for (int j = 0; j < 27; ++j)
            t[j] = gather_resolved(p, p.src, ixs[j * CELLS]);
        }
      T ij0 = 0, ij1 = 0, ij2 = 0, cdet = 0;
      if (!GENERAL)
        {
          ij0  = p.inv_jac[cell];
          ij1  = p.inv_jac[p.ncp + cell];
          ij2  = p.inv_jac[2 * p.ncp + cell];
          cdet = p.jxw[cell];
        }
      T d1c = 0, d2c = 0;
      if (CELLWISE)
        {
          d1c = p.d1c[cell];
          d2c = p.d2c[cell];
        }

      // ---- interpolate to the quadrature points in x and y (registers only) --------------
#pragma unroll
      for (int l = 0; l < 9; ++l)
        {
          const T a = t[3 * l], b = t[3 * l + 1], d = t[3 * l + 2];
#pragma unroll
          for (int q = 0; q < 3; ++q)
            t[3 * l + q] = sh.S[q * 3] * a + sh.S[q * 3 + 1] * b + sh.S[q * 3 + 2] * d;
        }
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < 3; ++i)
          {
            const T a = t[i + 9 * k], b = t[i + 3 + 9 * k], d = t[i + 6 + 9 * k];
#pragma unroll
            for (int q = 0; q < 3; ++q)
              t[i + 3 * q + 9 * k] = sh.S[q * 3] * a + sh.S[q * 3 + 1] * b + sh.S[q * 3 + 2] * d;
          }

      T acc[27];
#pragma unroll
      for (int j = 0; j < 27; ++j)
        acc[j] = 0;

      // ---- quadrature layers --------------------------------------------------------------
#pragma unroll 1
      for (int qz = 0; qz < 3; ++qz)
        {
          // runtime (CTA-uniform) layer index: select from the constant bank instead of indexing it
          const T sz0 = sel3<T>(qz, sh.S[0], sh.S[3], sh.S[6]), sz1 = sel3<T>(qz, sh.S[1], sh.S[4], sh.S[7]),
                  sz2 = sel3<T>(qz, sh.S[2], sh.S[5], sh.S[8]);
          const T gz0 = sel3<T>(qz, sh.G[0], sh.G[3], sh.G[6]), gz1 = sel3<T>(qz, sh.G[1], sh.G[4], sh.G[7]),
                  gz2 = sel3<T>(qz, sh.G[2], sh.G[5], sh.G[8]);
          // test side: Cartesian cells carry the quadrature weights in the sweep matrices (Shape::Sw/Gw/Dt),
          // general cells get them with JxW from the table
          const T wz  = sel3<T>(qz, sh.w[0], sh.w[1], sh.w[2]);
          const T tz0 = GENERAL ? sz0 : sz0 * wz, tz1 = GENERAL ? sz1 : sz1 * wz, tz2 = GENERAL ? sz2 : sz2 * wz;
          const T hz0 = GENERAL ? gz0 : gz0 * wz, hz1 = GENERAL ? gz1 : gz1 * wz, hz2 = GENERAL ? gz2 : gz2 * wz;
          T       vl[9], wl[9];
#pragma unroll
          for (int a = 0; a < 9; ++a)
            {
              vl[a] = sz0 * t[a] + sz1 * t[a + 9] + sz2 * t[a + 18];
              wl[a] = 0;
            }
          const T *tb = nullptr;
#pragma unroll
          for (int qy = 0; qy < 3; ++qy)
            {
              if (qy % ROWS == 0)
                {
                  mbar_wait(&full[slot], par);
                  tb = tab + (size_t)slot * stage_elems<T, ROWS>(F);
                }
              constexpr int QPS = 3 * ROWS;
              const int     qlo = (qy % ROWS) * 3; // first point of this row inside the stage
#define GLSB_TAB(f, x) tb[((f)*QPS + qlo + (x)) * CELLS + col]          /* fields without a component row */
#define GLSB_TABR(f, x, r) tb[((f)*QPS + qlo + (x)) * CELLS + colr[r]] /* rotated by 4 * row, see qoff() */
#pragma unroll
              for (int qx = 0; qx < 3; ++qx)
                {
                  const int a   = 3 * qy + qx;
                  const T   val = vl[a];
                  const T rx = sh.D[qx * 3] * vl[3 * qy] + sh.D[qx * 3 + 1] * vl[3 * qy + 1] + sh.D[qx * 3 + 2] * vl[3 * qy + 2];
                  const T ry = sh.D[qy * 3] * vl[qx] + sh.D[qy * 3 + 1] * vl[qx + 3] + sh.D[qy * 3 + 2] * vl[qx + 6];
                  const T rz = gz0 * t[a] + gz1 * t[a + 9] + gz2 * t[a + 18];
                  // geometry: physical gradient of this lane's component
                  T g0, g1, g2, jq = 0;
                  T J00, J01, J02, J10, J11, J12, J20, J21, J22;
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
                  T *xs = xw + (a & 1) * XSLOT;
                  xs[xpos]            = val;
                  xs[XROW + xpos]     = g0;
                  xs[2 * XROW + xpos] = g1;
                  xs[3 * XROW + xpos] = g2;
                  // tables (field offsets of the prefix are fixed: U 0..2, grad U 3..11, grad P 12..14)
                  const T U0 = GLSB_TABR(0, qx, 0), U1 = GLSB_TABR(1, qx, 1), U2 = GLSB_TABR(2, qx, 2);
                  const T H0 = GLSB_TABR(3 + 3 * cv, qx, 3), H1 = GLSB_TABR(4 + 3 * cv, qx, 3),
                          H2 = GLSB_TABR(5 + 3 * cv, qx, 3);
                  const T Pc = GLSB_TABR(12 + cv, qx, 3);
                  const T d1 = CELLWISE ? d1c : GLSB_TAB(p.fd1q, qx);
                  const T d2 = CELLWISE ? d2c : GLSB_TAB(p.fd2q, qx);
                  __syncwarp();
                  const Pair<T> u01 = *reinterpret_cast<const Pair<T> *>(xs + xp0),
                                u2p = *reinterpret_cast<const Pair<T> *>(xs + xp1);
                  const T u0 = u01.a, u1 = u01.b, u2 = u2p.a, pp = u2p.b;
                  const T div = xs[XROW + xp0] + xs[2 * XROW + xp0 + 1] + xs[3 * XROW + xp1];
                  // column c of grad u and d_c p: row (1 + c), the 4 elements of this cell
                  const T      *xc  = xs + (1 + cv) * XROW;
                  const Pair<T> G01 = *reinterpret_cast<const Pair<T> *>(xc + xp0),
                                G2p = *reinterpret_cast<const Pair<T> *>(xc + xp1);
                  const T Gc0 = G01.a, Gc1 = G01.b, Gc2 = G2p.a, gpc = G2p.b;
                  const T  td  = val * w;
                  const T  sgu = g0 * U0 + g1 * U1 + g2 * U2; // U . grad u_c
                  const T  ugs = H0 * u0 + H1 * u1 + H2 * u2; // u . grad U_c
                  T        y   = sgu + ugs;
                  if (CTD)
                    y = td + y;
                  // velocity row c: SUPG residual of the increment, delta_1 (y + d_c p)
                  const T r0  = d1 * (y + gpc);
                  // ---- exchange round 2: the pressure row is (grad q, residual_0): it needs r0 of the three
                  // velocity rows (operator_ns.cc:1166-1172), published as they are
                  xs[4 * XROW + xpos] = r0;
                  const T sgs = H0 * U0 + H1 * U1 + H2 * U2; // U . grad U_c
                  T       rb  = Pc + sgs;
                  if (CTD)
                    rb = (GLSB_TABR(cv, qx, 3) * w + GLSB_TABR(p.fO + cv, qx, 3)) + rb;
                  const T rr1  = d1 * rb;
                  const T diag = d2 * div - pp;
                  T       vo   = td + sgu + ugs;
                  T       o0   = nu * (g0 + Gc0) + U0 * r0 + u0 * rr1 + (c == 0 ? diag : T(0));
                  T       o1   = nu * (g1 + Gc1) + U1 * r0 + u1 * rr1 + (c == 1 ? diag : T(0));
                  T       o2   = nu * (g2 + Gc2) + U2 * r0 + u2 * rr1 + (c == 2 ? diag : T(0));
                  __syncwarp();
                  // pressure row: (q, div u) and (grad q, residual_0)
                  const Pair<T> r01 = *reinterpret_cast<const Pair<T> *>(xs + 4 * XROW + xp0);
                  vo = is_p ? div : vo;
                  o0 = is_p ? r01.a : o0;
                  o1 = is_p ? r01.b : o1;
                  o2 = is_p ? xs[4 * XROW + xp1] : o2;
                  // submit_value / submit_gradient: times JxW, back to the reference cell
                  T ox, oy, oz;
                  if (GENERAL)
                    {
                      vo *= jq;
                      ox = (J00 * o0 + J01 * o1 + J02 * o2) * jq;
                      oy = (J10 * o0 + J11 * o1 + J12 * o2) * jq;
                      oz = (J20 * o0 + J21 * o1 + J22 * o2) * jq;
                    }
                  else
                    {
                      // weights live in Dt / Sw / Gw, det J is applied once per dof before the scatter
                      ox = o0 * ij0, oy = o1 * ij1, oz = o2 * ij2;
                    }
                  // integrate: collocation derivative transposed in x and y inside the layer, z into acc
                  wl[a] += vo;
#pragma unroll
                  for (int i = 0; i < 3; ++i)
                    {
                      wl[i + 3 * qy] += (GENERAL ? sh.D[qx * 3 + i] : sh.Dt[qx * 3 + i]) * ox;
                      wl[qx + 3 * i] += (GENERAL ? sh.D[qy * 3 + i] : sh.Dt[qy * 3 + i]) * oy;
                    }
                  acc[a] += hz0 * oz;
                  acc[a + 9] += hz1 * oz;
                  acc[a + 18] += hz2 * oz;
                }
#undef GLSB_TAB
#undef GLSB_TABR
              // release the ring slot; the last warp to do so refills it with the stage nst stages ahead
              if ((qy + 1) % ROWS == 0)
                {
                  __syncwarp();
                  if (lane == 0)
                    {
                      if (atomicAdd(&cnt[slot], 1u) == TPB / 32 - 1)
                        {
                          cnt[slot] = 0;
                          if (it + nst < n_stages)
                            issue_stage<T, ROWS>(p, F, tab, full, it + nst, slot);
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
          for (int a = 0; a < 9; ++a)
            {
              acc[a] += tz0 * wl[a];
              acc[a + 9] += tz1 * wl[a];
              acc[a + 18] += tz2 * wl[a];
            }
        }

      // ---- test with the basis in y and x (transposed sweeps) -----------------------------
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < 3; ++i)
          {
            const T a = acc[i + 9 * k], b = acc[i + 3 + 9 * k], d = acc[i + 6 + 9 * k];
#pragma unroll
            for (int q = 0; q < 3; ++q)
              acc[i + 3 * q + 9 * k] = GENERAL ? sh.S[q] * a + sh.S[3 + q] * b + sh.S[6 + q] * d :
                                                 sh.Sw[q] * a + sh.Sw[3 + q] * b + sh.Sw[6 + q] * d;
          }
#pragma unroll
      for (int l = 0; l < 9; ++l)
        {
          const T a = acc[3 * l], b = acc[3 * l + 1], d = acc[3 * l + 2];
#pragma unroll
          for (int q = 0; q < 3; ++q)
            acc[3 * l + q] = GENERAL ? sh.S[q] * a + sh.S[3 + q] * b + sh.S[6 + q] * d :
                                       sh.Sw[q] * a + sh.Sw[3 + q] * b + sh.Sw[6 + q] * d;
        }
      if (!GENERAL)
        {
#pragma unroll
          for (int j = 0; j < 27; ++j)
            acc[j] *= cdet;
        }

      // ---- scatter (distribute_local_to_global) ------------------------------------------
      if (active)
        {
          if (!slow)
            {
#pragma unroll
              for (int j = 0; j < 27; ++j)
                atomic_add(p.dst + ixs[j * CELLS], acc[j]);
            }
          else
            {
#pragma unroll
              for (int j = 0; j < 27; ++j)
                scatter_resolved(p, p.dst, ixs[j * CELLS], acc[j]);
            }
        }
      // release the index ring slot; the last warp refills it with the block two batches ahead
      __syncwarp();
      if (lane == 0)
        {
          if (atomicAdd(&cnt[MAX_NST + (bi & 1)], 1u) == TPB / 32 - 1)
            {
              cnt[MAX_NST + (bi & 1)] = 0;
              if (bi + 2 < my_n)
                issue_idx(bi + 2);
            }
        }
    }
}

template <typename T, int ROWS>
size_t smem_bytes(int F, int nst)
{
  return (nst * stage_elems<T, ROWS>(F) + (TPB / 32) * 2 * XSLOT) * sizeof(T) +
         2 * IDX_ELEMS * 4 + (MAX_NST + 2) * (sizeof(uint64_t) + sizeof(uint32_t));
}

// ring depth: 2 stages of a whole layer; with row stages as deep as the target occupancy allows (<= 4)
template <typename T, int ROWS>
int ring_depth(int F)
{
  const int ctas = sizeof(T) == 4 ? GLSB_Q2_F32_CTAS : GLSB_Q2_F64_CTAS;
  int       nst  = 2;
  while (ROWS == 1 && nst < 4 && ctas * (smem_bytes<T, ROWS>(F, nst + 1) + 1024) <= 228 * 1024)
    ++nst;
  return nst;
}

template <typename T, bool GENERAL, bool CTD, bool CELLWISE, int ROWS>
static int launch_rows(const KParams<T> &p, const Shape<T, 3> &S, int F, cudaStream_t s)
{
  static int   n_sm = 0;
  static const int env_nst = getenv("GLSB_Q2_NST") ? atoi(getenv("GLSB_Q2_NST")) : 0;
  const int    nst  = env_nst >= 2 && env_nst <= MAX_NST ? env_nst : ring_depth<T, ROWS>(F);
  const size_t smem = smem_bytes<T, ROWS>(F, nst);
  auto         kern = k_vmult_q2_newton<T, GENERAL, CTD, CELLWISE, ROWS>;
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
  const uint32_t n_batches = (p.cell_end - p.cell_begin + CELLS - 1) / CELLS;
  const int      use_sm    = (p.sm_reserve > 0 && p.sm_reserve < n_sm) ? n_sm - p.sm_reserve : n_sm;
  const uint32_t grid      = n_batches < (uint32_t)(use_sm * bps) ? n_batches : (uint32_t)(use_sm * bps);
  kern<<<grid, TPB, smem, s>>>(p, S, F, nst);
  return cudaGetLastError() != cudaSuccess;
}

template <typename T, bool GENERAL, bool CTD, bool CELLWISE>
static int launch(const KParams<T> &p, const Shape<T, 3> &S, int F, cudaStream_t s)
{
  if (p.QG == 9)
    return launch_rows<T, GENERAL, CTD, CELLWISE, 3>(p, S, F, s);
  return launch_rows<T, GENERAL, CTD, CELLWISE, 1>(p, S, F, s);
}

template <typename T, bool GENERAL>
static int launch_flags(const KParams<T> &p, const Shape<T, 3> &S, int F, cudaStream_t s)
{
  if (p.ctd)
    return p.cell_wise ? launch<T, GENERAL, true, true>(p, S, F, s) : launch<T, GENERAL, true, false>(p, S, F, s);
  return p.cell_wise ? launch<T, GENERAL, false, true>(p, S, F, s) : launch<T, GENERAL, false, false>(p, S, F, s);
}

} // namespace q2
} // namespace glsb
