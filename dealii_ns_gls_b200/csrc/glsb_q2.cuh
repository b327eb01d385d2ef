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
//     [batch][layer][field][9][32 cells]; the block a CTA needs for one quadrature layer is
//     contiguous and is streamed HBM -> shared memory with ONE cp.async.bulk (TMA engine,
//     mbarrier complete_tx) per layer into an NST-deep ring, issued 1-2 layers ahead.
//   * gather straight from global memory (the 4 components of a node are adjacent lanes, so a
//     node-major numbering gives full 32 B sectors), scatter with RED.ADD.F64 atomics.
#pragma once
#include "glsb_kernels.cuh"

namespace glsb
{
namespace q2
{
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

template <typename T>
__host__ __device__ constexpr size_t stage_elems(int F)
{
  return (size_t)F * 9 * CELLS;
}

constexpr int IDX_ELEMS = 108 * CELLS; // dof indices of one batch
constexpr int XROW  = 33;       // row stride of the exchange scratch (elements)
constexpr int XSLOT = 5 * XROW; // rows: value, d_0, d_1, d_2, y

// issue the bulk copy of stage j (quadrature layer j % 3 of this CTA's batch j / 3) into ring slot
// j % NST; called by all lanes of warp 0, one elected lane issues
template <typename T, int NST>
__device__ __forceinline__ void issue_stage(const KParams<T> &p, int F, T *tab, uint64_t *full, uint64_t *empty,
                                            uint32_t j, uint32_t cell0, int layer, int lane)
{
  const uint32_t b = j % NST, r = j / NST;
  if (r >= 1)
    mbar_wait(&empty[b], (r - 1) & 1);
  if (lane == 0)
    {
      const uint32_t bytes = (uint32_t)(stage_elems<T>(F) * sizeof(T));
      const T *src = p.Q + (((uint64_t)(cell0 >> 5) * 3 + layer) * p.FT) * (9 * CELLS);
      mbar_expect_tx(&full[b], bytes);
      bulk_g2s(tab + (size_t)b * stage_elems<T>(F), src, bytes, &full[b]);
    }
  __syncwarp();
}

template <typename T, bool GENERAL, bool CTD, bool CELLWISE, int NST>
__global__ void __launch_bounds__(TPB, 2) k_vmult_q2_newton(const KParams<T> p, const Shape<T, 3> sh, const int F)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T        *tab   = reinterpret_cast<T *>(smem_raw);
  T        *xch   = tab + NST * stage_elems<T>(F);                    // [warp][2][XSLOT]
  uint32_t *ibuf  = reinterpret_cast<uint32_t *>(xch + (TPB / 32) * 2 * XSLOT); // [2][108][32] dof indices
  uint64_t *full  = reinterpret_cast<uint64_t *>(ibuf + 2 * IDX_ELEMS);
  uint64_t *empty = full + NST;
  uint64_t *ifull = empty + NST, *iempty = ifull + 2;

  const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int      c = lane & 3, col = 8 * warp + (lane >> 2);
  const bool     is_p = (c == 3);
  const int      cv   = is_p ? 0 : c; // table row used by this lane (pressure lane: any valid row)
  const int      bl   = lane & ~3;
  const T        m0 = (c == 0) ? T(1) : T(0), m1 = (c == 1) ? T(1) : T(0), m2 = (c == 2) ? T(1) : T(0);
  T             *xw   = xch + warp * 2 * XSLOT;
  // column of this lane's cell in a 32-cell table row, for fields of component row 0, 1, 2 and cv
  const int      colr[4] = {col, (col + 4) & 31, (col + 8) & 31, (col + 4 * cv) & 31};

  if (threadIdx.x == 0)
    {
      for (int s = 0; s < NST; ++s)
        {
          mbar_init(&full[s], 1);
          mbar_init(&empty[s], TPB / 32);
        }
      for (int s = 0; s < 2; ++s)
        {
          mbar_init(&ifull[s], 1);
          mbar_init(&iempty[s], TPB / 32);
        }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  __syncthreads();

  const uint32_t n_batches = (p.cell_end - p.cell_begin + CELLS - 1) / CELLS;
  const uint32_t my_n      = (blockIdx.x < n_batches) ? (n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t n_stages  = my_n * 3;

  // prologue: fill NST-1 ring slots
  if (warp == 0)
    for (uint32_t j = 0; j < (uint32_t)(NST - 1) && j < n_stages; ++j)
      {
        const uint32_t bj = blockIdx.x + (j / 3) * gridDim.x;
        issue_stage<T, NST>(p, F, tab, full, empty, j, p.cell_begin + bj * CELLS, j % 3, lane);
      }

  // dof indices of batch bi of this CTA -> index ring slot bi & 1 (one bulk copy of 13.5 KB)
  auto issue_idx = [&](uint32_t bi) {
    const uint32_t b = bi & 1, r = bi >> 1;
    if (r >= 1)
      mbar_wait(&iempty[b], (r - 1) & 1);
    if (lane == 0)
      {
        const uint32_t cell0 = p.cell_begin + (blockIdx.x + bi * gridDim.x) * CELLS;
        mbar_expect_tx(&ifull[b], IDX_ELEMS * 4);
        bulk_g2s(ibuf + b * IDX_ELEMS, p.idx + (uint64_t)(cell0 >> 5) * IDX_ELEMS, IDX_ELEMS * 4, &ifull[b]);
      }
    __syncwarp();
  };
  if (warp == 0 && my_n > 0)
    issue_idx(0);

  const T  w = p.weight, nu = p.nu;
  uint32_t it = 0;
  for (uint32_t bi = 0; bi < my_n; ++bi)
    {
      const uint32_t batch  = blockIdx.x + bi * gridDim.x;
      const uint32_t cell0  = p.cell_begin + batch * CELLS;
      const uint32_t cellr  = cell0 + col;
      const bool     active = cell_active(p, cellr);
      const uint32_t cell   = cellr < p.cell_end ? cellr : p.cell_end - 1;
      // this lane's 27 dof indices sit in shared memory (staged one batch ahead)
      const uint32_t *ixs = ibuf + (bi & 1) * IDX_ELEMS + (c * 27) * CELLS + ((col + 8 * c) & 31);
      // cells with constrained dofs are rare: one warp-uniform test instead of one per dof
      const bool slow = __any_sync(0xffffffffu, p.cell_flags[cell] != 0);
      mbar_wait(&ifull[bi & 1], (bi >> 1) & 1);

      // ---- gather (read_dof_values) ------------------------------------------------------
      T t[27];
      {
        uint32_t iv[27];
#pragma unroll
        for (int j = 0; j < 27; ++j)
          iv[j] = ixs[j * CELLS];
        if (!slow)
          {
#pragma unroll
            for (int j = 0; j < 27; ++j)
              t[j] = p.src[iv[j]];
          }
        else
          {
#pragma unroll
            for (int j = 0; j < 27; ++j)
              t[j] = gather_resolved(p, p.src, iv[j]);
          }
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
      for (int qz = 0; qz < 3; ++qz, ++it)
        {
          // runtime (CTA-uniform) layer index: select from the constant bank instead of indexing it
          const T sz0 = sel3<T>(qz, sh.S[0], sh.S[3], sh.S[6]), sz1 = sel3<T>(qz, sh.S[1], sh.S[4], sh.S[7]),
                  sz2 = sel3<T>(qz, sh.S[2], sh.S[5], sh.S[8]);
          const T gz0 = sel3<T>(qz, sh.G[0], sh.G[3], sh.G[6]), gz1 = sel3<T>(qz, sh.G[1], sh.G[4], sh.G[7]),
                  gz2 = sel3<T>(qz, sh.G[2], sh.G[5], sh.G[8]);
          const T wz  = sel3<T>(qz, sh.w[0], sh.w[1], sh.w[2]);
          T       vl[9], wl[9];
#pragma unroll
          for (int a = 0; a < 9; ++a)
            {
              vl[a] = sz0 * t[a] + sz1 * t[a + 9] + sz2 * t[a + 18];
              wl[a] = 0;
            }
          const uint32_t slot = it % NST;
          mbar_wait(&full[slot], (it / NST) & 1);
          const T *tb = tab + (size_t)slot * stage_elems<T>(F);
#define GLSB_TAB(f, a) tb[((f)*9 + (a)) * CELLS + col]          /* fields without a component row */
#define GLSB_TABR(f, a, r) tb[((f)*9 + (a)) * CELLS + colr[r]] /* rotated by 4 * row, see qoff() */

#pragma unroll
          for (int a = 0; a < 9; ++a)
            {
              const int qx = a % 3, qy = a / 3;
              const T   val = vl[a];
              const T rx = sh.D[qx * 3] * vl[3 * qy] + sh.D[qx * 3 + 1] * vl[3 * qy + 1] + sh.D[qx * 3 + 2] * vl[3 * qy + 2];
              const T ry = sh.D[qy * 3] * vl[qx] + sh.D[qy * 3 + 1] * vl[qx + 3] + sh.D[qy * 3 + 2] * vl[qx + 6];
              const T rz = gz0 * t[a] + gz1 * t[a + 9] + gz2 * t[a + 18];
              // geometry: physical gradient of this lane's component
              T g0, g1, g2, jq;
              T J00, J01, J02, J10, J11, J12, J20, J21, J22;
              if (GENERAL)
                {
                  J00 = GLSB_TAB(p.fJ + 0, a), J01 = GLSB_TAB(p.fJ + 1, a), J02 = GLSB_TAB(p.fJ + 2, a);
                  J10 = GLSB_TAB(p.fJ + 3, a), J11 = GLSB_TAB(p.fJ + 4, a), J12 = GLSB_TAB(p.fJ + 5, a);
                  J20 = GLSB_TAB(p.fJ + 6, a), J21 = GLSB_TAB(p.fJ + 7, a), J22 = GLSB_TAB(p.fJ + 8, a);
                  jq  = GLSB_TAB(p.fjxw, a);
                  g0  = J00 * rx + J10 * ry + J20 * rz;
                  g1  = J01 * rx + J11 * ry + J21 * rz;
                  g2  = J02 * rx + J12 * ry + J22 * rz;
                }
              else
                {
                  g0 = rx * ij0, g1 = ry * ij1, g2 = rz * ij2;
                  jq = cdet * wz * (sh.w[qx] * sh.w[qy]);
                }
              // ---- exchange round 1: publish value and gradient of this component --------------
              T *xs = xw + (a & 1) * XSLOT;
              xs[lane]            = val;
              xs[XROW + lane]     = g0;
              xs[2 * XROW + lane] = g1;
              xs[3 * XROW + lane] = g2;
              // tables (field offsets of the prefix are fixed: U 0..2, grad U 3..11, grad P 12..14)
              const T U0 = GLSB_TABR(0, a, 0), U1 = GLSB_TABR(1, a, 1), U2 = GLSB_TABR(2, a, 2);
              const T H0 = GLSB_TABR(3 + 3 * cv, a, 3), H1 = GLSB_TABR(4 + 3 * cv, a, 3), H2 = GLSB_TABR(5 + 3 * cv, a, 3);
              const T Pc = GLSB_TABR(12 + cv, a, 3);
              const T d1 = CELLWISE ? d1c : GLSB_TAB(p.fd1q, a);
              const T d2 = CELLWISE ? d2c : GLSB_TAB(p.fd2q, a);
              __syncwarp();
              const T u0 = xs[bl], u1 = xs[bl + 1], u2 = xs[bl + 2], pp = xs[bl + 3];
              const T div = xs[XROW + bl] + xs[2 * XROW + bl + 1] + xs[3 * XROW + bl + 2];
              // column c of grad u and d_c p: entry (1 + c) of lanes bl .. bl + 3
              const T *xc  = xs + (1 + cv) * XROW + bl;
              const T  Gc0 = xc[0], Gc1 = xc[1], Gc2 = xc[2], gpc = xc[3];
              const T  td  = val * w;
              const T  sgu = g0 * U0 + g1 * U1 + g2 * U2; // U . grad u_c
              const T  ugs = H0 * u0 + H1 * u1 + H2 * u2; // u . grad U_c
              T        y   = sgu + ugs;
              if (CTD)
                y = td + y;
              // ---- exchange round 2: the pressure row needs y of the three velocity rows --------
              xs[4 * XROW + lane] = y;
              // velocity row c
              const T r0  = d1 * (y + gpc);
              const T sgs = H0 * U0 + H1 * U1 + H2 * U2; // U . grad U_c
              T       rb  = Pc + sgs;
              if (CTD)
                rb = (GLSB_TABR(cv, a, 3) * w + GLSB_TABR(p.fO + cv, a, 3)) + rb;
              const T rr1  = d1 * rb;
              const T diag = d2 * div - pp;
              T       vo   = td + sgu + ugs;
              T       o0   = nu * (g0 + Gc0) + U0 * r0 + u0 * rr1 + m0 * diag;
              T       o1   = nu * (g1 + Gc1) + U1 * r0 + u1 * rr1 + m1 * diag;
              T       o2   = nu * (g2 + Gc2) + U2 * r0 + u2 * rr1 + m2 * diag;
              __syncwarp();
              // pressure row: (q, div u) and delta_1 (grad q, residual_0)
              const T y0 = xs[4 * XROW + bl], y1 = xs[4 * XROW + bl + 1], y2 = xs[4 * XROW + bl + 2];
              const T q0 = d1 * (y0 + g0), q1 = d1 * (y1 + g1), q2 = d1 * (y2 + g2);
              vo = is_p ? div : vo;
              o0 = is_p ? q0 : o0;
              o1 = is_p ? q1 : o1;
              o2 = is_p ? q2 : o2;
              // submit_value / submit_gradient: times JxW, back to the reference cell
              vo *= jq;
              T ox, oy, oz;
              if (GENERAL)
                {
                  ox = (J00 * o0 + J01 * o1 + J02 * o2) * jq;
                  oy = (J10 * o0 + J11 * o1 + J12 * o2) * jq;
                  oz = (J20 * o0 + J21 * o1 + J22 * o2) * jq;
                }
              else
                {
                  ox = o0 * (ij0 * jq), oy = o1 * (ij1 * jq), oz = o2 * (ij2 * jq);
                }
              // integrate: collocation derivative transposed in x and y inside the layer, z into acc
              wl[a] += vo;
#pragma unroll
              for (int i = 0; i < 3; ++i)
                {
                  wl[i + 3 * qy] += sh.D[qx * 3 + i] * ox;
                  wl[qx + 3 * i] += sh.D[qy * 3 + i] * oy;
                }
              acc[a] += gz0 * oz;
              acc[a + 9] += gz1 * oz;
              acc[a + 18] += gz2 * oz;
            }
#undef GLSB_TAB
#undef GLSB_TABR
          // release the ring slot, then let warp 0 refill the slot released one layer ago
          __syncwarp();
          if (lane == 0)
            mbar_arrive(&empty[slot]);
          if (warp == 0)
            {
              const uint32_t j = it + NST - 1;
              if (j < n_stages)
                {
                  const uint32_t bj = blockIdx.x + (j / 3) * gridDim.x;
                  issue_stage<T, NST>(p, F, tab, full, empty, j, p.cell_begin + bj * CELLS, j % 3, lane);
                }
              if (qz == 0 && bi + 1 < my_n)
                issue_idx(bi + 1);
            }
#pragma unroll
          for (int a = 0; a < 9; ++a)
            {
              acc[a] += sz0 * wl[a];
              acc[a + 9] += sz1 * wl[a];
              acc[a + 18] += sz2 * wl[a];
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
              acc[i + 3 * q + 9 * k] = sh.S[q] * a + sh.S[3 + q] * b + sh.S[6 + q] * d;
          }
#pragma unroll
      for (int l = 0; l < 9; ++l)
        {
          const T a = acc[3 * l], b = acc[3 * l + 1], d = acc[3 * l + 2];
#pragma unroll
          for (int q = 0; q < 3; ++q)
            acc[3 * l + q] = sh.S[q] * a + sh.S[3 + q] * b + sh.S[6 + q] * d;
        }

      // ---- scatter (distribute_local_to_global): all index loads first, then the atomics ---
      if (active)
        {
          uint32_t iv[27];
#pragma unroll
          for (int j = 0; j < 27; ++j)
            iv[j] = ixs[j * CELLS];
          if (!slow)
            {
#pragma unroll
              for (int j = 0; j < 27; ++j)
                atomic_add(p.dst + iv[j], acc[j]);
            }
          else
            {
#pragma unroll
              for (int j = 0; j < 27; ++j)
                scatter_resolved(p, p.dst, iv[j], acc[j]);
            }
        }
      // release the index ring slot
      __syncwarp();
      if (lane == 0)
        mbar_arrive(&iempty[bi & 1]);
    }
}

template <typename T>
size_t smem_bytes(int F, int nst)
{
  return (nst * stage_elems<T>(F) + (TPB / 32) * 2 * XSLOT) * sizeof(T) + 2 * IDX_ELEMS * 4 + (2 * nst + 4) * sizeof(uint64_t);
}

template <typename T, bool GENERAL, bool CTD, bool CELLWISE, int NST>
static int launch(const KParams<T> &p, const Shape<T, 3> &S, int F, cudaStream_t s)
{
  static int   n_sm = 0;
  const size_t smem = smem_bytes<T>(F, NST);
  auto         kern = k_vmult_q2_newton<T, GENERAL, CTD, CELLWISE, NST>;
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
  kern<<<grid, TPB, smem, s>>>(p, S, F);
  return cudaGetLastError() != cudaSuccess;
}

template <typename T, bool GENERAL, int NST>
static int launch_flags(const KParams<T> &p, const Shape<T, 3> &S, int F, cudaStream_t s)
{
  if (p.ctd)
    return p.cell_wise ? launch<T, GENERAL, true, true, NST>(p, S, F, s) : launch<T, GENERAL, true, false, NST>(p, S, F, s);
  return p.cell_wise ? launch<T, GENERAL, false, true, NST>(p, S, F, s) : launch<T, GENERAL, false, false, NST>(p, S, F, s);
}

} // namespace q2
} // namespace glsb
