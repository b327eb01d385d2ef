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
//     u, p, the needed entries of grad u / grad p and the SUPG residual with warp shuffles
//     (4 broadcasts + a 2-step butterfly for div u + a 3-step rotation in which every source lane
//     serves exactly one reader).
//   * the q-point tables (U, grad U, grad P [, du/dt_old, delta_q, J^-T, JxW]) of a batch are
//     streamed HBM -> shared memory one quadrature layer ahead with cp.async.bulk (TMA engine,
//     mbarrier complete_tx) into an NST-deep ring; rows are 32 cells wide (256 B in FP64).
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
__device__ __forceinline__ T sel4(int c, T a0, T a1, T a2, T a3)
{
  return c == 0 ? a0 : (c == 1 ? a1 : (c == 2 ? a2 : a3));
}

template <typename T>
__host__ __device__ constexpr size_t stage_elems(int F)
{
  return (size_t)F * 9 * CELLS;
}

// issue the bulk copies of stage j (batch bi of this CTA, quadrature layer j % 3) into ring slot
// j % NST; executed by all lanes of warp 0
template <typename T, int NST>
__device__ __forceinline__ void issue_stage(const KParams<T> &p, const Q2Stage<T> &sd, T *tab, uint64_t *full,
                                            uint64_t *empty, uint32_t j, uint32_t cell0, int layer, int lane)
{
  const uint32_t b = j % NST, r = j / NST;
  if (r >= 1)
    mbar_wait(&empty[b], (r - 1) & 1);
  T *dst = tab + (size_t)b * stage_elems<T>(sd.F);
  if (lane == 0)
    mbar_expect_tx(&full[b], (uint32_t)(sd.F * 9 * CELLS * sizeof(T)));
  __syncwarp();
  int fo = 0;
  for (int g = 0; g < sd.n_groups; ++g)
    {
      const int rows = sd.nf[g] * 9;
      for (int rr = lane; rr < rows; rr += 32)
        {
          const int f = rr / 9, q9 = rr - 9 * f;
          bulk_g2s(dst + ((size_t)(fo + f) * 9 + q9) * CELLS,
                   sd.base[g] + ((size_t)(f * 27 + layer * 9 + q9) * p.ncp + cell0),
                   (uint32_t)(CELLS * sizeof(T)), &full[b]);
        }
      fo += sd.nf[g];
    }
}

template <typename T, bool GENERAL, int NST>
__global__ void __launch_bounds__(TPB, 2) k_vmult_q2_newton(const KParams<T> p, const Shape<T, 3> sh,
                                                            const Q2Stage<T> sd)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T        *tab   = reinterpret_cast<T *>(smem_raw);
  uint64_t *full  = reinterpret_cast<uint64_t *>(smem_raw + NST * stage_elems<T>(sd.F) * sizeof(T));
  uint64_t *empty = full + NST;

  const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int      c = lane & 3, col = 8 * warp + (lane >> 2);
  const bool     is_p = (c == 3);
  const int      cv   = is_p ? 0 : c; // table row used by this lane (pressure lane: any valid row)
  const unsigned bl   = lane & ~3u;
  const T        m0 = (c == 0) ? T(1) : T(0), m1 = (c == 1) ? T(1) : T(0), m2 = (c == 2) ? T(1) : T(0);

  if (threadIdx.x == 0)
    {
      for (int s = 0; s < NST; ++s)
        {
          mbar_init(&full[s], 1);
          mbar_init(&empty[s], TPB / 32);
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
        issue_stage<T, NST>(p, sd, tab, full, empty, j, p.cell_begin + bj * CELLS, j % 3, lane);
      }

  const T w = p.weight, nu = p.nu;
  uint32_t it = 0;
  for (uint32_t bi = 0; bi < my_n; ++bi)
    {
      const uint32_t batch  = blockIdx.x + bi * gridDim.x;
      const uint32_t cell0  = p.cell_begin + batch * CELLS;
      const uint32_t cellr  = cell0 + col;
      const bool     active = cell_active(p, cellr);
      const uint32_t cell   = cellr < p.cell_end ? cellr : p.cell_end - 1;

      // ---- gather (read_dof_values) ------------------------------------------------------
      T t[27];
      {
        uint32_t iv[27];
#pragma unroll
        for (int j = 0; j < 27; ++j)
          iv[j] = p.idx[(uint64_t)(c * 27 + j) * p.ncp + cell];
#pragma unroll
        for (int j = 0; j < 27; ++j)
          t[j] = gather_resolved(p, p.src, iv[j]);
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
      if (p.cell_wise)
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
          const T *tb = tab + (size_t)slot * stage_elems<T>(sd.F) + col;
#define GLSB_TAB(f, a) tb[((f)*9 + (a)) * CELLS]

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
                  J00 = GLSB_TAB(sd.oJ + 0, a), J01 = GLSB_TAB(sd.oJ + 1, a), J02 = GLSB_TAB(sd.oJ + 2, a);
                  J10 = GLSB_TAB(sd.oJ + 3, a), J11 = GLSB_TAB(sd.oJ + 4, a), J12 = GLSB_TAB(sd.oJ + 5, a);
                  J20 = GLSB_TAB(sd.oJ + 6, a), J21 = GLSB_TAB(sd.oJ + 7, a), J22 = GLSB_TAB(sd.oJ + 8, a);
                  jq  = GLSB_TAB(sd.ojxw, a);
                  g0  = J00 * rx + J10 * ry + J20 * rz;
                  g1  = J01 * rx + J11 * ry + J21 * rz;
                  g2  = J02 * rx + J12 * ry + J22 * rz;
                }
              else
                {
                  g0 = rx * ij0, g1 = ry * ij1, g2 = rz * ij2;
                  jq = cdet * wz * (sh.w[qx] * sh.w[qy]);
                }
              // tables
              const T U0 = GLSB_TAB(sd.oU, a), U1 = GLSB_TAB(sd.oU + 1, a), U2 = GLSB_TAB(sd.oU + 2, a);
              const T H0 = GLSB_TAB(sd.oH + 3 * cv, a), H1 = GLSB_TAB(sd.oH + 3 * cv + 1, a),
                      H2 = GLSB_TAB(sd.oH + 3 * cv + 2, a);
              const T Pc = GLSB_TAB(sd.oP + cv, a);
              const T d1 = p.cell_wise ? d1c : GLSB_TAB(sd.od1q, a);
              const T d2 = p.cell_wise ? d2c : GLSB_TAB(sd.od2q, a);
              // values of all four components of this cell
              const T u0 = __shfl_sync(0xffffffffu, val, bl), u1 = __shfl_sync(0xffffffffu, val, bl + 1),
                      u2 = __shfl_sync(0xffffffffu, val, bl + 2), pp = __shfl_sync(0xffffffffu, val, bl + 3);
              // div u: butterfly over the three velocity lanes (pressure lane contributes 0)
              const T gcc = sel4<T>(c, g0, g1, g2, T(0));
              const T dh  = gcc + __shfl_xor_sync(0xffffffffu, gcc, 1);
              const T div = dh + __shfl_xor_sync(0xffffffffu, dh, 2);
              const T td  = val * w;
              const T sgu = g0 * U0 + g1 * U1 + g2 * U2; // U . grad u_c
              const T ugs = H0 * u0 + H1 * u1 + H2 * u2; // u . grad U_c
              T       y   = sgu + ugs;
              if (p.ctd)
                y = td + y;
              // rotation exchange: in step k lane j is read by lane (j - k) & 3 only
              const T su1 = sel4<T>(c, y, g0, g1, g2);
              const T su2 = sel4<T>(c, g2, y, g0, g1);
              const T su3 = sel4<T>(c, g1, g2, y, g0);
              const T r1  = __shfl_sync(0xffffffffu, su1, bl + ((c + 1) & 3));
              const T r2  = __shfl_sync(0xffffffffu, su2, bl + ((c + 2) & 3));
              const T r3  = __shfl_sync(0xffffffffu, su3, bl + ((c + 3) & 3));
              const T Gc0 = sel3<T>(c, g0, r3, r2); // d_c u_0
              const T Gc1 = sel3<T>(c, r1, g1, r3); // d_c u_1
              const T Gc2 = sel3<T>(c, r2, r1, g2); // d_c u_2
              const T gpc = sel3<T>(c, r3, r2, r1); // d_c p
              // velocity row c
              const T r0  = d1 * (y + gpc);
              const T sgs = H0 * U0 + H1 * U1 + H2 * U2; // U . grad U_c
              T       rb  = Pc + sgs;
              if (p.ctd)
                rb = (GLSB_TAB(sd.oU + cv, a) * w + GLSB_TAB(sd.oO + cv, a)) + rb;
              const T rr1  = d1 * rb;
              const T diag = d2 * div - pp;
              T       vo   = td + sgu + ugs;
              T       o0   = nu * (g0 + Gc0) + U0 * r0 + u0 * rr1 + m0 * diag;
              T       o1   = nu * (g1 + Gc1) + U1 * r0 + u1 * rr1 + m1 * diag;
              T       o2   = nu * (g2 + Gc2) + U2 * r0 + u2 * rr1 + m2 * diag;
              // pressure row: (q, div u) and delta_1 (grad q, residual_0); r1..r3 = y_0..y_2 there
              if (is_p)
                {
                  vo = div;
                  o0 = d1 * (r1 + g0);
                  o1 = d1 * (r2 + g1);
                  o2 = d1 * (r3 + g2);
                }
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
                  issue_stage<T, NST>(p, sd, tab, full, empty, j, p.cell_begin + bj * CELLS, j % 3, lane);
                }
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

      // ---- scatter (distribute_local_to_global) --------------------------------------------
      if (active)
        {
#pragma unroll
          for (int j = 0; j < 27; ++j)
            {
              const uint32_t iv = p.idx[(uint64_t)(c * 27 + j) * p.ncp + cell];
              scatter_resolved(p, p.dst, iv, acc[j]);
            }
        }
    }
}

template <typename T, bool GENERAL, int NST>
static int launch(const KParams<T> &p, const Shape<T, 3> &S, const Q2Stage<T> &sd, cudaStream_t s)
{
  static int blocks_per_sm = -1, n_sm = 0;
  const size_t smem = NST * stage_elems<T>(sd.F) * sizeof(T) + 2 * NST * sizeof(uint64_t);
  auto         kern = k_vmult_q2_newton<T, GENERAL, NST>;
  if (smem > 227 * 1024)
    return -1;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return 1;
  int bps = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, TPB, smem) != cudaSuccess || bps < 1)
    return -1;
  if (blocks_per_sm < 0)
    {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    }
  blocks_per_sm            = bps;
  const uint32_t n_batches = (p.cell_end - p.cell_begin + CELLS - 1) / CELLS;
  const uint32_t grid      = n_batches < (uint32_t)(n_sm * bps) ? n_batches : (uint32_t)(n_sm * bps);
  kern<<<grid, TPB, smem, s>>>(p, S, sd);
  return cudaGetLastError() != cudaSuccess;
}

} // namespace q2
} // namespace glsb
