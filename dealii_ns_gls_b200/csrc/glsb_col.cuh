// Column kernel: cell loop for the degrees whose 2 n^3 values per (cell, component) do not fit the registers of
// the register-tiled kernel (glsb_q2.cuh): FP64 Q3 / Q4 (and any branch of do_vmult_cell, 3-D).
//
// Mapping: one thread per (cell, x-y column): n^2 threads per cell, CPB cells per CTA (thread t -> cell t % CPB,
// column t / CPB, so that the CPB cells of a table row are consecutive lanes).  A thread keeps the n values of
// its column for all dim + 1 components in registers; the z sweeps run in registers with the 1-D matrices as
// immediate constant-bank operands, the x and y sweeps go through shared memory on whole [C][n^3][CPB] volumes
// (6 CTA barriers per batch of cells instead of ~26 in the generic kernel, half the shared-memory reads, and
// no per-access index arithmetic: all offsets are compile-time multiples of CPB).  The quadrature-point physics
// is local to the thread (all components of a point live in one thread): qpoint_physics / load_tables / GeomQ of
// glsb_kernels.cuh are reused unchanged.
#pragma once
#include "glsb_kernels.cuh"


namespace glsb
{
namespace col
{
template <int n, typename T>
struct Cfg
{
  static constexpr int CPB     = (n == 5 && sizeof(T) == 8) ? 4 : 8;
  static constexpr int THREADS = CPB * n * n;
  static constexpr int VOL     = 4 * n * n * n * CPB; // one [C][n^3][CPB] volume
  static constexpr size_t smem = sizeof(T) * (3 * (size_t)VOL + 2 * n * n);
  static constexpr int CTAS = 2; // resident CTAs per SM the kernel is compiled for
};

template <int n, typename T, int BR>
__global__ void __launch_bounds__((Cfg<n, T>::THREADS), (Cfg<n, T>::CTAS)) k_vmult_col(const KParams<T> p, const Shape<T, n> sh)
{
  using G           = Cfg<n, T>;
  constexpr int C   = 4, N2 = n * n, N3 = n * n * n, CPB = G::CPB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *B0 = reinterpret_cast<T *>(smem_raw), *B1 = B0 + G::VOL, *B2 = B1 + G::VOL;
  T *sS = B2 + G::VOL, *sD = sS + N2;
  const int cb = threadIdx.x % CPB, colid = threadIdx.x / CPB, i = colid % n, j = colid / n;
  for (int k = threadIdx.x; k < N2; k += blockDim.x)
    {
      sS[k] = sh.S[k];
      sD[k] = sh.D[k];
    }
  const uint32_t cell0  = p.cell_begin + blockIdx.x * CPB + cb;
  const bool     active = cell_active(p, cell0);
  const uint32_t cell   = cell0 < p.cell_end ? cell0 : p.cell_end - 1;
  // element (c, k, jj, ii) of a volume, this thread's cell
#define GLSB_V(B, c, k, jj, ii) (B)[((((c)*n + (k)) * n + (jj)) * n + (ii)) * CPB + cb]

  // ---- gather: dofs (c, k) of this column ------------------------------------------------------------------
  // the indices stay in 4 n registers for the scatter: reading them again (to compile for 3 resident CTAs at
  // 168 registers) measured slower -- Q3 15.0 against 16.9 GDoF/s, Q4 10.3 against 15.3: a CTA touches only
  // CPB * 4 bytes of each 128-byte index row, so a second pass over the rows costs more than the occupancy gains
  uint32_t iv[C][n];
  T        u[C][n];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < n; ++k)
      iv[c][k] = p.idx[idx_at(p, (uint32_t)(c * N3 + (k * n + j) * n + i), cell)];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < n; ++k)
      {
        u[c][k] = (BR == BR_RESIDUAL) ? p.src[plain_index(p, iv[c][k])] : gather_resolved(p, p.src, iv[c][k]);
        GLSB_V(B0, c, k, j, i) = u[c][k];
      }
  __syncthreads();
  // ---- interpolate in x (B0 -> B1), y (B1 -> registers), z (registers) -----------------------------------
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < n; ++k)
      {
        T s = 0;
#pragma unroll
        for (int a = 0; a < n; ++a)
          s += sS[i * n + a] * GLSB_V(B0, c, k, j, a);
        GLSB_V(B1, c, k, j, i) = s;
      }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < C; ++c)
    {
      T t[n];
#pragma unroll
      for (int k = 0; k < n; ++k)
        {
          T s = 0;
#pragma unroll
          for (int a = 0; a < n; ++a)
            s += sS[j * n + a] * GLSB_V(B1, c, k, a, i);
          t[k] = s;
        }
#pragma unroll
      for (int k = 0; k < n; ++k)
        {
          T s = sh.S[k * n] * t[0];
#pragma unroll
          for (int a = 1; a < n; ++a)
            s += sh.S[k * n + a] * t[a];
          u[c][k]               = s; // value at quadrature point (i, j, k)
          GLSB_V(B0, c, k, j, i) = s; // B0 was last read before the previous barrier
        }
    }
  __syncthreads();
  // ---- quadrature points of the column: derivatives, physics, z part of the integration --------------------
  T out[C][n];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < n; ++k)
      out[c][k] = 0;
#pragma unroll
  for (int k = 0; k < n; ++k)
    {
      const int        q     = (k * n + j) * n + i;
      const int        qi[3] = {i, j, k};
      GeomQ<3, n, T>   geo;
      geo.load(p, sh, cell, q, qi);
      QTables<3, T> tb;
      load_tables<3, T, BR>(p, qpos(p, (uint32_t)q, cell), cell, tb);
      T val[C], rg[C][3], g[C][3], vout[C], gout[C][3], rgq[C][3];
#pragma unroll
      for (int c = 0; c < C; ++c)
        {
          T rx = 0, ry = 0, rz = 0;
#pragma unroll
          for (int a = 0; a < n; ++a)
            {
              rx += sD[i * n + a] * GLSB_V(B0, c, k, j, a);
              ry += sD[j * n + a] * GLSB_V(B0, c, k, a, i);
              rz += sh.D[k * n + a] * u[c][a];
            }
          val[c]   = u[c][k];
          rg[c][0] = rx, rg[c][1] = ry, rg[c][2] = rz;
        }
      geo.template to_physical<C>(rg, g);
      qpoint_physics<3, T, BR>(p, tb, val, g, vout, gout);
      geo.template to_reference<C>(gout, rgq);
#pragma unroll
      for (int c = 0; c < C; ++c)
        {
          out[c][k] += vout[c] * geo.jxw;
#pragma unroll
          for (int a = 0; a < n; ++a)
            out[c][a] += sh.D[k * n + a] * rgq[c][2];
          GLSB_V(B1, c, k, j, i) = rgq[c][0]; // B1 was last read before the previous barrier
          GLSB_V(B2, c, k, j, i) = rgq[c][1];
        }
    }
  __syncthreads();
  // ---- x and y parts of the integration, then the transposed interpolation z (registers), y, x -------------
#pragma unroll
  for (int c = 0; c < C; ++c)
    {
#pragma unroll
      for (int k = 0; k < n; ++k)
        {
          T s = out[c][k];
#pragma unroll
          for (int a = 0; a < n; ++a)
            s += sD[a * n + i] * GLSB_V(B1, c, k, j, a) + sD[a * n + j] * GLSB_V(B2, c, k, a, i);
          out[c][k] = s;
        }
      T w[n];
#pragma unroll
      for (int kk = 0; kk < n; ++kk)
        {
          T s = sh.S[kk] * out[c][0];
#pragma unroll
          for (int k = 1; k < n; ++k)
            s += sh.S[k * n + kk] * out[c][k];
          w[kk] = s;
        }
#pragma unroll
      for (int k = 0; k < n; ++k)
        GLSB_V(B0, c, k, j, i) = w[k]; // B0 was last read before the previous barrier
    }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < n; ++k)
      {
        T s = 0;
#pragma unroll
        for (int a = 0; a < n; ++a)
          s += sS[a * n + j] * GLSB_V(B0, c, k, a, i);
        GLSB_V(B1, c, k, j, i) = s; // B1 / B2 were last read before the previous barrier
      }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < n; ++k)
      {
        T s = 0;
#pragma unroll
        for (int a = 0; a < n; ++a)
          s += sS[a * n + i] * GLSB_V(B1, c, k, j, a);
        if (p.sign_negative)
          s = -s;
        if (active)
          scatter_resolved(p, p.dst, iv[c][k], s);
      }
#undef GLSB_V
}

template <int n, typename T>
static int launch(int branch, const KParams<T> &p, const Shape<T, n> &S, cudaStream_t s)
{
  using G = Cfg<n, T>;
  if (p.cell_end <= p.cell_begin)
    return 0;
  const unsigned grid = (unsigned)((p.cell_end - p.cell_begin + G::CPB - 1) / G::CPB);
#define GLSB_COL_LAUNCH(BR)                                                                                        \
  {                                                                                                                \
    if (G::smem > 48 * 1024 &&                                                                                     \
        cudaFuncSetAttribute(k_vmult_col<n, T, BR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem) != \
          cudaSuccess)                                                                                             \
      return 1;                                                                                                    \
    k_vmult_col<n, T, BR><<<grid, G::THREADS, G::smem, s>>>(p, S);                                                 \
  }
  if (branch == BR_NEWTON)
    GLSB_COL_LAUNCH(BR_NEWTON)
  else if (branch == BR_FIXED_POINT)
    GLSB_COL_LAUNCH(BR_FIXED_POINT)
  else
    GLSB_COL_LAUNCH(BR_RESIDUAL)
#undef GLSB_COL_LAUNCH
  return cudaGetLastError() != cudaSuccess;
}
} // namespace col
} // namespace glsb
