// Multigrid transfer between two levels and the vector operations of a device-resident Krylov solver
// (include/glsb200.h, sections "multigrid transfer on the device" and "device-resident Krylov vectors";
// SURVEY.md section 8f, ranks 2 and 3).  The reference does both on the host through deal.II:
// MGTransferGlobalCoarsening / MGTwoLevelTransfer (main.cc:540-563) and SolverGMRES (solver_l.cc:46-74).
#include "../../include/glsb200.h"
#include "glsb_common.h"

#include <cstdio>
#include <string>
#include <vector>

using namespace glsb;

namespace
{
std::string g_transfer_create_error;

template <typename T>
struct TParams
{
  const uint32_t *cidx, *fidx, *row_ptr, *ecol;
  const T        *eval, *w;
  int             n, nloc, ndof, nch, dim;
  T               P[2 * MAX_N * MAX_N]; // P[(a * n + l) * n + j]
  T               R[MAX_N * MAX_N];     // R[j * n + l]
  int             rchild[MAX_N];
};

template <typename T>
__device__ __forceinline__ T coarse_read(const TParams<T> &p, const T *__restrict__ v, uint32_t iv)
{
  if (!(iv & GLSB_CONSTRAINED_BIT))
    return v[iv];
  const uint32_t r = iv & ~GLSB_CONSTRAINED_BIT;
  T              s = 0;
  for (uint32_t e = p.row_ptr[r]; e < p.row_ptr[r + 1]; ++e)
    s += p.eval[e] * v[p.ecol[e]];
  return s;
}

// One 1-D sweep of the tensor-product embedding through shared memory.  A cell and its two children per
// direction: x' = a * n + l is node l of child a (2n entries, the node the children share appears twice, as in
// the cell-wise loop of deal.II).  expand: out[r][x'][i] = sum_j P[x'][j] in[r][j][i]; contract is the transpose.
template <typename T>
__device__ __forceinline__ void expand_axis(const T *__restrict__ in, T *__restrict__ out, const T *__restrict__ P,
                                            int n, int outer, int inner)
{
  const int total = outer * 2 * n * inner;
  for (int o = threadIdx.x; o < total; o += blockDim.x)
    {
      const int i = o % inner, x = (o / inner) % (2 * n), r = o / (inner * 2 * n);
      const T  *src = in + (r * n) * inner + i, *pr = P + x * n;
      T         s = 0;
      for (int j = 0; j < n; ++j)
        s += pr[j] * src[j * inner];
      out[o] = s;
    }
}
template <typename T>
__device__ __forceinline__ void contract_axis(const T *__restrict__ in, T *__restrict__ out, const T *__restrict__ P,
                                              int n, int outer, int inner)
{
  const int total = outer * n * inner;
  for (int o = threadIdx.x; o < total; o += blockDim.x)
    {
      const int i = o % inner, j = (o / inner) % n, r = o / (inner * n);
      const T  *src = in + (r * 2 * n) * inner + i;
      T         s = 0;
      for (int x = 0; x < 2 * n; ++x)
        s += P[x * n + j] * src[x * inner];
      out[o] = s;
    }
}

// position of (child, local node l) of component c in the expanded [C][(2n)^dim] block
template <int dim>
__device__ __forceinline__ int expanded_pos(int child, int c, int l, int n)
{
  const int m = 2 * n;
  const int x = (child & 1) * n + l % n, y = ((child >> 1) & 1) * n + (l / n) % n;
  if (dim == 2)
    return (c * m + y) * m + x;
  const int z = ((child >> 2) & 1) * n + l / (n * n);
  return ((c * m + z) * m + y) * m + x;
}

// fine += W P C_c coarse: one CTA per coarse cell, sum-factorised embedding (dim sweeps through shared memory)
template <typename T, int dim>
__global__ void __launch_bounds__(128) k_prolongate(const TParams<T> p, T *__restrict__ dst, const T *__restrict__ src)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int  n = p.n, nloc = p.nloc, ndof = p.ndof, m = 2 * n, C = dim + 1;
  const int  big = C * (dim == 3 ? m * m * m : m * m);
  T         *X = reinterpret_cast<T *>(smem_raw), *Y = X + big, *sP = Y + big;
  const uint64_t cell = blockIdx.x;
  for (int d = threadIdx.x; d < ndof; d += blockDim.x)
    X[d] = coarse_read(p, src, p.cidx[cell * ndof + d]); // [C][n^dim], constraints resolved
  for (int k = threadIdx.x; k < 2 * n * n; k += blockDim.x)
    sP[k] = p.P[k];
  __syncthreads();
  if (dim == 3)
    {
      expand_axis(X, Y, sP, n, C * n * n, 1); // [C][n][n][2n]
      __syncthreads();
      expand_axis(Y, X, sP, n, C * n, m);     // [C][n][2n][2n]
      __syncthreads();
      expand_axis(X, Y, sP, n, C, m * m);     // [C][2n][2n][2n]
    }
  else
    {
      expand_axis(X, Y, sP, n, C * n, 1); // [C][n][2n]
      __syncthreads();
      expand_axis(Y, X, sP, n, C, m);     // [C][2n][2n]
    }
  __syncthreads();
  const T  *E     = dim == 3 ? Y : X;
  const int total = p.nch * ndof;
  for (int o = threadIdx.x; o < total; o += blockDim.x)
    {
      const int      child = o / ndof, d = o - child * ndof, c = d / nloc, l = d - c * nloc;
      const uint32_t fi = p.fidx[cell * total + o];
      const T        wv = p.w ? p.w[fi] : T(1);
      if (wv != T(0))
        atomicAdd(dst + fi, wv * E[expanded_pos<dim>(child, c, l, n)]);
    }
}

// coarse += C_c^T P^T W fine
template <typename T, int dim>
__global__ void __launch_bounds__(128) k_restrict(const TParams<T> p, T *__restrict__ dst, const T *__restrict__ src)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int  n = p.n, nloc = p.nloc, ndof = p.ndof, m = 2 * n, C = dim + 1;
  const int  big = C * (dim == 3 ? m * m * m : m * m);
  T         *X = reinterpret_cast<T *>(smem_raw), *Y = X + big, *sP = Y + big;
  const uint64_t cell  = blockIdx.x;
  const int      total = p.nch * ndof;
  for (int o = threadIdx.x; o < total; o += blockDim.x)
    {
      const int      child = o / ndof, d = o - child * ndof, c = d / nloc, l = d - c * nloc;
      const uint32_t fi = p.fidx[cell * total + o];
      X[expanded_pos<dim>(child, c, l, n)] = (p.w ? p.w[fi] : T(1)) * src[fi];
    }
  for (int k = threadIdx.x; k < 2 * n * n; k += blockDim.x)
    sP[k] = p.P[k];
  __syncthreads();
  if (dim == 3)
    {
      contract_axis(X, Y, sP, n, C, m * m);     // [C][n][2n][2n]
      __syncthreads();
      contract_axis(Y, X, sP, n, C * n, m);     // [C][n][n][2n]
      __syncthreads();
      contract_axis(X, Y, sP, n, C * n * n, 1); // [C][n][n][n]
    }
  else
    {
      contract_axis(X, Y, sP, n, C, m);     // [C][n][2n]
      __syncthreads();
      contract_axis(Y, X, sP, n, C * n, 1); // [C][n][n]
    }
  __syncthreads();
  const T *R = dim == 3 ? Y : X;
  for (int d = threadIdx.x; d < ndof; d += blockDim.x)
    {
      const T        s  = R[d];
      const uint32_t iv = p.cidx[cell * ndof + d];
      if (!(iv & GLSB_CONSTRAINED_BIT))
        atomicAdd(dst + iv, s);
      else
        {
          const uint32_t r = iv & ~GLSB_CONSTRAINED_BIT;
          for (uint32_t e = p.row_ptr[r]; e < p.row_ptr[r + 1]; ++e)
            atomicAdd(dst + p.ecol[e], p.eval[e] * s);
        }
    }
}

// coarse = fine function at the coarse support points (set, not add: the fine function is continuous, so
// every cell writes the same value to a shared dof)
template <typename T, int dim>
__global__ void __launch_bounds__(128) k_interpolate(const TParams<T> p, T *__restrict__ dst, const T *__restrict__ src)
{
  const int      n = p.n, nloc = p.nloc, ndof = p.ndof;
  const uint64_t cell = blockIdx.x;
  for (int d = threadIdx.x; d < ndof; d += blockDim.x)
    {
      const uint32_t iv = p.cidx[cell * ndof + d];
      if (iv & GLSB_CONSTRAINED_BIT)
        continue;
      const int c = d / nloc, j = d - c * nloc;
      const int j0 = j % n, j1 = (j / n) % n, j2 = j / (n * n);
      const int child = p.rchild[j0] + 2 * p.rchild[j1] + (dim == 3 ? 4 * p.rchild[j2] : 0);
      const uint32_t *fi = p.fidx + (cell * p.nch + child) * ndof + c * nloc;
      T               s  = 0;
      for (int l2 = 0; l2 < (dim == 3 ? n : 1); ++l2)
        for (int l1 = 0; l1 < n; ++l1)
          for (int l0 = 0; l0 < n; ++l0)
            {
              T r = p.R[j0 * n + l0] * p.R[j1 * n + l1];
              if (dim == 3)
                r *= p.R[j2 * n + l2];
              if (r != T(0))
                s += r * src[fi[(l2 * n + l1) * n + l0]];
            }
      dst[iv] = s;
    }
}

// ---- vector kernels -----------------------------------------------------------------------------------
constexpr int DOT_KC      = 8;    // inner products per pass over w
constexpr int DOT_THREADS = 256;
constexpr int DOT_BLOCKS  = 1184; // 148 SMs x 8

// partial[b * DOT_KC + j] = this block's share of V_j . w
template <typename T>
__global__ void __launch_bounds__(DOT_THREADS)
  k_multi_dot(double *__restrict__ partial, const T *__restrict__ V, uint64_t stride, int k, const T *__restrict__ w,
              uint64_t n)
{
  double acc[DOT_KC];
#pragma unroll
  for (int j = 0; j < DOT_KC; ++j)
    acc[j] = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    {
      const double wi = (double)w[i];
#pragma unroll
      for (int j = 0; j < DOT_KC; ++j)
        if (j < k)
          acc[j] += (double)V[j * stride + i] * wi;
    }
  __shared__ double red[DOT_KC][DOT_THREADS / 32];
#pragma unroll
  for (int j = 0; j < DOT_KC; ++j)
    {
      double v = acc[j];
      for (int o = 16; o > 0; o >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0)
        red[j][threadIdx.x >> 5] = v;
    }
  __syncthreads();
  if (threadIdx.x < DOT_KC)
    {
      double v = 0;
      for (int wv = 0; wv < DOT_THREADS / 32; ++wv)
        v += red[threadIdx.x][wv];
      partial[(uint64_t)blockIdx.x * DOT_KC + threadIdx.x] = v;
    }
}
// out[j] = sum over blocks, in block order (deterministic)
__global__ void k_dot_finish(double *__restrict__ out, const double *__restrict__ partial, int n_blocks, int k)
{
  const int j = blockIdx.x;
  if (j >= k)
    return;
  __shared__ double red[32];
  double            v = 0;
  for (int b = threadIdx.x; b < n_blocks; b += blockDim.x)
    v += partial[(uint64_t)b * DOT_KC + j];
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0)
    red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0)
    {
      double s = 0;
      for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv)
        s += red[wv];
      out[j] = s;
    }
}

template <typename T>
__global__ void k_multi_axpy(T *__restrict__ w, const T *__restrict__ V, uint64_t stride, int k,
                             const double *__restrict__ coef, double scale, uint64_t n)
{
  __shared__ double sc[64];
  for (int j0 = 0; j0 < k; j0 += 64)
    {
      const int kk = min(64, k - j0);
      __syncthreads();
      if (threadIdx.x < kk)
        sc[threadIdx.x] = scale * coef[j0 + threadIdx.x];
      __syncthreads();
      for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        {
          double s = 0;
          for (int j = 0; j < kk; ++j)
            s += sc[j] * (double)V[(uint64_t)(j0 + j) * stride + i];
          w[i] = (T)((double)w[i] + s);
        }
    }
}

template <typename T>
__global__ void k_axpby(T *__restrict__ y, T a, const T *__restrict__ x, T b, uint64_t n)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    y[i] = (b == T(0)) ? a * x[i] : a * x[i] + b * y[i];
}

template <typename D, typename S>
__global__ void k_convert(D *__restrict__ dst, const S *__restrict__ src, uint64_t n)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = (D)src[i];
}

template <typename T>
__global__ void k_zero_indexed(T *__restrict__ v, const uint32_t *__restrict__ idx, uint64_t n)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    v[idx[i]] = T(0);
}

// one warp per row
template <typename T>
__global__ void k_dense_apply(T *__restrict__ y, const double *__restrict__ A, const T *__restrict__ x, uint32_t m,
                              uint32_t n)
{
  const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= m)
    return;
  const double *a = A + (uint64_t)row * n;
  double        s = 0;
  for (uint32_t c = lane; c < n; c += 32)
    s += a[c] * (double)x[c];
  for (int o = 16; o > 0; o >>= 1)
    s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0)
    y[row] = (T)s;
}

unsigned grid_for(uint64_t n, int threads, unsigned cap)
{
  const uint64_t g = (n + threads - 1) / threads;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

struct DevMem
{
  void *p = nullptr;
  ~DevMem()
  {
    if (p)
      cudaFree(p);
  }
  template <typename U>
  bool upload(const U *host, size_t count)
  {
    if (count == 0)
      return true;
    if (cudaMalloc(&p, count * sizeof(U)) != cudaSuccess)
      return false;
    return cudaMemcpy(p, host, count * sizeof(U), cudaMemcpyHostToDevice) == cudaSuccess;
  }
};

// scratch of the two-stage reduction, one per device, grown on demand (calls on one device are expected
// from one host thread, like every other entry point of this library)
double *dot_scratch(int device)
{
  static double *buf[64] = {};
  if (device < 0 || device >= 64)
    return nullptr;
  if (!buf[device])
    if (cudaMalloc(&buf[device], sizeof(double) * DOT_BLOCKS * DOT_KC) != cudaSuccess)
      buf[device] = nullptr;
  return buf[device];
}
} // namespace

struct glsb_transfer
{
  int      dim = 0, degree = 0, n = 0, number_type = 0, device = 0, ndof = 0, nloc = 0, nch = 0;
  uint64_t n_coarse_cells = 0, n_fine = 0, n_coarse = 0;
  uint32_t n_rows = 0;
  DevMem   cidx, fidx, row_ptr, ecol, eval, w;
  double   P[2 * MAX_N * MAX_N], R[MAX_N * MAX_N];
  int      rchild[MAX_N];
  std::string error;
};

namespace
{
template <typename T>
TParams<T> make_params(const glsb_transfer *t)
{
  TParams<T> p;
  p.cidx    = static_cast<const uint32_t *>(t->cidx.p);
  p.fidx    = static_cast<const uint32_t *>(t->fidx.p);
  p.row_ptr = static_cast<const uint32_t *>(t->row_ptr.p);
  p.ecol    = static_cast<const uint32_t *>(t->ecol.p);
  p.eval    = static_cast<const T *>(t->eval.p);
  p.w       = static_cast<const T *>(t->w.p);
  p.n = t->n, p.nloc = t->nloc, p.ndof = t->ndof, p.nch = t->nch, p.dim = t->dim;
  for (int i = 0; i < 2 * t->n * t->n; ++i)
    p.P[i] = (T)t->P[i];
  for (int i = 0; i < t->n * t->n; ++i)
    p.R[i] = (T)t->R[i];
  for (int i = 0; i < t->n; ++i)
    p.rchild[i] = t->rchild[i];
  return p;
}

enum TransferOp
{
  OP_PROLONGATE,
  OP_RESTRICT,
  OP_INTERPOLATE
};

template <typename T, int dim>
int run_transfer(glsb_transfer *t, int which, void *dst, const void *src, cudaStream_t s)
{
  if (t->n_coarse_cells == 0)
    return 0;
  const TParams<T> p    = make_params<T>(t);
  const unsigned   grid = (unsigned)t->n_coarse_cells;
  const int        m    = 2 * t->n;
  const size_t     sm   = sizeof(T) * (2 * (size_t)(dim + 1) * (dim == 3 ? m * m * m : m * m) + 2 * t->n * t->n);
  if (which != OP_INTERPOLATE && sm > 48 * 1024)
    {
      auto kern = which == OP_PROLONGATE ? k_prolongate<T, dim> : k_restrict<T, dim>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
        return 1;
    }
  if (which == OP_PROLONGATE)
    k_prolongate<T, dim><<<grid, 128, sm, s>>>(p, static_cast<T *>(dst), static_cast<const T *>(src));
  else if (which == OP_RESTRICT)
    k_restrict<T, dim><<<grid, 128, sm, s>>>(p, static_cast<T *>(dst), static_cast<const T *>(src));
  else
    k_interpolate<T, dim><<<grid, 128, 0, s>>>(p, static_cast<T *>(dst), static_cast<const T *>(src));
  return cudaGetLastError() != cudaSuccess;
}

int transfer_call(glsb_transfer *t, int which, void *dst, const void *src, void *stream)
{
  if (!t || !dst || !src)
    return 1;
  cudaSetDevice(t->device);
  cudaStream_t s  = static_cast<cudaStream_t>(stream);
  int          rc = 1;
  if (t->number_type == GLSB_F64)
    rc = t->dim == 3 ? run_transfer<double, 3>(t, which, dst, src, s) : run_transfer<double, 2>(t, which, dst, src, s);
  else
    rc = t->dim == 3 ? run_transfer<float, 3>(t, which, dst, src, s) : run_transfer<float, 2>(t, which, dst, src, s);
  if (rc)
    t->error = std::string("transfer kernel launch failed: ") + cudaGetErrorString(cudaGetLastError());
  return rc;
}
} // namespace

extern "C" {

int glsb_transfer_create(const glsb_transfer_desc *d, glsb_transfer **out)
{
  if (!d || !out)
    {
      g_transfer_create_error = "null argument";
      return 1;
    }
  *out = nullptr;
  if (d->abi_version != GLSB_ABI_VERSION)
    {
      g_transfer_create_error = "abi_version mismatch";
      return 1;
    }
  if ((d->dim != 2 && d->dim != 3) || d->degree < 1 || d->degree > MAX_N - 1 ||
      (d->number_type != GLSB_F64 && d->number_type != GLSB_F32))
    {
      g_transfer_create_error = "unsupported dim / degree / number_type";
      return 1;
    }
  if (d->n_coarse_cells > 0x7fffffffull || (d->n_coarse_cells && (!d->coarse_dof_indices || !d->fine_dof_indices)))
    {
      g_transfer_create_error = "bad cell arrays";
      return 1;
    }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || d->device < 0 || d->device >= n_dev ||
      cudaSetDevice(d->device) != cudaSuccess)
    {
      g_transfer_create_error = "no usable CUDA device (there is no CPU fallback)";
      return 1;
    }
  glsb_transfer *t  = new glsb_transfer;
  t->dim            = d->dim;
  t->degree         = d->degree;
  t->n              = d->degree + 1;
  t->number_type    = d->number_type;
  t->device         = d->device;
  t->nloc           = d->dim == 2 ? t->n * t->n : t->n * t->n * t->n;
  t->ndof           = (d->dim + 1) * t->nloc;
  t->nch            = 1 << d->dim;
  t->n_coarse_cells = d->n_coarse_cells;
  t->n_fine         = d->n_fine_dofs;
  t->n_coarse       = d->n_coarse_dofs;
  t->n_rows         = d->n_constraint_rows;
  compute_transfer_host(d->degree, t->P, t->R, t->rchild);
  bool ok = t->cidx.upload(d->coarse_dof_indices, (size_t)d->n_coarse_cells * t->ndof) &&
            t->fidx.upload(d->fine_dof_indices, (size_t)d->n_coarse_cells * t->nch * t->ndof);
  if (ok && d->n_constraint_rows)
    {
      const uint32_t ne = d->row_ptr[d->n_constraint_rows];
      ok = t->row_ptr.upload(d->row_ptr, (size_t)d->n_constraint_rows + 1) && t->ecol.upload(d->entry_col, ne);
      if (ok && ne)
        {
          if (d->number_type == GLSB_F64)
            ok = t->eval.upload(d->entry_val, ne);
          else
            {
              std::vector<float> f(d->entry_val, d->entry_val + ne);
              ok = t->eval.upload(f.data(), ne);
            }
        }
    }
  if (ok && d->weights)
    {
      if (d->number_type == GLSB_F64)
        ok = t->w.upload(d->weights, (size_t)d->n_fine_dofs);
      else
        {
          std::vector<float> f(d->weights, d->weights + d->n_fine_dofs);
          ok = t->w.upload(f.data(), (size_t)d->n_fine_dofs);
        }
    }
  if (!ok)
    {
      g_transfer_create_error = "device allocation / upload failed";
      delete t;
      return 1;
    }
  *out = t;
  return 0;
}

void glsb_transfer_destroy(glsb_transfer *t)
{
  if (t)
    {
      cudaSetDevice(t->device);
      delete t;
    }
}

const char *glsb_transfer_last_error(const glsb_transfer *t) { return t ? t->error.c_str() : g_transfer_create_error.c_str(); }

int glsb_transfer_prolongate_and_add(glsb_transfer *t, void *dst_fine, const void *src_coarse, void *stream)
{
  return transfer_call(t, OP_PROLONGATE, dst_fine, src_coarse, stream);
}
int glsb_transfer_restrict_and_add(glsb_transfer *t, void *dst_coarse, const void *src_fine, void *stream)
{
  return transfer_call(t, OP_RESTRICT, dst_coarse, src_fine, stream);
}
int glsb_transfer_interpolate(glsb_transfer *t, void *dst_coarse, const void *src_fine, void *stream)
{
  return transfer_call(t, OP_INTERPOLATE, dst_coarse, src_fine, stream);
}

int glsb_vec_multi_dot(double *out_dev, const void *V, uint64_t stride, int k, const void *w, uint64_t n, int type,
                       void *stream)
{
  if (!out_dev || !V || !w || k < 0 || (type != GLSB_F64 && type != GLSB_F32))
    return 1;
  int dev = 0;
  cudaGetDevice(&dev);
  double *partial = dot_scratch(dev);
  if (!partial)
    return 1;
  cudaStream_t   s      = static_cast<cudaStream_t>(stream);
  const unsigned blocks = grid_for(n, DOT_THREADS * 4, DOT_BLOCKS);
  for (int j0 = 0; j0 < k; j0 += DOT_KC)
    {
      const int kk = k - j0 < DOT_KC ? k - j0 : DOT_KC;
      if (type == GLSB_F64)
        k_multi_dot<double><<<blocks, DOT_THREADS, 0, s>>>(partial, static_cast<const double *>(V) + (uint64_t)j0 * stride,
                                                           stride, kk, static_cast<const double *>(w), n);
      else
        k_multi_dot<float><<<blocks, DOT_THREADS, 0, s>>>(partial, static_cast<const float *>(V) + (uint64_t)j0 * stride,
                                                          stride, kk, static_cast<const float *>(w), n);
      k_dot_finish<<<kk, 256, 0, s>>>(out_dev + j0, partial, (int)blocks, kk);
    }
  return cudaGetLastError() != cudaSuccess;
}

int glsb_vec_multi_axpy(void *w, const void *V, uint64_t stride, int k, const double *coef_dev, double scale,
                        uint64_t n, int type, void *stream)
{
  if (!w || !V || !coef_dev || k < 0 || (type != GLSB_F64 && type != GLSB_F32))
    return 1;
  if (k == 0 || n == 0)
    return 0;
  cudaStream_t   s      = static_cast<cudaStream_t>(stream);
  const unsigned blocks = grid_for(n, 256, 148 * 16);
  if (type == GLSB_F64)
    k_multi_axpy<double><<<blocks, 256, 0, s>>>(static_cast<double *>(w), static_cast<const double *>(V), stride, k,
                                                coef_dev, scale, n);
  else
    k_multi_axpy<float><<<blocks, 256, 0, s>>>(static_cast<float *>(w), static_cast<const float *>(V), stride, k,
                                               coef_dev, scale, n);
  return cudaGetLastError() != cudaSuccess;
}

int glsb_vec_axpby(void *y, double a, const void *x, double b, uint64_t n, int type, void *stream)
{
  if (!y || !x || (type != GLSB_F64 && type != GLSB_F32))
    return 1;
  if (n == 0)
    return 0;
  cudaStream_t   s      = static_cast<cudaStream_t>(stream);
  const unsigned blocks = grid_for(n, 256, 148 * 16);
  if (type == GLSB_F64)
    k_axpby<double><<<blocks, 256, 0, s>>>(static_cast<double *>(y), a, static_cast<const double *>(x), b, n);
  else
    k_axpby<float><<<blocks, 256, 0, s>>>(static_cast<float *>(y), (float)a, static_cast<const float *>(x), (float)b, n);
  return cudaGetLastError() != cudaSuccess;
}

int glsb_vec_convert(void *dst, int dst_type, const void *src, int src_type, uint64_t n, void *stream)
{
  if (!dst || !src)
    return 1;
  if (n == 0)
    return 0;
  cudaStream_t   s      = static_cast<cudaStream_t>(stream);
  const unsigned blocks = grid_for(n, 256, 148 * 16);
  if (dst_type == GLSB_F32 && src_type == GLSB_F64)
    k_convert<float, double><<<blocks, 256, 0, s>>>(static_cast<float *>(dst), static_cast<const double *>(src), n);
  else if (dst_type == GLSB_F64 && src_type == GLSB_F32)
    k_convert<double, float><<<blocks, 256, 0, s>>>(static_cast<double *>(dst), static_cast<const float *>(src), n);
  else if (dst_type == src_type && (dst_type == GLSB_F64 || dst_type == GLSB_F32))
    return cudaMemcpyAsync(dst, src, n * (dst_type == GLSB_F64 ? 8 : 4), cudaMemcpyDeviceToDevice, s) != cudaSuccess;
  else
    return 1;
  return cudaGetLastError() != cudaSuccess;
}

int glsb_vec_set_zero_indexed(void *v, const uint32_t *idx_dev, uint64_t n_idx, int type, void *stream)
{
  if (!v || (n_idx && !idx_dev) || (type != GLSB_F64 && type != GLSB_F32))
    return 1;
  if (n_idx == 0)
    return 0;
  cudaStream_t   s      = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)((n_idx + 255) / 256);
  if (type == GLSB_F64)
    k_zero_indexed<double><<<blocks, 256, 0, s>>>(static_cast<double *>(v), idx_dev, n_idx);
  else
    k_zero_indexed<float><<<blocks, 256, 0, s>>>(static_cast<float *>(v), idx_dev, n_idx);
  return cudaGetLastError() != cudaSuccess;
}

int glsb_dense_apply(void *y, const double *A_dev, const void *x, uint32_t m, uint32_t n, int type, void *stream)
{
  if (!y || !A_dev || !x || (type != GLSB_F64 && type != GLSB_F32))
    return 1;
  if (m == 0)
    return 0;
  cudaStream_t   s      = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (m + 7) / 8;
  if (type == GLSB_F64)
    k_dense_apply<double><<<blocks, 256, 0, s>>>(static_cast<double *>(y), A_dev, static_cast<const double *>(x), m, n);
  else
    k_dense_apply<float><<<blocks, 256, 0, s>>>(static_cast<float *>(y), A_dev, static_cast<const float *>(x), m, n);
  return cudaGetLastError() != cudaSuccess;
}

} // extern "C"
