// Generic cell kernels: any dim in {2,3}, degree 1..4, Number in {double,float},
// Cartesian or general geometry, all three branches of do_vmult_cell, constraints.
//
// Thread mapping: one thread per (cell, quadrature point / node); CPB cells per CTA,
// thread t -> (cb = t % CPB, l = t / CPB) so that the [field][q][cell] SoA tables and the
// [dof][cell] index array are read with CPB consecutive cells per request.  Sum
// factorisation sweeps go through shared memory.  This is the reference-complete
// fallback path; the register-tiled Q2 kernel in glsb_q2.cuh is the fast path.
#pragma once
#include "glsb_common.h"
#include "../../include/glsb200.h"

#ifndef GLSB_GEN_CTAS
#define GLSB_GEN_CTAS 2 // resident CTAs per SM the generic vmult kernel is compiled for (3 measured slower: spills)
#endif

namespace glsb
{
template <int dim, int n>
struct Geo
{
  static constexpr int n_loc = (dim == 2) ? n * n : n * n * n;
  static constexpr int C     = dim + 1;
  static constexpr int nq    = n_loc;
  static constexpr int cpb()
  {
    int c = 1;
    while (2 * c * n_loc <= 256)
      c *= 2;
    return c;
  }
  static constexpr int CPB     = cpb();
  static constexpr int THREADS = CPB * n_loc;
};

__device__ __forceinline__ void atomic_add(double *a, double v) { atomicAdd(a, v); }
__device__ __forceinline__ void atomic_add(float *a, float v) { atomicAdd(a, v); }

// slow paths (constrained dofs are rare): kept out of line so that the hot kernels stay small
template <typename T>
__device__ __noinline__ T gather_constrained(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ ecol,
                                             const T *__restrict__ eval, const T *__restrict__ src, uint32_t r)
{
  T s = 0;
  for (uint32_t e = row_ptr[r]; e < row_ptr[r + 1]; ++e)
    s += eval[e] * src[ecol[e]];
  return s;
}

template <typename T>
__device__ __noinline__ void scatter_constrained(const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ ecol,
                                                 const T *__restrict__ eval, T *__restrict__ dst, uint32_t r, T v)
{
  for (uint32_t e = row_ptr[r]; e < row_ptr[r + 1]; ++e)
    atomic_add(dst + ecol[e], eval[e] * v);
}

template <typename T>
__device__ __forceinline__ T gather_resolved(const KParams<T> &p, const T *__restrict__ src, uint32_t iv)
{
  if (!(iv & GLSB_CONSTRAINED_BIT))
    return src[iv];
  return gather_constrained<T>(p.row_ptr, p.ecol, p.eval, src, iv & ~GLSB_CONSTRAINED_BIT);
}

template <typename T>
__device__ __forceinline__ uint32_t plain_index(const KParams<T> &p, uint32_t iv)
{
  return (iv & GLSB_CONSTRAINED_BIT) ? p.row_dof[iv & ~GLSB_CONSTRAINED_BIT] : iv;
}

template <typename T>
__device__ __forceinline__ void scatter_resolved(const KParams<T> &p, T *__restrict__ dst, uint32_t iv, T v)
{
  if (!(iv & GLSB_CONSTRAINED_BIT))
    {
      atomic_add(dst + iv, v);
      return;
    }
  scatter_constrained<T>(p.row_ptr, p.ecol, p.eval, dst, iv & ~GLSB_CONSTRAINED_BIT, v);
}

// ---------------------------------------------------------------------------------------
// per-thread context of the generic kernels
// ---------------------------------------------------------------------------------------
template <int dim, int n, typename T>
struct Ctx
{
  using G                    = Geo<dim, n>;
  static constexpr int n_loc = G::n_loc, C = G::C, CPB = G::CPB;

  T  *v;  // [C][n_loc][CPB]
  T  *sg; // [C*dim][n_loc][CPB]
  T  *sS, *sD;
  int cb, l, ii[3];

  __device__ __forceinline__ int at(int c, int ll) const { return (c * n_loc + ll) * CPB + cb; }
  __device__ __forceinline__ int stride(int e) const { return e == 0 ? 1 : (e == 1 ? n : n * n); }

  __device__ void init(unsigned char *smem, const Shape<T, n> &sh)
  {
    v  = reinterpret_cast<T *>(smem);
    sg = v + C * n_loc * CPB;
    sS = sg + C * dim * n_loc * CPB;
    sD = sS + n * n;
    cb = threadIdx.x % CPB;
    l  = threadIdx.x / CPB;
    ii[0] = l % n;
    ii[1] = (l / n) % n;
    ii[2] = l / (n * n);
    for (int k = threadIdx.x; k < n * n; k += blockDim.x)
      {
        sS[k] = sh.S[k];
        sD[k] = sh.D[k];
      }
  }

  // one 1-D sweep over all C components, in place in v.  transpose = false: out_q = sum_i M[q][i] in_i
  __device__ void sweep(const T *M, int e, bool transpose)
  {
    const int ie = ii[e], st = stride(e), base = l - ie * st;
    T         out[C];
#pragma unroll
    for (int c = 0; c < C; ++c)
      {
        T s = 0;
#pragma unroll
        for (int i = 0; i < n; ++i)
          s += (transpose ? M[i * n + ie] : M[ie * n + i]) * v[at(c, base + i * st)];
        out[c] = s;
      }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < C; ++c)
      v[at(c, l)] = out[c];
    __syncthreads();
  }

  // v holds dof values on entry (after a __syncthreads); on exit v holds the values at the
  // quadrature points, val[c] the value and rg[c][e] the reference-cell gradient at point l.
  __device__ void evaluate(T (&val)[C], T (&rg)[C][dim])
  {
#pragma unroll
    for (int e = 0; e < dim; ++e)
      sweep(sS, e, false);
#pragma unroll
    for (int c = 0; c < C; ++c)
      val[c] = v[at(c, l)];
#pragma unroll
    for (int e = 0; e < dim; ++e)
      {
        const int ie = ii[e], st = stride(e), base = l - ie * st;
#pragma unroll
        for (int c = 0; c < C; ++c)
          {
            T s = 0;
#pragma unroll
            for (int i = 0; i < n; ++i)
              s += sD[ie * n + i] * v[at(c, base + i * st)];
            rg[c][e] = s;
          }
      }
  }

  // vq[c] (already times JxW) and rgq[c][e] (reference gradient, times JxW) at point l
  // -> v holds the local result vector (tested with all basis functions) on exit.
  __device__ void integrate(const T (&vq)[C], const T (&rgq)[C][dim])
  {
    __syncthreads(); // everyone is done reading v
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int e = 0; e < dim; ++e)
        sg[((c * dim + e) * n_loc + l) * CPB + cb] = rgq[c][e];
    __syncthreads();
    T out[C];
#pragma unroll
    for (int c = 0; c < C; ++c)
      out[c] = vq[c];
#pragma unroll
    for (int e = 0; e < dim; ++e)
      {
        const int ie = ii[e], st = stride(e), base = l - ie * st;
#pragma unroll
        for (int c = 0; c < C; ++c)
          {
            T s = 0;
#pragma unroll
            for (int i = 0; i < n; ++i)
              s += sD[i * n + ie] * sg[((c * dim + e) * n_loc + base + i * st) * CPB + cb];
            out[c] += s;
          }
      }
#pragma unroll
    for (int c = 0; c < C; ++c)
      v[at(c, l)] = out[c];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < dim; ++e)
      sweep(sS, e, true);
  }
};

// ---------------------------------------------------------------------------------------
// geometry: reference gradient <-> physical gradient
// ---------------------------------------------------------------------------------------
template <int dim, int n, typename T>
struct GeomQ
{
  T ij[dim][dim]; // (J^-1)_{e j}; Cartesian: only the diagonal is used
  T jxw;
  int cart;

  __device__ __forceinline__ void load(const KParams<T> &p, const Shape<T, n> &sh, uint32_t cell, int q, const int (&qi)[3])
  {
    cart = (p.geom == GLSB_GEOM_CARTESIAN);
    if (cart)
      {
        T w = sh.w[qi[0]] * sh.w[qi[1]];
        if (dim == 3)
          w *= sh.w[qi[2]];
#pragma unroll
        for (int e = 0; e < dim; ++e)
          ij[e][e] = p.inv_jac[e * p.ncp + cell];
        jxw = p.jxw[cell] * w;
      }
    else
      {
        const QPos<T> qp = qpos(p, (uint32_t)q, cell);
#pragma unroll
        for (int e = 0; e < dim; ++e)
#pragma unroll
          for (int j = 0; j < dim; ++j)
            ij[e][j] = p.Q[qoff(qp, p.fJ + e * dim + j, 0)];
        jxw = p.Q[qoff(qp, p.fjxw, 0)];
      }
  }
  template <int C>
  __device__ __forceinline__ void to_physical(const T (&rg)[C][dim], T (&g)[C][dim]) const
  {
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int j = 0; j < dim; ++j)
        {
          if (cart)
            g[c][j] = rg[c][j] * ij[j][j];
          else
            {
              T s = 0;
#pragma unroll
              for (int e = 0; e < dim; ++e)
                s += ij[e][j] * rg[c][e];
              g[c][j] = s;
            }
        }
  }
  // test side: rgq[c][e] = sum_j (J^-1)_{e j} gout[c][j] * JxW
  template <int C>
  __device__ __forceinline__ void to_reference(const T (&gout)[C][dim], T (&rgq)[C][dim]) const
  {
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int e = 0; e < dim; ++e)
        {
          if (cart)
            rgq[c][e] = gout[c][e] * ij[e][e] * jxw;
          else
            {
              T s = 0;
#pragma unroll
              for (int j = 0; j < dim; ++j)
                s += ij[e][j] * gout[c][j];
              rgq[c][e] = s * jxw;
            }
        }
  }
};

// ---------------------------------------------------------------------------------------
// quadrature-point physics (operator_ns.cc:949-1182), one point
// ---------------------------------------------------------------------------------------
template <int dim, typename T>
__device__ __forceinline__ void symm_add(T (&gout)[dim + 1][dim], const T (&B)[dim][dim], T factor)
{
  // operator_ns.cc:899-916
#pragma unroll
  for (int d = 0; d < dim; ++d)
    gout[d][d] += B[d][d] * factor;
#pragma unroll
  for (int e = 0; e < dim; ++e)
#pragma unroll
    for (int d = e + 1; d < dim; ++d)
      {
        const T tmp = (B[d][e] + B[e][d]) * (factor * T(0.5));
        gout[d][e] += tmp;
        gout[e][d] += tmp;
      }
}

// the table entries of one quadrature point that a cell application reads (operator_ns.h:117-132);
// loaded once, BEFORE the sum-factorisation sweeps, so that the global loads complete behind them
// (and once per cell, not once per unit vector, in the diagonal kernels)
template <int dim, typename T>
struct QTables
{
  T d1, d2;
  T U[dim], H[dim][dim], P[dim], O[dim], Gold[dim][dim], gold_p[dim];
};

template <int dim, typename T, int BR>
__device__ __forceinline__ void load_tables(const KParams<T> &p, const QPos<T> qp, uint32_t cell, QTables<dim, T> &t)
{
#define GLSB_QF(f, row) p.Q[qoff(qp, (f), (row))]
  t.d1 = p.cell_wise ? p.d1c[cell] : GLSB_QF(p.fd1q, 0);
  t.d2 = p.cell_wise ? p.d2c[cell] : GLSB_QF(p.fd2q, 0);
#pragma unroll
  for (int j = 0; j < dim; ++j)
    t.U[j] = GLSB_QF(p.fU + j, j);
  if (BR == BR_NEWTON)
    {
#pragma unroll
      for (int c = 0; c < dim; ++c)
        {
          t.P[c] = GLSB_QF(p.fP + c, c);
          t.O[c] = p.ctd ? GLSB_QF(p.fO + c, c) : T(0);
#pragma unroll
          for (int j = 0; j < dim; ++j)
            t.H[c][j] = GLSB_QF(p.fH + c * dim + j, c);
        }
    }
  else
    {
      constexpr bool res = (BR == BR_RESIDUAL);
#pragma unroll
      for (int c = 0; c < dim; ++c)
        {
          t.O[c]      = (res && p.has_o) ? GLSB_QF(p.fO + c, c) : T(0);
          t.gold_p[c] = (res && p.theta_ne_1) ? GLSB_QF(p.fgoldp + c, 0) : T(0);
#pragma unroll
          for (int j = 0; j < dim; ++j)
            t.Gold[c][j] = (res && p.theta_ne_1) ? GLSB_QF(p.fGold + c * dim + j, 0) : T(0);
        }
    }
#undef GLSB_QF
}

template <int dim, typename T, int BR>
__device__ __forceinline__ void qpoint_physics(const KParams<T> &p, const QTables<dim, T> &tb,
                                               const T (&val)[dim + 1], const T (&g)[dim + 1][dim],
                                               T (&vout)[dim + 1], T (&gout)[dim + 1][dim])
{
  const T d1 = tb.d1, d2 = tb.d2;
  const T w  = p.weight;
  T       U[dim];
#pragma unroll
  for (int j = 0; j < dim; ++j)
    U[j] = tb.U[j];
#pragma unroll
  for (int c = 0; c <= dim; ++c)
#pragma unroll
    for (int j = 0; j < dim; ++j)
      gout[c][j] = 0;

  if (BR == BR_NEWTON)
    {
      T H[dim][dim], P[dim];
#pragma unroll
      for (int c = 0; c < dim; ++c)
        {
          P[c] = tb.P[c];
#pragma unroll
          for (int j = 0; j < dim; ++j)
            H[c][j] = tb.H[c][j];
        }
      T Gm[dim][dim];
      T div = 0;
#pragma unroll
      for (int c = 0; c < dim; ++c)
        {
#pragma unroll
          for (int j = 0; j < dim; ++j)
            Gm[c][j] = g[c][j];
          div += g[c][c];
        }
      T r0[dim], r1[dim];
#pragma unroll
      for (int c = 0; c < dim; ++c)
        {
          T sgu = 0, ugs = 0, sgs = 0;
#pragma unroll
          for (int j = 0; j < dim; ++j)
            {
              sgu += Gm[c][j] * U[j];
              ugs += H[c][j] * val[j];
              sgs += H[c][j] * U[j];
            }
          const T td = val[c] * w;
          vout[c]    = td + sgu + ugs;
          T a        = g[dim][c] + sgu + ugs;
          T b        = P[c] + sgs;
          if (p.ctd)
            {
              a = td + a;
              b = (U[c] * w + tb.O[c]) + b;
            }
          r0[c] = d1 * a;
          r1[c] = d1 * b;
        }
#pragma unroll
      for (int d = 0; d < dim; ++d)
        gout[d][d] -= val[dim];
      symm_add<dim, T>(gout, Gm, p.nu * T(2));
#pragma unroll
      for (int d0 = 0; d0 < dim; ++d0)
#pragma unroll
        for (int d1i = 0; d1i < dim; ++d1i)
          gout[d0][d1i] += U[d1i] * r0[d0] + val[d1i] * r1[d0];
#pragma unroll
      for (int d = 0; d < dim; ++d)
        gout[d][d] += d2 * div;
      vout[dim] = div;
#pragma unroll
      for (int j = 0; j < dim; ++j)
        gout[dim][j] = r0[j];
    }
  else
    {
      constexpr bool res = (BR == BR_RESIDUAL);
      const T        th  = p.theta;
      T              B[dim][dim], pbar[dim], td[dim];
#pragma unroll
      for (int c = 0; c < dim; ++c)
        {
          td[c]   = val[c] * w;
          pbar[c] = th * g[dim][c];
#pragma unroll
          for (int j = 0; j < dim; ++j)
            B[c][j] = th * g[c][j];
        }
      if (res && p.has_o)
#pragma unroll
        for (int c = 0; c < dim; ++c)
          td[c] += tb.O[c];
      if (res && p.theta_ne_1)
        {
          const T omt = T(1) - th;
#pragma unroll
          for (int c = 0; c < dim; ++c)
            {
              pbar[c] += omt * tb.gold_p[c];
#pragma unroll
              for (int j = 0; j < dim; ++j)
                B[c][j] += omt * tb.Gold[c][j];
            }
        }
      T divb = 0;
#pragma unroll
      for (int c = 0; c < dim; ++c)
        divb += B[c][c];
      T sgb[dim];
#pragma unroll
      for (int c = 0; c < dim; ++c)
        {
          T s = 0;
#pragma unroll
          for (int j = 0; j < dim; ++j)
            s += B[c][j] * U[j];
          sgb[c]  = s;
          vout[c] = td[c] + s;
        }
#pragma unroll
      for (int d = 0; d < dim; ++d)
        gout[d][d] -= val[dim];
      symm_add<dim, T>(gout, B, p.nu * T(2));
#pragma unroll
      for (int d0 = 0; d0 < dim; ++d0)
        {
          const T tdc = p.ctd ? td[d0] : T(0);
          const T r0  = d1 * (tdc + pbar[d0] + sgb[d0]);
#pragma unroll
          for (int d1i = 0; d1i < dim; ++d1i)
            gout[d0][d1i] += U[d1i] * r0;
          gout[dim][d0] = d1 * (tdc + g[dim][d0] + sgb[d0]);
        }
#pragma unroll
      for (int d = 0; d < dim; ++d)
        gout[d][d] += d2 * divb;
      vout[dim] = divb;
    }
}

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
template <int dim, int n, typename T>
constexpr size_t generic_smem_bytes()
{
  using G = Geo<dim, n>;
  return sizeof(T) * ((size_t)(G::C + G::C * dim) * G::n_loc * G::CPB + 2 * n * n);
}

// full cell operator on the local vector in ctx.v (dof values in, tested result out); geo and tb are this
// thread's quadrature-point geometry and table entries, loaded by the caller before the dof values
template <int dim, int n, typename T, int BR>
__device__ __forceinline__ void cell_apply(Ctx<dim, n, T> &ctx, const KParams<T> &p, const GeomQ<dim, n, T> &geo,
                                           const QTables<dim, T> &tb)
{
  constexpr int C = dim + 1;
  T             val[C], rg[C][dim], g[C][dim], vout[C], gout[C][dim], rgq[C][dim];
  ctx.evaluate(val, rg);
  geo.template to_physical<C>(rg, g);
  qpoint_physics<dim, T, BR>(p, tb, val, g, vout, gout);
  geo.template to_reference<C>(gout, rgq);
#pragma unroll
  for (int c = 0; c < C; ++c)
    vout[c] *= geo.jxw;
  ctx.integrate(vout, rgq);
}

// value of the unit vector e_col at (possibly constrained) index iv, resolved like read_dof_values
template <typename T>
__device__ __forceinline__ T unit_resolved(const KParams<T> &p, uint32_t iv, uint32_t col)
{
  if (!(iv & GLSB_CONSTRAINED_BIT))
    return iv == col ? T(1) : T(0);
  const uint32_t r = iv & ~GLSB_CONSTRAINED_BIT;
  T              s = 0;
  for (uint32_t e = p.row_ptr[r]; e < p.row_ptr[r + 1]; ++e)
    if (p.ecol[e] == col)
      s += p.eval[e];
  return s;
}

// UNIT: get_system_matrix -- blockIdx.y is a column j: src is the unit vector e_j (never materialised), dst the
// j-th column of a column-major n x n matrix (p.dst + j * p.unit_stride): all columns in ONE launch
template <int dim, int n, typename T, int BR, bool UNIT = false>
__global__ void __launch_bounds__(Geo<dim, n>::THREADS, GLSB_GEN_CTAS) k_vmult_generic(const KParams<T> p, const Shape<T, n> sh)
{
  using G = Geo<dim, n>;
  extern __shared__ __align__(16) unsigned char smem[];
  Ctx<dim, n, T> ctx;
  ctx.init(smem, sh);
  const uint32_t cell0  = p.cell_begin + blockIdx.x * G::CPB + ctx.cb;
  const bool     active = cell_active(p, cell0);
  const uint32_t cell   = cell0 < p.cell_end ? cell0 : p.cell_end - 1;
  T             *dst    = UNIT ? p.dst + (uint64_t)blockIdx.y * p.unit_stride : p.dst;
  uint32_t       iv[G::C];
#pragma unroll
  for (int c = 0; c < G::C; ++c)
    iv[c] = p.idx[idx_at(p, c * G::n_loc + ctx.l, cell)];
  // this point's geometry and table entries: issued now, consumed after the evaluate sweeps
  GeomQ<dim, n, T> geo;
  geo.load(p, sh, cell, ctx.l, ctx.ii);
  QTables<dim, T> tb;
  load_tables<dim, T, BR>(p, qpos(p, (uint32_t)ctx.l, cell), cell, tb);
#pragma unroll
  for (int c = 0; c < G::C; ++c)
    ctx.v[ctx.at(c, ctx.l)] = UNIT ? unit_resolved(p, iv[c], blockIdx.y) :
                              (BR == BR_RESIDUAL) ? p.src[plain_index(p, iv[c])] : gather_resolved(p, p.src, iv[c]);
  __syncthreads();
  cell_apply<dim, n, T, BR>(ctx, p, geo, tb);
  if (active)
    {
#pragma unroll
      for (int c = 0; c < G::C; ++c)
        {
          T r = ctx.v[ctx.at(c, ctx.l)];
          if (p.sign_negative)
            r = -r;
          if (!UNIT || r != T(0))
            scatter_resolved(p, dst, iv[c], r);
        }
    }
}

// set_linearization_point + compute_penalty_parameters (operator_ns.cc:570-620, :322-420)
template <int dim, int n, typename T>
__global__ void __launch_bounds__(Geo<dim, n>::THREADS) k_linearization(const KParams<T> p, const Shape<T, n> sh)
{
  using G         = Geo<dim, n>;
  constexpr int C = G::C;
  extern __shared__ __align__(16) unsigned char smem[];
  Ctx<dim, n, T> ctx;
  ctx.init(smem, sh);
  const uint32_t cell0  = p.cell_begin + blockIdx.x * G::CPB + ctx.cb;
  const bool     active = cell_active(p, cell0);
  const uint32_t cell   = cell0 < p.cell_end ? cell0 : p.cell_end - 1;
#pragma unroll
  for (int c = 0; c < C; ++c)
    ctx.v[ctx.at(c, ctx.l)] = p.src[plain_index(p, p.idx[idx_at(p, c * G::n_loc + ctx.l, cell)])];
  __syncthreads();
  T val[C], rg[C][dim], g[C][dim];
  ctx.evaluate(val, rg);
  GeomQ<dim, n, T> geo;
  geo.load(p, sh, cell, ctx.l, ctx.ii);
  geo.template to_physical<C>(rg, g);
  const QPos<T> qp = qpos(p, (uint32_t)ctx.l, cell);
  T             u2 = 0;
#pragma unroll
  for (int c = 0; c < dim; ++c)
    u2 += val[c] * val[c];
  // cell-wise u_max: reduce over the points of the cell through shared memory
  ctx.sg[ctx.l * G::CPB + ctx.cb] = sqrt(u2);
  __syncthreads();
  T umax = 0;
  for (int q = 0; q < G::nq; ++q)
    umax = max(umax, ctx.sg[q * G::CPB + ctx.cb]);
  const double h = p.h_min[cell];
  double       d1c, d2c;
  if (p.nu_d < h)
    {
      d1c = p.c1 / sqrt(p.stau * p.stau + (double)umax * (double)umax / (h * h));
      d2c = p.c2 * h;
    }
  else
    {
      d1c = p.c1 * h * h;
      d2c = p.c2 * h * h;
    }
  const double meas = p.measure[cell];
  const T      hq   = (dim == 2) ? T(sqrt(4. * meas / 3.14159265358979323846) / p.degree) :
                                   T(pow(6. * meas / 3.14159265358979323846, 1. / 3.) / p.degree);
  const T      um2  = T(1e-12) + u2;
  const T      x    = T(4) * p.nu / (hq * hq);
  const T      d1q  = T(1) / sqrt(T(p.stau * p.stau) + T(4) * um2 / hq / hq + T(9) * (x * x));
  const T      d2q  = sqrt(um2) * hq * T(0.5);
  if (!active)
    return;
#define GLSB_QF(f, row) p.Q[qoff(qp, (f), (row))]
  if (ctx.l == 0)
    {
      p.d1c[cell] = T(d1c);
      p.d2c[cell] = T(d2c);
    }
  GLSB_QF(p.fd1q, 0) = d1q;
  GLSB_QF(p.fd2q, 0) = d2q;
#pragma unroll
  for (int c = 0; c < dim; ++c)
    {
      GLSB_QF(p.fU + c, c) = val[c];
      GLSB_QF(p.fP + c, c) = g[dim][c];
#pragma unroll
      for (int j = 0; j < dim; ++j)
        GLSB_QF(p.fH + c * dim + j, c) = g[c][j];
    }
#undef GLSB_QF
}

// set_previous_solution (operator_ns.cc:234-320).  GRAD = false: u_time_derivative_old from
// vec_old = sum_i w_i hist_i; GRAD = true: u_old_gradient, p_old_gradient from hist[0].
template <int dim, int n, typename T, bool GRAD>
__global__ void __launch_bounds__(Geo<dim, n>::THREADS) k_previous(const KParams<T> p, const Shape<T, n> sh)
{
  using G         = Geo<dim, n>;
  constexpr int C = G::C;
  extern __shared__ __align__(16) unsigned char smem[];
  Ctx<dim, n, T> ctx;
  ctx.init(smem, sh);
  const uint32_t cell0  = p.cell_begin + blockIdx.x * G::CPB + ctx.cb;
  const bool     active = cell_active(p, cell0);
  const uint32_t cell   = cell0 < p.cell_end ? cell0 : p.cell_end - 1;
#pragma unroll
  for (int c = 0; c < C; ++c)
    {
      const uint32_t i = plain_index(p, p.idx[idx_at(p, c * G::n_loc + ctx.l, cell)]);
      T              s = 0;
      for (int k = 0; k < p.hist_n; ++k)
        s += p.hist_w[k] * p.hist[k][i];
      ctx.v[ctx.at(c, ctx.l)] = s;
    }
  __syncthreads();
  T val[C], rg[C][dim], g[C][dim];
  ctx.evaluate(val, rg);
  if (!active)
    return;
  const QPos<T> qp = qpos(p, (uint32_t)ctx.l, cell);
#define GLSB_QF(f, row) p.Q[qoff(qp, (f), (row))]
  if (!GRAD)
    {
#pragma unroll
      for (int c = 0; c < dim; ++c)
        GLSB_QF(p.fO + c, c) = val[c];
    }
  else
    {
      GeomQ<dim, n, T> geo;
      geo.load(p, sh, cell, ctx.l, ctx.ii);
      geo.template to_physical<C>(rg, g);
#pragma unroll
      for (int c = 0; c < dim; ++c)
        {
          GLSB_QF(p.fgoldp + c, 0) = g[dim][c];
#pragma unroll
          for (int j = 0; j < dim; ++j)
            GLSB_QF(p.fGold + c * dim + j, 0) = g[c][j];
        }
    }
#undef GLSB_QF
}

// get_max_u (operator_ns.cc:530-568)
template <int dim, int n, typename T>
__global__ void __launch_bounds__(Geo<dim, n>::THREADS) k_max_u(const KParams<T> p, const Shape<T, n> sh)
{
  using G         = Geo<dim, n>;
  constexpr int C = G::C;
  extern __shared__ __align__(16) unsigned char smem[];
  Ctx<dim, n, T> ctx;
  ctx.init(smem, sh);
  const uint32_t cell0  = p.cell_begin + blockIdx.x * G::CPB + ctx.cb;
  const bool     active = cell_active(p, cell0);
  const uint32_t cell   = cell0 < p.cell_end ? cell0 : p.cell_end - 1;
#pragma unroll
  for (int c = 0; c < C; ++c)
    ctx.v[ctx.at(c, ctx.l)] = p.src[plain_index(p, p.idx[idx_at(p, c * G::n_loc + ctx.l, cell)])];
  __syncthreads();
  T val[C], rg[C][dim];
  ctx.evaluate(val, rg);
  T u2 = 0;
#pragma unroll
  for (int c = 0; c < dim; ++c)
    u2 += val[c] * val[c];
  // |u| >= 0: the bit pattern of a non-negative double orders like the value
  __shared__ unsigned long long smax;
  if (threadIdx.x == 0)
    smax = 0ull;
  __syncthreads();
  const double m = active ? (double)sqrt(u2) : 0.0;
  atomicMax(&smax, (unsigned long long)__double_as_longlong(m));
  __syncthreads();
  if (threadIdx.x == 0)
    atomicMax(p.max_bits, smax);
}

// diag(A_cell) by unit vectors (MatrixFreeTools::compute_diagonal, operator_ns.cc:210-218) for
// cells without weighted constraint rows; cells with weighted rows go through k_diag_columns.
template <int dim, int n, typename T, int BR>
__global__ void __launch_bounds__(Geo<dim, n>::THREADS) k_diag_generic(const KParams<T> p, const Shape<T, n> sh,
                                                                      const uint8_t *__restrict__ skip_cell)
{
  using G         = Geo<dim, n>;
  constexpr int C = G::C;
  extern __shared__ __align__(16) unsigned char smem[];
  Ctx<dim, n, T> ctx;
  ctx.init(smem, sh);
  const uint32_t cell0  = p.cell_begin + blockIdx.x * G::CPB + ctx.cb;
  const bool     active = cell_active(p, cell0);
  const uint32_t cell   = cell0 < p.cell_end ? cell0 : p.cell_end - 1;
  T              mine[C];
  GeomQ<dim, n, T> geo;
  geo.load(p, sh, cell, ctx.l, ctx.ii);
  QTables<dim, T> tb;
  load_tables<dim, T, BR>(p, qpos(p, (uint32_t)ctx.l, cell), cell, tb);
  for (int j = 0; j < C * G::n_loc; ++j)
    {
      __syncthreads();
#pragma unroll
      for (int c = 0; c < C; ++c)
        ctx.v[ctx.at(c, ctx.l)] = (c * G::n_loc + ctx.l == j) ? T(1) : T(0);
      __syncthreads();
      cell_apply<dim, n, T, BR>(ctx, p, geo, tb);
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (c * G::n_loc + ctx.l == j)
          mine[c] = ctx.v[ctx.at(c, ctx.l)];
    }
  if (!active || (skip_cell != nullptr && skip_cell[cell]))
    return;
#pragma unroll
  for (int c = 0; c < C; ++c)
    {
      const uint32_t iv = p.idx[idx_at(p, c * G::n_loc + ctx.l, cell)];
      if (!(iv & GLSB_CONSTRAINED_BIT))
        atomic_add(p.dst + iv, mine[c]);
    }
}

// diag(A_cell) without unit vectors.  For a trial/test function pair of the SAME component c the
// quadrature-point operator (operator_ns.cc:1067-1182 resp. :955-1066) restricted to that component is
//   value_out = alpha v + beta . g,     grad_out_j = gamma_j v + sum_m K_jm g_m       (v = phi, g = grad phi)
// Newton branch, velocity c:  alpha = w + H_cc, beta = U, gamma_j = U_j delta_1 (t w + H_cc) + [j = c] r1_c,
//                             K = nu I + (nu + delta_2) e_c e_c^T + delta_1 U U^T,
//                             r1_c = delta_1 (t (w U_c + o_c) + P_c + H_c. U)
// fixed-point, velocity c:    alpha = w, beta = theta U, gamma_j = U_j delta_1 t w,
//                             K = nu theta I + theta (nu + delta_2) e_c e_c^T + delta_1 theta U U^T
// pressure (both):            alpha = beta = gamma = 0, K = delta_1 I
// so that  A_ii = sum_q JxW [ alpha phi_i^2 + (beta + gamma) . grad phi_i phi_i + grad phi_i . K grad phi_i ].
// With phi_i a tensor product this is a sum-factorised contraction of 1 + dim + dim (dim + 1) / 2 coefficient
// fields with the 1-D matrices S^2, S G, G^2: O(10 n^4) flops per component and cell instead of the
// (dim + 1) n^dim full cell applications of MatrixFreeTools::compute_diagonal (operator_ns.cc:210-218), same
// result up to round-off.  Cells with weighted constraint rows go through k_diag_columns.
template <int dim, int n, typename T, int BR>
__global__ void __launch_bounds__(Geo<dim, n>::THREADS) k_diag_sumfac(const KParams<T> p, const Shape<T, n> sh,
                                                                     const uint8_t *__restrict__ skip_cell)
{
  using G             = Geo<dim, n>;
  constexpr int C     = G::C;
  constexpr int NTERM = 1 + dim + dim * (dim + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem[];
  Ctx<dim, n, T> ctx;
  ctx.init(smem, sh);
  // 1-D products of the basis with itself, [q * n + i]; they live behind v in the (unused) gradient scratch
  T *mSS = ctx.sg, *mSG = mSS + n * n, *mGG = mSG + n * n;
  for (int k = threadIdx.x; k < n * n; k += blockDim.x)
    {
      mSS[k] = sh.S[k] * sh.S[k];
      mSG[k] = sh.S[k] * sh.G[k];
      mGG[k] = sh.G[k] * sh.G[k];
    }
  const uint32_t cell0  = p.cell_begin + blockIdx.x * G::CPB + ctx.cb;
  const bool     active = cell_active(p, cell0);
  const uint32_t cell   = cell0 < p.cell_end ? cell0 : p.cell_end - 1;
  GeomQ<dim, n, T> geo;
  geo.load(p, sh, cell, ctx.l, ctx.ii);
  QTables<dim, T> tb;
  load_tables<dim, T, BR>(p, qpos(p, (uint32_t)ctx.l, cell), cell, tb);
  // (J^-1)_{e m}, dense
  T ji[dim][dim];
#pragma unroll
  for (int e = 0; e < dim; ++e)
#pragma unroll
    for (int m = 0; m < dim; ++m)
      ji[e][m] = geo.cart ? (e == m ? geo.ij[e][e] : T(0)) : geo.ij[e][m];

  // coefficient fields of this quadrature point, per component: [A | b_e | K_ee | 2 K_ef (e < f)]
  T coef[C][NTERM];
  {
    const T w = p.weight, nu = p.nu, d1 = tb.d1, d2 = tb.d2;
    const T th = (BR == BR_NEWTON) ? T(1) : p.theta;
    const T tw = p.ctd ? w : T(0);
#pragma unroll
    for (int c = 0; c < C; ++c)
      {
        T alpha = 0, bg[dim], K[dim][dim];
        if (c < dim)
          {
            T Hcc = 0, r1 = 0;
            if (BR == BR_NEWTON)
              {
                Hcc   = tb.H[c][c];
                T sgs = 0;
#pragma unroll
                for (int j = 0; j < dim; ++j)
                  sgs += tb.H[c][j] * tb.U[j];
                T b = tb.P[c] + sgs;
                if (p.ctd)
                  b = (tb.U[c] * w + tb.O[c]) + b;
                r1 = d1 * b;
              }
            alpha = w + Hcc;
#pragma unroll
            for (int j = 0; j < dim; ++j)
              {
                bg[j] = th * tb.U[j] + tb.U[j] * d1 * (tw + Hcc) + (j == c ? r1 : T(0));
#pragma unroll
                for (int m = 0; m < dim; ++m)
                  K[j][m] = (j == m ? nu * th : T(0)) + ((j == c && m == c) ? th * (nu + d2) : T(0)) +
                            d1 * th * tb.U[j] * tb.U[m];
              }
          }
        else
          {
#pragma unroll
            for (int j = 0; j < dim; ++j)
              {
                bg[j] = 0;
#pragma unroll
                for (int m = 0; m < dim; ++m)
                  K[j][m] = (j == m) ? d1 : T(0);
              }
          }
        // to the reference cell: b_e = sum_m J^-1_{e m} bg_m, Khat = J^-1 K J^-T, all times JxW
        coef[c][0] = geo.jxw * alpha;
        T kj[dim][dim]; // K J^-T: kj[j][f] = sum_m K[j][m] ji[f][m]
#pragma unroll
        for (int j = 0; j < dim; ++j)
#pragma unroll
          for (int f = 0; f < dim; ++f)
            {
              T s = 0;
#pragma unroll
              for (int m = 0; m < dim; ++m)
                s += K[j][m] * ji[f][m];
              kj[j][f] = s;
            }
        int t = 1 + dim;
#pragma unroll
        for (int e = 0; e < dim; ++e)
          {
            T s = 0, kee = 0;
#pragma unroll
            for (int m = 0; m < dim; ++m)
              {
                s += ji[e][m] * bg[m];
                kee += ji[e][m] * kj[m][e];
              }
            coef[c][1 + e]       = geo.jxw * s;
            coef[c][1 + dim + e] = geo.jxw * kee;
          }
        t = 1 + 2 * dim;
#pragma unroll
        for (int e = 0; e < dim; ++e)
#pragma unroll
          for (int f = e + 1; f < dim; ++f)
            {
              T s = 0;
#pragma unroll
              for (int m = 0; m < dim; ++m)
                s += ji[e][m] * kj[m][f];
              coef[c][t++] = T(2) * geo.jxw * s;
            }
      }
  }

  // contract every field with its triple of 1-D matrices (transposed sweeps of all C components at once)
  T diag[C];
#pragma unroll
  for (int c = 0; c < C; ++c)
    diag[c] = 0;
  __syncthreads(); // matrices written
#pragma unroll
  for (int t = 0; t < NTERM; ++t)
    {
      // which direction carries a derivative factor (0 = none) once or twice
      int de = -1, df = -1;
      if (t >= 1 && t <= dim)
        de = t - 1;
      else if (t > dim && t <= 2 * dim)
        de = df = t - 1 - dim;
      else if (t > 2 * dim)
        {
          int idx = t - 1 - 2 * dim, cnt = 0;
#pragma unroll
          for (int e = 0; e < dim; ++e)
#pragma unroll
            for (int f = e + 1; f < dim; ++f)
              {
                if (cnt == idx)
                  {
                    de = e;
                    df = f;
                  }
                ++cnt;
              }
        }
#pragma unroll
      for (int c = 0; c < C; ++c)
        ctx.v[ctx.at(c, ctx.l)] = coef[c][t];
      __syncthreads();
#pragma unroll
      for (int e = 0; e < dim; ++e)
        {
          const int cnt = (de == e ? 1 : 0) + (df == e ? 1 : 0);
          ctx.sweep(cnt == 0 ? mSS : (cnt == 1 ? mSG : mGG), e, true);
        }
#pragma unroll
      for (int c = 0; c < C; ++c)
        diag[c] += ctx.v[ctx.at(c, ctx.l)];
      __syncthreads();
    }
  if (!active || (skip_cell != nullptr && skip_cell[cell]))
    return;
#pragma unroll
  for (int c = 0; c < C; ++c)
    {
      const uint32_t iv = p.idx[idx_at(p, c * G::n_loc + ctx.l, cell)];
      if (!(iv & GLSB_CONSTRAINED_BIT))
        atomic_add(p.dst + iv, diag[c]);
    }
}

// diag(C_cell^T A_cell C_cell) for cells with weighted constraint rows: one CTA handles CPB
// slots of the same cell list entry; for every global column g of the cell, x = C_cell[:, g],
// y = A_cell x, diag[g] += x . y.  Column data is CSR built on the host at create time.
template <int dim, int n, typename T, int BR>
__global__ void __launch_bounds__(Geo<dim, n>::THREADS) k_diag_columns(const KParams<T> p, const Shape<T, n> sh,
                                                                      const DiagColumns dc)
{
  using G         = Geo<dim, n>;
  constexpr int C = G::C;
  extern __shared__ __align__(16) unsigned char smem[];
  Ctx<dim, n, T> ctx;
  ctx.init(smem, sh);
  // all CPB slots of the CTA work on the same cell, slot cb takes columns cb, cb + CPB, ...
  const uint32_t li   = blockIdx.x;
  const uint32_t cell = dc.cell[li];
  const uint32_t c0 = dc.col_ptr[li], c1 = dc.col_ptr[li + 1];
  const uint32_t nrounds = (c1 - c0 + G::CPB - 1) / G::CPB;
  GeomQ<dim, n, T> geo;
  geo.load(p, sh, cell, ctx.l, ctx.ii);
  QTables<dim, T> tb;
  load_tables<dim, T, BR>(p, qpos(p, (uint32_t)ctx.l, cell), cell, tb);
  for (uint32_t r = 0; r < nrounds; ++r)
    {
      const uint32_t col = c0 + r * G::CPB + ctx.cb;
      const bool     ok  = col < c1;
      __syncthreads();
      T x[C];
#pragma unroll
      for (int c = 0; c < C; ++c)
        {
          x[c] = 0;
          if (ok)
            for (uint32_t e = dc.ent_ptr[col]; e < dc.ent_ptr[col + 1]; ++e)
              if (dc.ent_loc[e] == (uint32_t)(c * G::n_loc + ctx.l))
                x[c] += T(dc.ent_val[e]);
          ctx.v[ctx.at(c, ctx.l)] = x[c];
        }
      __syncthreads();
      cell_apply<dim, n, T, BR>(ctx, p, geo, tb);
      T s = 0;
#pragma unroll
      for (int c = 0; c < C; ++c)
        s += x[c] * ctx.v[ctx.at(c, ctx.l)];
      if (ok && s != T(0))
        atomic_add(p.dst + dc.col_dof[col], s);
    }
}

} // namespace glsb
