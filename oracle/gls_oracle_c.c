/* CPU restatement (C, sum-factorised, SIMD over cell batches, OpenMP) of the cell loop of
 * NavierStokesOperator::vmult / evaluate_residual.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ as a second, independently written checker next
 * to oracle/gls_oracle.py, and by bench.py as the timed CPU baseline ("port": this is a
 * restatement of the reference algorithm, not deal.II).  Pinned through gls_oracle.py, with
 * which it agrees to round-off: that file's quadrature-point physics is compared with the
 * reference's own object code (oracle/_ref), deal.II's parts stay unpinned (see its header).
 *
 * Follows /root/reference:
 *   include/operator_ns.cc:806-830    do_vmult_range: gather, cell kernel, scatter-add
 *   include/operator_ns.cc:949-1066   fixed-point / residual branch
 *   include/operator_ns.cc:1067-1182  Newton branch
 *   include/operator_ns.cc:899-916    symm_scalar_product_add
 * and mimics how deal.II executes it: VectorizedArray over LANES cells, even cells per
 * batch, q-point tables stored per batch, 1-D sweeps with the collocation derivative.
 *
 * Plain (unconstrained) gather/scatter only: callers resolve constraints around it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LANES 8
#define MAXN 5
#define MAXLOC 125

typedef struct
{
  int       dim, n, C, n_loc, branch, ctd, cell_wise, has_o, theta_ne_1, cartesian;
  int64_t   n_cells, n_batches, n_dofs;
  double    nu, theta;
  double    S[MAXN * MAXN], D[MAXN * MAXN], w[MAXN];
  uint32_t *idx;   /* [batch][C*n_loc][LANES] */
  double   *U, *H, *P, *O, *Gold, *gold_p; /* [batch][field][q][LANES] */
  double   *d1, *d2;                         /* cell-wise: [batch][LANES]; q-wise: [batch][q][LANES] */
  double   *ij, *jxw;                        /* cart: [batch][dim][LANES], [batch][LANES]; general: [batch][dim*dim][q][LANES], [batch][q][LANES] */
  uint8_t  *shared; /* [n_dofs] 1 if the dof is touched by more than one thread chunk */
  int       n_threads_prepared;
} glso_t;

static double *interleave(const double *src, int64_t n_cells, int64_t nb, int per_cell)
{
  /* src[cell][per_cell] -> out[batch][per_cell][LANES]; padding lanes repeat the last cell */
  double *out = (double *)malloc(sizeof(double) * nb * per_cell * LANES);
  for (int64_t b = 0; b < nb; ++b)
    for (int f = 0; f < per_cell; ++f)
      for (int l = 0; l < LANES; ++l)
        {
          int64_t c = b * LANES + l;
          if (c >= n_cells)
            c = n_cells - 1;
          out[(b * per_cell + f) * LANES + l] = src ? src[c * per_cell + f] : 0.0;
        }
  return out;
}

glso_t *glso_create(int dim, int degree, int64_t n_cells, int64_t n_dofs, const uint32_t *cell_dofs,
                    const double *S, const double *D, const double *w, int cartesian, const double *inv_jac,
                    const double *jxw, double nu, double theta, int branch, int ctd, int cell_wise)
{
  glso_t *o    = (glso_t *)calloc(1, sizeof(glso_t));
  o->dim       = dim;
  o->n         = degree + 1;
  o->C         = dim + 1;
  o->n_loc     = dim == 2 ? o->n * o->n : o->n * o->n * o->n;
  o->n_cells   = n_cells;
  o->n_batches = (n_cells + LANES - 1) / LANES;
  o->n_dofs    = n_dofs;
  o->nu        = nu;
  o->theta     = theta;
  o->branch    = branch;
  o->ctd       = ctd;
  o->cell_wise = cell_wise;
  o->theta_ne_1 = theta != 1.0;
  o->cartesian = cartesian;
  memcpy(o->S, S, sizeof(double) * o->n * o->n);
  memcpy(o->D, D, sizeof(double) * o->n * o->n);
  memcpy(o->w, w, sizeof(double) * o->n);
  const int ndof = o->C * o->n_loc;
  o->idx         = (uint32_t *)malloc(sizeof(uint32_t) * o->n_batches * ndof * LANES);
  for (int64_t b = 0; b < o->n_batches; ++b)
    for (int j = 0; j < ndof; ++j)
      for (int l = 0; l < LANES; ++l)
        {
          int64_t c = b * LANES + l;
          if (c >= n_cells)
            c = n_cells - 1;
          o->idx[(b * ndof + j) * LANES + l] = cell_dofs[c * ndof + j];
        }
  if (cartesian)
    {
      o->ij  = interleave(inv_jac, n_cells, o->n_batches, dim);
      o->jxw = interleave(jxw, n_cells, o->n_batches, 1);
    }
  else
    {
      /* inv_jac[cell][q][e][j] -> [batch][e*dim+j][q][LANES] */
      const int nq = o->n_loc, dd = dim * dim;
      double   *t  = (double *)malloc(sizeof(double) * n_cells * dd * nq);
      for (int64_t c = 0; c < n_cells; ++c)
        for (int q = 0; q < nq; ++q)
          for (int f = 0; f < dd; ++f)
            t[(c * dd + f) * nq + q] = inv_jac[(c * nq + q) * dd + f];
      o->ij = interleave(t, n_cells, o->n_batches, dd * nq);
      free(t);
      o->jxw = interleave(jxw, n_cells, o->n_batches, nq);
    }
  return o;
}

/* tables given as [cell][field][q] (field-major per cell) */
void glso_set_tables(glso_t *o, const double *U, const double *H, const double *P, const double *O,
                     const double *Gold, const double *gold_p, const double *d1, const double *d2)
{
  const int d = o->dim, nq = o->n_loc;
  free(o->U), free(o->H), free(o->P), free(o->O), free(o->Gold), free(o->gold_p), free(o->d1), free(o->d2);
  o->U      = interleave(U, o->n_cells, o->n_batches, d * nq);
  o->H      = H ? interleave(H, o->n_cells, o->n_batches, d * d * nq) : NULL;
  o->P      = P ? interleave(P, o->n_cells, o->n_batches, d * nq) : NULL;
  o->O      = O ? interleave(O, o->n_cells, o->n_batches, d * nq) : NULL;
  o->Gold   = Gold ? interleave(Gold, o->n_cells, o->n_batches, d * d * nq) : NULL;
  o->gold_p = gold_p ? interleave(gold_p, o->n_cells, o->n_batches, d * nq) : NULL;
  o->has_o  = O != NULL;
  o->d1     = interleave(d1, o->n_cells, o->n_batches, o->cell_wise ? 1 : nq);
  o->d2     = interleave(d2, o->n_cells, o->n_batches, o->cell_wise ? 1 : nq);
}

void glso_destroy(glso_t *o)
{
  if (!o)
    return;
  free(o->idx), free(o->U), free(o->H), free(o->P), free(o->O), free(o->Gold), free(o->gold_p);
  free(o->d1), free(o->d2), free(o->ij), free(o->jxw), free(o->shared);
  free(o);
}

typedef double vd __attribute__((vector_size(LANES * sizeof(double)), aligned(8)));

/* apply a 1-D matrix along direction e to x[n_loc] (vector lanes), M[out*n+in] (or transposed) */
static inline __attribute__((always_inline)) void sweep(const int dim, const int n, const double *M, const int transpose,
                                                       const int e, const vd *in, vd *out, const int add)
{
  const int st = e == 0 ? 1 : (e == 1 ? n : n * n);
  const int n_loc = dim == 2 ? n * n : n * n * n;
  for (int base = 0; base < n_loc; ++base)
    {
      if ((base / st) % n != 0)
        continue;
      for (int q = 0; q < n; ++q)
        {
          vd s = add ? out[base + q * st] : (vd){0};
          for (int i = 0; i < n; ++i)
            s += (transpose ? M[i * n + q] : M[q * n + i]) * in[base + i * st];
          out[base + q * st] = s;
        }
    }
}

static inline __attribute__((always_inline)) void cell_batch(const glso_t *o, const int dim, const int n, int64_t b,
                                                            const double *src, vd *res /* [C*n_loc] */,
                                                            const double weight,
                                                            const vd *local_in /* [C*n_loc] or NULL: gather from src */)
{
  const int C = dim + 1, n_loc = dim == 2 ? n * n : n * n * n, nq = n_loc;
  vd        val[4][MAXLOC], tmp[MAXLOC], rg[4][3][MAXLOC];
  const uint32_t *ix = o->idx + b * C * n_loc * LANES;
  for (int c = 0; c < C; ++c)
    {
      if (local_in)
        memcpy(val[c], local_in + c * n_loc, sizeof(vd) * n_loc);
      else
        for (int i = 0; i < n_loc; ++i)
          for (int l = 0; l < LANES; ++l)
            val[c][i][l] = src[ix[(c * n_loc + i) * LANES + l]];
      /* evaluate: interpolate to q points, then collocation derivatives */
      for (int e = 0; e < dim; ++e)
        {
          sweep(dim, n, o->S, 0, e, val[c], tmp, 0);
          memcpy(val[c], tmp, sizeof(vd) * n_loc);
        }
      for (int e = 0; e < dim; ++e)
        sweep(dim, n, o->D, 0, e, val[c], rg[c][e], 0);
    }
  const double *Ub = o->U + b * dim * nq * LANES;
  const vd     *U_ = (const vd *)Ub;
  const vd     *H_ = o->H ? (const vd *)(o->H + b * dim * dim * nq * LANES) : NULL;
  const vd     *P_ = o->P ? (const vd *)(o->P + b * dim * nq * LANES) : NULL;
  const vd     *O_ = o->O ? (const vd *)(o->O + b * dim * nq * LANES) : NULL;
  const vd     *Go = o->Gold ? (const vd *)(o->Gold + b * dim * dim * nq * LANES) : NULL;
  const vd     *go = o->gold_p ? (const vd *)(o->gold_p + b * dim * nq * LANES) : NULL;
  const double  nu = o->nu, th = o->theta;
  for (int q = 0; q < nq; ++q)
    {
      vd ij[3][3], jxw;
      if (o->cartesian)
        {
          double wq = o->w[q % n] * o->w[(q / n) % n] * (dim == 3 ? o->w[q / (n * n)] : 1.0);
          for (int e = 0; e < dim; ++e)
            for (int j = 0; j < dim; ++j)
              ij[e][j] = (e == j) ? *(const vd *)(o->ij + (b * dim + e) * LANES) : (vd){0};
          jxw = *(const vd *)(o->jxw + b * LANES) * wq;
        }
      else
        {
          for (int e = 0; e < dim; ++e)
            for (int j = 0; j < dim; ++j)
              ij[e][j] = *(const vd *)(o->ij + ((b * dim * dim + e * dim + j) * nq + q) * LANES);
          jxw = *(const vd *)(o->jxw + (b * nq + q) * LANES);
        }
      vd g[4][3], v[4];
      for (int c = 0; c < C; ++c)
        {
          v[c] = val[c][q];
          for (int j = 0; j < dim; ++j)
            {
              vd s = {0};
              for (int e = 0; e < dim; ++e)
                s += ij[e][j] * rg[c][e][q];
              g[c][j] = s;
            }
        }
      const vd d1 = o->cell_wise ? *(const vd *)(o->d1 + b * LANES) : *(const vd *)(o->d1 + (b * nq + q) * LANES);
      const vd d2 = o->cell_wise ? *(const vd *)(o->d2 + b * LANES) : *(const vd *)(o->d2 + (b * nq + q) * LANES);
      vd       Uq[3], vout[4], gout[4][3];
      for (int j = 0; j < dim; ++j)
        Uq[j] = U_[j * nq + q];
      for (int c = 0; c < C; ++c)
        for (int j = 0; j < dim; ++j)
          gout[c][j] = (vd){0};
      if (o->branch == 0)
        {
          vd div = {0}, r0[3], r1[3];
          for (int c = 0; c < dim; ++c)
            div += g[c][c];
          for (int c = 0; c < dim; ++c)
            {
              vd sgu = {0}, ugs = {0}, sgs = {0};
              for (int j = 0; j < dim; ++j)
                {
                  sgu += g[c][j] * Uq[j];
                  ugs += H_[(c * dim + j) * nq + q] * v[j];
                  sgs += H_[(c * dim + j) * nq + q] * Uq[j];
                }
              vd td   = v[c] * weight;
              vout[c] = td + sgu + ugs;
              vd a = g[dim][c] + sgu + ugs, bb = P_[c * nq + q] + sgs;
              if (o->ctd)
                {
                  a  = td + a;
                  bb = (Uq[c] * weight + O_[c * nq + q]) + bb;
                }
              r0[c] = d1 * a;
              r1[c] = d1 * bb;
            }
          for (int c = 0; c < dim; ++c)
            gout[c][c] -= v[dim];
          for (int c = 0; c < dim; ++c)
            gout[c][c] += g[c][c] * (nu * 2.0);
          for (int e = 0; e < dim; ++e)
            for (int c = e + 1; c < dim; ++c)
              {
                vd t = (g[c][e] + g[e][c]) * (nu * 2.0 * 0.5);
                gout[c][e] += t;
                gout[e][c] += t;
              }
          for (int a = 0; a < dim; ++a)
            for (int bq = 0; bq < dim; ++bq)
              gout[a][bq] += Uq[bq] * r0[a] + v[bq] * r1[a];
          for (int c = 0; c < dim; ++c)
            gout[c][c] += d2 * div;
          vout[dim] = div;
          for (int j = 0; j < dim; ++j)
            gout[dim][j] = r0[j];
        }
      else
        {
          const int res = o->branch == 2;
          vd        B[3][3], pbar[3], td[3], divb = {0}, sgb[3];
          for (int c = 0; c < dim; ++c)
            {
              td[c]   = v[c] * weight;
              pbar[c] = th * g[dim][c];
              for (int j = 0; j < dim; ++j)
                B[c][j] = th * g[c][j];
            }
          if (res && o->has_o)
            for (int c = 0; c < dim; ++c)
              td[c] += O_[c * nq + q];
          if (res && o->theta_ne_1)
            for (int c = 0; c < dim; ++c)
              {
                pbar[c] += (1.0 - th) * go[c * nq + q];
                for (int j = 0; j < dim; ++j)
                  B[c][j] += (1.0 - th) * Go[(c * dim + j) * nq + q];
              }
          for (int c = 0; c < dim; ++c)
            divb += B[c][c];
          for (int c = 0; c < dim; ++c)
            {
              vd s = {0};
              for (int j = 0; j < dim; ++j)
                s += B[c][j] * Uq[j];
              sgb[c]  = s;
              vout[c] = td[c] + s;
            }
          for (int c = 0; c < dim; ++c)
            gout[c][c] -= v[dim];
          for (int c = 0; c < dim; ++c)
            gout[c][c] += B[c][c] * (nu * 2.0);
          for (int e = 0; e < dim; ++e)
            for (int c = e + 1; c < dim; ++c)
              {
                vd t = (B[c][e] + B[e][c]) * (nu * 2.0 * 0.5);
                gout[c][e] += t;
                gout[e][c] += t;
              }
          for (int a = 0; a < dim; ++a)
            {
              vd tdc = o->ctd ? td[a] : (vd){0};
              vd r0  = d1 * (tdc + pbar[a] + sgb[a]);
              for (int bq = 0; bq < dim; ++bq)
                gout[a][bq] += Uq[bq] * r0;
              gout[dim][a] = d1 * (tdc + g[dim][a] + sgb[a]);
            }
          for (int c = 0; c < dim; ++c)
            gout[c][c] += d2 * divb;
          vout[dim] = divb;
        }
      /* submit_value / submit_gradient */
      for (int c = 0; c < C; ++c)
        {
          val[c][q] = vout[c] * jxw;
          for (int e = 0; e < dim; ++e)
            {
              vd s = {0};
              for (int j = 0; j < dim; ++j)
                s += ij[e][j] * gout[c][j];
              rg[c][e][q] = s * jxw;
            }
        }
    }
  /* integrate */
  for (int c = 0; c < C; ++c)
    {
      for (int e = 0; e < dim; ++e)
        sweep(dim, n, o->D, 1, e, rg[c][e], val[c], 1);
      for (int e = 0; e < dim; ++e)
        {
          sweep(dim, n, o->S, 1, e, val[c], tmp, 0);
          memcpy(val[c], tmp, sizeof(vd) * n_loc);
        }
      memcpy(res + c * n_loc, val[c], sizeof(vd) * n_loc);
    }
}

#define DISPATCH(DIM, N)                       \
  if (o->dim == DIM && o->n == N)              \
    {                                          \
      cell_batch(o, DIM, N, b, src, res, weight, local_in); \
      return;                                  \
    }

static void cell_batch_dispatch(const glso_t *o, int64_t b, const double *src, vd *res, double weight,
                                const vd *local_in)
{
  DISPATCH(3, 3)
  DISPATCH(3, 2)
  DISPATCH(3, 4)
  DISPATCH(3, 5)
  DISPATCH(2, 2)
  DISPATCH(2, 3)
  DISPATCH(2, 4)
  DISPATCH(2, 5)
}

static void prepare_shared(glso_t *o, int nt)
{
  /* batches are split into nt contiguous chunks; mark dofs touched by more than one chunk */
  free(o->shared);
  o->shared            = (uint8_t *)calloc(o->n_dofs, 1);
  int32_t *owner       = (int32_t *)malloc(sizeof(int32_t) * o->n_dofs);
  const int ndof       = o->C * o->n_loc;
  for (int64_t i = 0; i < o->n_dofs; ++i)
    owner[i] = -1;
  for (int t = 0; t < nt; ++t)
    {
      int64_t b0 = o->n_batches * t / nt, b1 = o->n_batches * (t + 1) / nt;
      for (int64_t k = b0 * ndof * LANES; k < b1 * ndof * LANES; ++k)
        {
          uint32_t d = o->idx[k];
          if (owner[d] == -1)
            owner[d] = t;
          else if (owner[d] != t)
            o->shared[d] = 1;
        }
    }
  free(owner);
  o->n_threads_prepared = nt;
}

int glso_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* dst += sum_cells scatter(A_cell gather(src)); dst is NOT zeroed here unless zero_dst */
void glso_apply(glso_t *o, double *dst, const double *src, double weight, int zero_dst, int n_threads)
{
  const int ndof = o->C * o->n_loc;
#ifdef _OPENMP
  if (n_threads <= 0)
    n_threads = omp_get_max_threads();
#else
  n_threads = 1;
#endif
  if (o->n_threads_prepared != n_threads)
    prepare_shared(o, n_threads);
#pragma omp parallel num_threads(n_threads)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    if (zero_dst)
      {
        int64_t i0 = o->n_dofs * t / n_threads, i1 = o->n_dofs * (t + 1) / n_threads;
        memset(dst + i0, 0, sizeof(double) * (i1 - i0));
      }
#pragma omp barrier
    vd      res[4 * MAXLOC];
    int64_t b0 = o->n_batches * t / n_threads, b1 = o->n_batches * (t + 1) / n_threads;
    for (int64_t b = b0; b < b1; ++b)
      {
        cell_batch_dispatch(o, b, src, res, weight, NULL);
        const uint32_t *ix    = o->idx + b * ndof * LANES;
        const int       lanes = (b == o->n_batches - 1) ? (int)(o->n_cells - b * LANES) : LANES;
        for (int j = 0; j < ndof; ++j)
          for (int l = 0; l < lanes; ++l)
            {
              const uint32_t d = ix[j * LANES + l];
              if (o->shared[d])
                {
#pragma omp atomic
                  dst[d] += res[j][l];
                }
              else
                dst[d] += res[j][l];
            }
      }
  }
}

/* diag += sum_cells scatter(diag(A_cell)): the cell operator applied to each of the C * n^dim local unit vectors,
 * the way MatrixFreeTools::compute_diagonal does it for the reference (operator_ns.cc:210-218).  Plain scatter:
 * the caller sets the constrained rows to 1 (zero-type constraints only) and inverts (operator_ns.cc:219-224). */
void glso_diagonal(glso_t *o, double *diag, double weight, int zero_dst, int n_threads)
{
  const int ndof = o->C * o->n_loc;
#ifdef _OPENMP
  if (n_threads <= 0)
    n_threads = omp_get_max_threads();
#else
  n_threads = 1;
#endif
  if (o->n_threads_prepared != n_threads)
    prepare_shared(o, n_threads);
#pragma omp parallel num_threads(n_threads)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    if (zero_dst)
      {
        int64_t i0 = o->n_dofs * t / n_threads, i1 = o->n_dofs * (t + 1) / n_threads;
        memset(diag + i0, 0, sizeof(double) * (i1 - i0));
      }
#pragma omp barrier
    vd      res[4 * MAXLOC], unit[4 * MAXLOC], dl[4 * MAXLOC];
    int64_t b0 = o->n_batches * t / n_threads, b1 = o->n_batches * (t + 1) / n_threads;
    memset(unit, 0, sizeof(vd) * ndof);
    for (int64_t b = b0; b < b1; ++b)
      {
        for (int j = 0; j < ndof; ++j)
          {
            for (int l = 0; l < LANES; ++l)
              unit[j][l] = 1.0;
            cell_batch_dispatch(o, b, NULL, res, weight, unit);
            dl[j] = res[j];
            for (int l = 0; l < LANES; ++l)
              unit[j][l] = 0.0;
          }
        const uint32_t *ix    = o->idx + b * ndof * LANES;
        const int       lanes = (b == o->n_batches - 1) ? (int)(o->n_cells - b * LANES) : LANES;
        for (int j = 0; j < ndof; ++j)
          for (int l = 0; l < lanes; ++l)
            {
              const uint32_t d = ix[j * LANES + l];
              if (o->shared[d])
                {
#pragma omp atomic
                  diag[d] += dl[j][l];
                }
              else
                diag[d] += dl[j][l];
            }
      }
  }
}
