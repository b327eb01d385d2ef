// 1-D tables of FE_Q(p) (Lagrange basis on the Gauss-Lobatto points of [0,1]) evaluated at the
// points of QGauss(p+1), what MatrixFree's ShapeInfo holds for the reference's
// FESystem(FE_Q(p), dim+1) / QGauss(p+1) pair (main.cc:251, performance.cc:33-38).
// Everything is computed in long double and rounded once.
#include "glsb_common.h"

#include <cmath>
#include <vector>

namespace glsb
{
typedef long double ld;

static ld legendre(int n, ld x, ld *dp)
{
  ld p0 = 1, p1 = x;
  if (n == 0)
    {
      if (dp)
        *dp = 0;
      return 1;
    }
  for (int k = 2; k <= n; ++k)
    {
      ld p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
      p0    = p1;
      p1    = p2;
    }
  if (dp)
    *dp = n * (x * p1 - p0) / (x * x - 1);
  return p1;
}

static void gauss(int n, ld *x, ld *w)
{
  const ld pi = acosl(-1.0L);
  for (int i = 0; i < n; ++i)
    {
      ld z = -cosl(pi * (i + 0.75L) / (n + 0.5L));
      for (int it = 0; it < 100; ++it)
        {
          ld dp, p = legendre(n, z, &dp);
          ld dz = p / dp;
          z -= dz;
          if (fabsl(dz) < 1e-19L)
            break;
        }
      ld dp;
      legendre(n, z, &dp);
      x[i] = z;
      w[i] = 2 / ((1 - z * z) * dp * dp);
    }
  for (int i = 0; i < n / 2; ++i) // symmetrise
    {
      ld a         = 0.5L * (x[n - 1 - i] - x[i]);
      x[i]         = -a;
      x[n - 1 - i] = a;
      ld b         = 0.5L * (w[i] + w[n - 1 - i]);
      w[i] = w[n - 1 - i] = b;
    }
  if (n % 2)
    x[n / 2] = 0;
}

static void gauss_lobatto(int p, ld *x)
{
  // roots of P'_p plus the end points
  const ld pi = acosl(-1.0L);
  x[0]        = -1;
  x[p]        = 1;
  for (int i = 1; i < p; ++i)
    {
      ld z = -cosl(pi * i / p);
      for (int it = 0; it < 100; ++it)
        {
          // f = P'_p(z); f' from the Legendre ODE: (1-z^2) P'' = 2 z P' - p(p+1) P
          ld dp, pv = legendre(p, z, &dp);
          ld ddp = (2 * z * dp - p * (p + 1) * pv) / (1 - z * z);
          ld dz  = dp / ddp;
          z -= dz;
          if (fabsl(dz) < 1e-19L)
            break;
        }
      x[i] = z;
    }
  for (int i = 0; i <= p / 2; ++i)
    {
      ld a     = 0.5L * (x[p - i] - x[i]);
      x[i]     = -a;
      x[p - i] = a;
    }
}

static void lagrange(int m, const ld *nodes, ld x, ld *val, ld *der)
{
  for (int i = 0; i < m; ++i)
    {
      ld v = 1;
      for (int a = 0; a < m; ++a)
        if (a != i)
          v *= (x - nodes[a]) / (nodes[i] - nodes[a]);
      val[i] = v;
      ld d   = 0;
      for (int l = 0; l < m; ++l)
        {
          if (l == i)
            continue;
          ld t = 1 / (nodes[i] - nodes[l]);
          for (int a = 0; a < m; ++a)
            if (a != i && a != l)
              t *= (x - nodes[a]) / (nodes[i] - nodes[a]);
          d += t;
        }
      der[i] = d;
    }
}

void compute_shape_host(int degree, ShapeHost &out)
{
  const int n = degree + 1;
  out.n       = n;
  ld nodes[MAX_N], xq[MAX_N], wq[MAX_N];
  gauss_lobatto(degree, nodes);
  gauss(n, xq, wq);
  for (int i = 0; i < n; ++i)
    {
      nodes[i] = 0.5L * (nodes[i] + 1);
      xq[i]    = 0.5L * (xq[i] + 1);
      wq[i]    = 0.5L * wq[i];
      out.nodes[i] = (double)nodes[i];
      out.xq[i]    = (double)xq[i];
      out.w[i]     = (double)wq[i];
    }
  for (int q = 0; q < n; ++q)
    {
      ld v[MAX_N], d[MAX_N];
      lagrange(n, nodes, xq[q], v, d);
      for (int i = 0; i < n; ++i)
        {
          out.S[q * n + i] = (double)v[i];
          out.G[q * n + i] = (double)d[i];
        }
      lagrange(n, xq, xq[q], v, d);
      for (int i = 0; i < n; ++i)
        out.D[q * n + i] = (double)d[i];
    }
}

// 1-D matrices of the two-level transfer between a cell and its two children per direction
// (what deal.II's MGTwoLevelTransfer takes from FE_Q::get_prolongation_matrix / get_restriction_matrix):
//   P[a][l][j] = phi^coarse_j((x_l + a) / 2)              value of coarse basis j at node l of child a
//   R[j][l]    = phi^fine_l(2 x_j - rchild[j])            interpolation: fine function of child rchild[j]
//                                                         evaluated at coarse node j
void compute_transfer_host(int degree, double *P, double *R, int *rchild)
{
  const int n = degree + 1;
  ld        nodes[MAX_N];
  gauss_lobatto(degree, nodes);
  for (int i = 0; i < n; ++i)
    nodes[i] = 0.5L * (nodes[i] + 1);
  ld v[MAX_N], d[MAX_N];
  for (int a = 0; a < 2; ++a)
    for (int l = 0; l < n; ++l)
      {
        lagrange(n, nodes, (nodes[l] + a) / 2, v, d);
        for (int j = 0; j < n; ++j)
          P[(a * n + l) * n + j] = (fabsl(v[j]) < 1e-17L) ? 0.0 : (double)v[j];
      }
  for (int j = 0; j < n; ++j)
    {
      const int a = (nodes[j] <= 0.5L) ? 0 : 1; // x = 1/2 belongs to both children (same value: continuity)
      rchild[j]   = a;
      lagrange(n, nodes, 2 * nodes[j] - a, v, d);
      for (int l = 0; l < n; ++l)
        R[j * n + l] = (fabsl(v[l]) < 1e-17L) ? 0.0 : (double)v[l];
    }
}

// values Nf[side * n + i] and derivatives Gf[side * n + i] of the 1-D basis at the face positions xi = 0, 1
void compute_face_basis_host(int degree, double *Nf, double *Gf)
{
  const int n = degree + 1;
  ld        nodes[MAX_N], v[MAX_N], d[MAX_N];
  gauss_lobatto(degree, nodes);
  for (int i = 0; i < n; ++i)
    nodes[i] = 0.5L * (nodes[i] + 1);
  for (int side = 0; side < 2; ++side)
    {
      lagrange(n, nodes, (ld)side, v, d);
      for (int i = 0; i < n; ++i)
        {
          Nf[side * n + i] = (fabsl(v[i]) < 1e-17L) ? 0.0 : (double)v[i];
          Gf[side * n + i] = (double)d[i];
        }
    }
}

} // namespace glsb
