// Boundary-face outflow terms: do_vmult_boundary (operator_ns.cc:1195-1301) and the face tables of
// compute_penalty_parameters (:423-521).  Faces of all_outflow_bcs_cut carry (v, beta min(0, U.n) u), faces of
// all_outflow_bcs_nitsche carry (v, beta u) - nu (v, grad u . n) - nu (grad v . n, u) (with u - u_target in the
// residual).  The reference runs them as the boundary lambda of MatrixFree::loop (:710-717); here one CTA handles
// one face after the cell kernel of the same launch sequence.  Boundary faces are O(N^(2/3)) of the work, so
// these kernels are written for clarity: full tensor-product basis at the face points, no sum factorisation.
#pragma once
#include "glsb_kernels.cuh"

namespace glsb
{
template <typename T>
struct FaceParams
{
  uint32_t        nf;
  int             n, nloc, dim, nqf;
  const uint32_t *slot, *no, *kind; // internal cell slot, 2 * direction + side, 1 = cut / 2 = Nitsche
  const T        *normal;           // [f][q][dim]
  const T        *jxw;              // [f][q]
  const T        *inv_jac;          // [f][q][e][j] = (J^-1)_{e j}
  const T        *target;           // [f][q][dim] (Nitsche residual)
  const T        *beta;             // [f] effective_beta_face
  T              *velocity;         // [f][q][dim] face_velocity (cut)
  T               nu;
  T S[MAX_N * MAX_N], G[MAX_N * MAX_N]; // values / derivatives of the 1-D basis at the Gauss points [q * n + i]
  T Nf[2 * MAX_N], Gf[2 * MAX_N];       // ... at xi = 0 and xi = 1 [side * n + i]
};

// value and reference gradient of basis function i at quadrature point q of face `no`
template <typename T, int dim>
__device__ __forceinline__ void face_basis(const FaceParams<T> &f, int no, int q, int i, T &phi, T (&dphi)[dim])
{
  const int n = f.n, dir = no >> 1, side = no & 1;
  T         v[dim], g[dim];
  int       qq = q, ii = i;
#pragma unroll
  for (int e = 0; e < dim; ++e)
    {
      const int ie = ii % n;
      ii /= n;
      if (e == dir)
        {
          v[e] = f.Nf[side * n + ie];
          g[e] = f.Gf[side * n + ie];
        }
      else
        {
          const int qe = qq % n;
          qq /= n;
          v[e] = f.S[qe * n + ie];
          g[e] = f.G[qe * n + ie];
        }
    }
  phi = v[0];
#pragma unroll
  for (int e = 1; e < dim; ++e)
    phi *= v[e];
#pragma unroll
  for (int e = 0; e < dim; ++e)
    {
      T d = g[e];
#pragma unroll
      for (int a = 0; a < dim; ++a)
        if (a != e)
          d *= v[a];
      dphi[e] = d;
    }
}

constexpr int FACE_THREADS = 128;

template <int dim>
constexpr size_t face_smem_elems(int n)
{
  const int nloc = dim == 2 ? n * n : n * n * n, nqf = dim == 2 ? n : n * n;
  return (size_t)dim * nloc + (size_t)nqf * dim + (size_t)nqf * dim * dim;
}

// dst += face terms applied to src (RES: evaluate_residual variant, read_dof_values_plain, sign of p.sign_negative)
template <typename T, int dim, bool RES>
__global__ void __launch_bounds__(FACE_THREADS) k_faces_apply(const KParams<T> p, const FaceParams<T> f)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n_loc = f.nloc, nqf = f.nqf;
  T        *u  = reinterpret_cast<T *>(smem_raw); // [dim][nloc]
  T        *vq = u + dim * n_loc;                 // [nqf][dim]
  T        *gq = vq + nqf * dim;                  // [nqf][dim][dim] reference gradient to test with
  const uint32_t face = blockIdx.x, slot = f.slot[face];
  const int      no = (int)f.no[face], kind = (int)f.kind[face];
  for (int d = threadIdx.x; d < dim * n_loc; d += blockDim.x)
    {
      const uint32_t iv = p.idx[idx_at(p, (uint32_t)d, slot)];
      u[d]              = RES ? p.src[plain_index(p, iv)] : gather_resolved(p, p.src, iv);
    }
  __syncthreads();
  const T beta = f.beta[face];
  for (int q = threadIdx.x; q < nqf; q += blockDim.x)
    {
      T val[dim], rg[dim][dim];
#pragma unroll
      for (int c = 0; c < dim; ++c)
        {
          val[c] = 0;
#pragma unroll
          for (int e = 0; e < dim; ++e)
            rg[c][e] = 0;
        }
      for (int i = 0; i < n_loc; ++i)
        {
          T phi, dphi[dim];
          face_basis<T, dim>(f, no, q, i, phi, dphi);
#pragma unroll
          for (int c = 0; c < dim; ++c)
            {
              const T ui = u[c * n_loc + i];
              val[c] += phi * ui;
#pragma unroll
              for (int e = 0; e < dim; ++e)
                rg[c][e] += dphi[e] * ui;
            }
        }
      const uint64_t fq = (uint64_t)face * nqf + q;
      T              nrm[dim], ij[dim][dim];
#pragma unroll
      for (int j = 0; j < dim; ++j)
        nrm[j] = f.normal[fq * dim + j];
#pragma unroll
      for (int e = 0; e < dim; ++e)
#pragma unroll
        for (int j = 0; j < dim; ++j)
          ij[e][j] = f.inv_jac[(fq * dim + e) * dim + j];
      const T jxw = f.jxw[fq];
      T       vr[dim], gr[dim][dim];
      if (kind == 1)
        {
          // (v, beta min(0, u* . n) u), operator_ns.cc:1213-1245
          T no_flux = 0;
#pragma unroll
          for (int j = 0; j < dim; ++j)
            no_flux += (RES ? val[j] : f.velocity[fq * dim + j]) * nrm[j];
          no_flux = no_flux < T(0) ? no_flux : T(0);
#pragma unroll
          for (int c = 0; c < dim; ++c)
            {
              vr[c] = beta * no_flux * val[c];
#pragma unroll
              for (int j = 0; j < dim; ++j)
                gr[c][j] = 0;
            }
        }
      else
        {
          // Nitsche outflow, operator_ns.cc:1247-1291
          if (RES)
#pragma unroll
            for (int c = 0; c < dim; ++c)
              val[c] -= f.target[fq * dim + c];
#pragma unroll
          for (int c = 0; c < dim; ++c)
            {
              T gn = 0;
#pragma unroll
              for (int j = 0; j < dim; ++j)
                {
                  T g = 0;
#pragma unroll
                  for (int e = 0; e < dim; ++e)
                    g += ij[e][j] * rg[c][e];
                  gn += g * nrm[j];
                  gr[c][j] = -f.nu * val[c] * nrm[j];
                }
              vr[c] = beta * val[c] - f.nu * gn;
            }
        }
#pragma unroll
      for (int c = 0; c < dim; ++c)
        {
          vq[q * dim + c] = vr[c] * jxw;
#pragma unroll
          for (int e = 0; e < dim; ++e)
            {
              T s = 0;
#pragma unroll
              for (int j = 0; j < dim; ++j)
                s += ij[e][j] * gr[c][j];
              gq[(q * dim + c) * dim + e] = s * jxw;
            }
        }
    }
  __syncthreads();
  for (int d = threadIdx.x; d < dim * n_loc; d += blockDim.x)
    {
      const int c = d / n_loc, i = d - c * n_loc;
      T         s = 0;
      for (int q = 0; q < nqf; ++q)
        {
          T phi, dphi[dim];
          face_basis<T, dim>(f, no, q, i, phi, dphi);
          s += vq[q * dim + c] * phi;
#pragma unroll
          for (int e = 0; e < dim; ++e)
            s += gq[(q * dim + c) * dim + e] * dphi[e];
        }
      if (p.sign_negative)
        s = -s;
      if (s != T(0))
        scatter_resolved(p, p.dst, p.idx[idx_at(p, (uint32_t)d, slot)], s);
    }
}

// face_velocity[face][q] = value of the linearization point (read_dof_values_plain), operator_ns.cc:460-476
template <typename T, int dim>
__global__ void __launch_bounds__(FACE_THREADS) k_faces_velocity(const KParams<T> p, const FaceParams<T> f)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n_loc = f.nloc, nqf = f.nqf;
  T        *u = reinterpret_cast<T *>(smem_raw);
  const uint32_t face = blockIdx.x, slot = f.slot[face];
  const int      no = (int)f.no[face];
  for (int d = threadIdx.x; d < dim * n_loc; d += blockDim.x)
    u[d] = p.src[plain_index(p, p.idx[idx_at(p, (uint32_t)d, slot)])];
  __syncthreads();
  for (int q = threadIdx.x; q < nqf; q += blockDim.x)
    {
      T val[dim];
#pragma unroll
      for (int c = 0; c < dim; ++c)
        val[c] = 0;
      for (int i = 0; i < n_loc; ++i)
        {
          T phi, dphi[dim];
          face_basis<T, dim>(f, no, q, i, phi, dphi);
#pragma unroll
          for (int c = 0; c < dim; ++c)
            val[c] += phi * u[c * n_loc + i];
        }
#pragma unroll
      for (int c = 0; c < dim; ++c)
        f.velocity[((uint64_t)face * nqf + q) * dim + c] = val[c];
    }
}

// diagonal of the face terms (the boundary lambda of MatrixFreeTools::compute_diagonal, operator_ns.cc:210-218):
// A_ii += sum_q JxW [coef phi_i^2 - 2 nu phi_i d_n phi_i (Nitsche only)], coef = beta min(0, U.n) or beta.
// Dofs with constraint rows are skipped (zero rows contribute nothing; weighted rows on outflow faces are not
// supported in the diagonal).
template <typename T, int dim>
__global__ void __launch_bounds__(FACE_THREADS) k_faces_diag(const KParams<T> p, const FaceParams<T> f)
{
  const int      n_loc = f.nloc, nqf = f.nqf;
  const uint32_t face = blockIdx.x, slot = f.slot[face];
  const int      no = (int)f.no[face], kind = (int)f.kind[face];
  const T        beta = f.beta[face];
  for (int d = threadIdx.x; d < dim * n_loc; d += blockDim.x)
    {
      const uint32_t iv = p.idx[idx_at(p, (uint32_t)d, slot)];
      if (iv & GLSB_CONSTRAINED_BIT)
        continue;
      const int i = d % n_loc;
      T         s = 0;
      for (int q = 0; q < nqf; ++q)
        {
          T phi, dphi[dim];
          face_basis<T, dim>(f, no, q, i, phi, dphi);
          if (phi == T(0))
            continue;
          const uint64_t fq = (uint64_t)face * nqf + q;
          T              coef = beta, dn = 0;
          if (kind == 1)
            {
              T no_flux = 0;
#pragma unroll
              for (int j = 0; j < dim; ++j)
                no_flux += f.velocity[fq * dim + j] * f.normal[fq * dim + j];
              coef = beta * (no_flux < T(0) ? no_flux : T(0));
            }
          else
            {
#pragma unroll
              for (int j = 0; j < dim; ++j)
                {
                  T g = 0;
#pragma unroll
                  for (int e = 0; e < dim; ++e)
                    g += f.inv_jac[(fq * dim + e) * dim + j] * dphi[e];
                  dn += g * f.normal[fq * dim + j];
                }
            }
          s += f.jxw[fq] * (coef * phi * phi - (kind == 2 ? T(2) * f.nu * phi * dn : T(0)));
        }
      if (s != T(0))
        atomic_add(p.dst + iv, s);
    }
}
} // namespace glsb
