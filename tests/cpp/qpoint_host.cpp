// The PRODUCT's quadrature-point physics on the host: the stretch of dealii_ns_gls_b200/csrc/glsb_kernels.cuh between
// the banners "quadrature-point physics" and "kernels" (symm_add, QTables, load_tables, qpoint_physics -- what the
// generic, column, diagonal and residual kernels execute per point) is cut out at build time
// (tests/cpp/build_qpoint_host.sh) and compiled with g++, where __device__ / __forceinline__ are plain
// attributes, into tests/cpp/libprod_qpoint.so.  tests/test_reference_qpoint.py feeds it the inputs of the
// reference record (tests/golden/reference/qpoint.npz = output of the reference's own do_vmult_cell) and compares:
// CUDA source against reference object code without a GPU and without the oracle in between.
// (The register-tiled Q2 kernel carries its own copy of the physics spread over four lanes per cell; it is
// compared with this code path on the GPU, tests/test_gpu_parity.py::test_q2_fast_kernel_matches_generic_and_oracle.)
#include "glsb_common.h"

namespace glsb
{
#include "qpoint_product_extract.inc"

  template <int dim, int BR, typename T = double>
  static void
  run(const double theta, const double nu, const double weight, const int ctd, const int has_o, const int n_q,
      const double *value, const double *grad, const double *u_star, const double *u_star_grad,
      const double *p_star_grad, const double *u_tdo, const double *u_old_grad, const double *p_old_grad,
      const double *d1, const double *d2, const int cell_wise, double *value_out, double *grad_out)
  {
    constexpr int   C = dim + 1;
    KParams<T> p{};
    p.weight     = T(weight);
    p.nu         = T(nu);
    p.theta      = T(theta);
    p.ctd        = ctd;
    p.has_o      = has_o;
    p.theta_ne_1 = theta != 1.0;
    p.cell_wise  = cell_wise;
    for (int q = 0; q < n_q; ++q)
      {
        QTables<dim, T> tb{};
        tb.d1 = T(cell_wise ? d1[0] : d1[q]);
        tb.d2 = T(cell_wise ? d2[0] : d2[q]);
        for (int i = 0; i < dim; ++i)
          {
            tb.U[i]      = T(u_star[q * dim + i]);
            tb.P[i]      = T(p_star_grad[q * dim + i]);
            tb.O[i]      = (u_tdo && (BR == BR_NEWTON ? ctd : (BR == BR_RESIDUAL && has_o))) ? T(u_tdo[q * dim + i]) : T(0);
            tb.gold_p[i] = (BR == BR_RESIDUAL && p.theta_ne_1) ? T(p_old_grad[q * dim + i]) : T(0);
            for (int j = 0; j < dim; ++j)
              {
                tb.H[i][j]    = T(u_star_grad[(q * dim + i) * dim + j]);
                tb.Gold[i][j] = (BR == BR_RESIDUAL && p.theta_ne_1) ? T(u_old_grad[(q * dim + i) * dim + j]) : T(0);
              }
          }
        T val[C], g[C][dim], vout[C], gout[C][dim];
        for (int c = 0; c < C; ++c)
          {
            val[c] = T(value[q * C + c]);
            for (int j = 0; j < dim; ++j)
              g[c][j] = T(grad[(q * C + c) * dim + j]);
          }
        qpoint_physics<dim, T, BR>(p, tb, val, g, vout, gout);
        for (int c = 0; c < C; ++c)
          {
            value_out[q * C + c] = vout[c];
            for (int j = 0; j < dim; ++j)
              grad_out[(q * C + c) * dim + j] = gout[c][j];
          }
      }
  }
} // namespace glsb

// branch: 0 Newton, 1 fixed point, 2 residual (glsb::Branch).  The table entries are filled the way load_tables
// fills them for that branch (entries the kernel does not load are zero).
// is_float != 0: T = float (the level operators)
extern "C" int
prod_qpoint(int is_float, int dim, int branch, double theta, double nu, double weight, int ctd, int has_o, int n_q,
            const double *value, const double *grad, const double *u_star, const double *u_star_grad,
            const double *p_star_grad, const double *u_tdo, const double *u_old_grad, const double *p_old_grad,
            const double *d1, const double *d2, int cell_wise, double *value_out, double *grad_out)
{
#define GO(D, B)                                                                                                   \
  if (dim == D && branch == B)                                                                                     \
    {                                                                                                              \
      if (is_float)                                                                                                \
        glsb::run<D, B, float>(theta, nu, weight, ctd, has_o, n_q, value, grad, u_star, u_star_grad, p_star_grad,  \
                               u_tdo, u_old_grad, p_old_grad, d1, d2, cell_wise, value_out, grad_out);             \
      else                                                                                                         \
        glsb::run<D, B>(theta, nu, weight, ctd, has_o, n_q, value, grad, u_star, u_star_grad, p_star_grad, u_tdo,  \
                        u_old_grad, p_old_grad, d1, d2, cell_wise, value_out, grad_out);                           \
      return 0;                                                                                                    \
    }
  GO(2, 0) GO(2, 1) GO(2, 2) GO(3, 0) GO(3, 1) GO(3, 2)
#undef GO
  return 1;
}
