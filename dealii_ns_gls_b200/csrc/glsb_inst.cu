// One translation unit per (dim, Number): compile with -DGLSB_DIM=2|3 -DGLSB_REAL=double|float.
#include "glsb_kernels.cuh"
#include <cstdlib>
#if GLSB_DIM == 3
#include "glsb_col.cuh"
#endif

#ifndef GLSB_DIM
#error "define GLSB_DIM"
#endif
#ifndef GLSB_REAL
#error "define GLSB_REAL"
#endif

namespace glsb
{

template <typename K>
static int ensure_smem(K kernel, size_t bytes)
{
  if (bytes > 48 * 1024)
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
      return 1;
  return 0;
}

#define GLSB_GRID(G, p) (unsigned)(((p).cell_end - (p).cell_begin + G::CPB - 1) / G::CPB)

template <int dim, int n, typename T>
static int launch_vmult(int branch, const KParams<T> &p, const ShapeHost &sh, cudaStream_t s)
{
  using G = Geo<dim, n>;
  if (p.cell_end <= p.cell_begin)
    return 0;
  const size_t sm = generic_smem_bytes<dim, n, T>();
  const auto   S  = to_shape<T, n>(sh);
  switch (branch)
    {
      case BR_NEWTON:
        if (ensure_smem(k_vmult_generic<dim, n, T, BR_NEWTON>, sm))
          return 1;
        k_vmult_generic<dim, n, T, BR_NEWTON><<<GLSB_GRID(G, p), G::THREADS, sm, s>>>(p, S);
        break;
      case BR_FIXED_POINT:
        if (ensure_smem(k_vmult_generic<dim, n, T, BR_FIXED_POINT>, sm))
          return 1;
        k_vmult_generic<dim, n, T, BR_FIXED_POINT><<<GLSB_GRID(G, p), G::THREADS, sm, s>>>(p, S);
        break;
      default:
        if (ensure_smem(k_vmult_generic<dim, n, T, BR_RESIDUAL>, sm))
          return 1;
        k_vmult_generic<dim, n, T, BR_RESIDUAL><<<GLSB_GRID(G, p), G::THREADS, sm, s>>>(p, S);
    }
  return cudaGetLastError() != cudaSuccess;
}

template <int dim, int n, typename T>
static int launch_matrix(int branch, uint32_t n_columns, const KParams<T> &p, const ShapeHost &sh, cudaStream_t s)
{
  using G = Geo<dim, n>;
  if (p.cell_end <= p.cell_begin || n_columns == 0)
    return 0;
  const size_t sm = generic_smem_bytes<dim, n, T>();
  const auto   S  = to_shape<T, n>(sh);
  const dim3   grid(GLSB_GRID(G, p), n_columns);
  if (branch == BR_NEWTON)
    {
      if (ensure_smem(k_vmult_generic<dim, n, T, BR_NEWTON, true>, sm))
        return 1;
      k_vmult_generic<dim, n, T, BR_NEWTON, true><<<grid, G::THREADS, sm, s>>>(p, S);
    }
  else
    {
      if (ensure_smem(k_vmult_generic<dim, n, T, BR_FIXED_POINT, true>, sm))
        return 1;
      k_vmult_generic<dim, n, T, BR_FIXED_POINT, true><<<grid, G::THREADS, sm, s>>>(p, S);
    }
  return cudaGetLastError() != cudaSuccess;
}

template <int dim, int n, typename T>
static int launch_lin(const KParams<T> &p, const ShapeHost &sh, cudaStream_t s)
{
  using G = Geo<dim, n>;
  if (p.cell_end <= p.cell_begin)
    return 0;
  const size_t sm = generic_smem_bytes<dim, n, T>();
  if (ensure_smem(k_linearization<dim, n, T>, sm))
    return 1;
  k_linearization<dim, n, T><<<GLSB_GRID(G, p), G::THREADS, sm, s>>>(p, to_shape<T, n>(sh));
  return cudaGetLastError() != cudaSuccess;
}

template <int dim, int n, typename T>
static int launch_prev(int grad, const KParams<T> &p, const ShapeHost &sh, cudaStream_t s)
{
  using G = Geo<dim, n>;
  if (p.cell_end <= p.cell_begin)
    return 0;
  const size_t sm = generic_smem_bytes<dim, n, T>();
  if (grad)
    {
      if (ensure_smem(k_previous<dim, n, T, true>, sm))
        return 1;
      k_previous<dim, n, T, true><<<GLSB_GRID(G, p), G::THREADS, sm, s>>>(p, to_shape<T, n>(sh));
    }
  else
    {
      if (ensure_smem(k_previous<dim, n, T, false>, sm))
        return 1;
      k_previous<dim, n, T, false><<<GLSB_GRID(G, p), G::THREADS, sm, s>>>(p, to_shape<T, n>(sh));
    }
  return cudaGetLastError() != cudaSuccess;
}

template <int dim, int n, typename T>
static int launch_maxu(const KParams<T> &p, const ShapeHost &sh, cudaStream_t s)
{
  using G = Geo<dim, n>;
  if (p.cell_end <= p.cell_begin)
    return 0;
  const size_t sm = generic_smem_bytes<dim, n, T>();
  if (ensure_smem(k_max_u<dim, n, T>, sm))
    return 1;
  k_max_u<dim, n, T><<<GLSB_GRID(G, p), G::THREADS, sm, s>>>(p, to_shape<T, n>(sh));
  return cudaGetLastError() != cudaSuccess;
}

template <int dim, int n, typename T, int BR>
static int launch_diag_br(const KParams<T> &p, const ShapeHost &sh, const uint8_t *skip, const DiagColumns &dc,
                          cudaStream_t s)
{
  using G         = Geo<dim, n>;
  const size_t sm = generic_smem_bytes<dim, n, T>();
  const auto   S  = to_shape<T, n>(sh);
  if (p.cell_end > p.cell_begin)
    {
      // sum-factorised diagonal; GLSB_DIAG_UNIT_VECTORS=1 selects the unit-vector kernel (cross-check)
      static const bool unit = getenv("GLSB_DIAG_UNIT_VECTORS") != nullptr;
      if (unit)
        {
          if (ensure_smem(k_diag_generic<dim, n, T, BR>, sm))
            return 1;
          k_diag_generic<dim, n, T, BR><<<GLSB_GRID(G, p), G::THREADS, sm, s>>>(p, S, skip);
        }
      else
        {
          if (ensure_smem(k_diag_sumfac<dim, n, T, BR>, sm))
            return 1;
          k_diag_sumfac<dim, n, T, BR><<<GLSB_GRID(G, p), G::THREADS, sm, s>>>(p, S, skip);
        }
    }
  if (dc.n_list > 0)
    {
      if (ensure_smem(k_diag_columns<dim, n, T, BR>, sm))
        return 1;
      k_diag_columns<dim, n, T, BR><<<dc.n_list, G::THREADS, sm, s>>>(p, S, dc);
    }
  return cudaGetLastError() != cudaSuccess;
}

template <int dim, int n, typename T>
static int launch_diag(int branch, const KParams<T> &p, const ShapeHost &sh, const uint8_t *skip,
                       const DiagColumns &dc, cudaStream_t s)
{
  if (branch == BR_NEWTON)
    return launch_diag_br<dim, n, T, BR_NEWTON>(p, sh, skip, dc, s);
  return launch_diag_br<dim, n, T, BR_FIXED_POINT>(p, sh, skip, dc, s);
}

#define GLSB_SWITCH_N(call)          \
  switch (n)                         \
    {                                \
      case 2:                        \
        return call(2);              \
      case 3:                        \
        return call(3);              \
      case 4:                        \
        return call(4);              \
      case 5:                        \
        return call(5);              \
      default:                       \
        return 2;                    \
    }

template <>
int Kernels<GLSB_DIM, GLSB_REAL>::vmult(int n, int branch, const KParams<GLSB_REAL> &p, const ShapeHost &sh,
                                        cudaStream_t s)
{
#if GLSB_DIM == 3
  // degrees 3 and 4: the column kernel (glsb_col.cuh); GLSB_NO_COL=1 keeps the generic kernel (cross-checks)
  static const bool no_col = getenv("GLSB_NO_COL") != nullptr;
  if (!no_col && n == 4)
    return col::launch<4, GLSB_REAL>(branch, p, to_shape<GLSB_REAL, 4>(sh), s);
  if (!no_col && n == 5)
    return col::launch<5, GLSB_REAL>(branch, p, to_shape<GLSB_REAL, 5>(sh), s);
#endif
#define CALL(N) launch_vmult<GLSB_DIM, N, GLSB_REAL>(branch, p, sh, s)
  GLSB_SWITCH_N(CALL)
#undef CALL
}

template <>
int Kernels<GLSB_DIM, GLSB_REAL>::linearization(int n, const KParams<GLSB_REAL> &p, const ShapeHost &sh,
                                                cudaStream_t s)
{
#define CALL(N) launch_lin<GLSB_DIM, N, GLSB_REAL>(p, sh, s)
  GLSB_SWITCH_N(CALL)
#undef CALL
}

template <>
int Kernels<GLSB_DIM, GLSB_REAL>::previous(int n, int with_gradients, const KParams<GLSB_REAL> &p,
                                           const ShapeHost &sh, cudaStream_t s)
{
#define CALL(N) launch_prev<GLSB_DIM, N, GLSB_REAL>(with_gradients, p, sh, s)
  GLSB_SWITCH_N(CALL)
#undef CALL
}

template <>
int Kernels<GLSB_DIM, GLSB_REAL>::diagonal(int n, int branch, const KParams<GLSB_REAL> &p, const ShapeHost &sh,
                                           const uint8_t *skip_cell, const DiagColumns &dc, cudaStream_t s)
{
#define CALL(N) launch_diag<GLSB_DIM, N, GLSB_REAL>(branch, p, sh, skip_cell, dc, s)
  GLSB_SWITCH_N(CALL)
#undef CALL
}

template <>
int Kernels<GLSB_DIM, GLSB_REAL>::vmult_q2(const KParams<GLSB_REAL> &p, const ShapeHost &sh, int F, cudaStream_t s)
{
#if GLSB_DIM == 3 && defined(GLSB_WITH_Q2)
  // one translation unit per (Number, degree) of the register-tiled kernel: glsb_inst_q2.cu
  switch (sh.n)
    {
      case 2:
        return q2::launch_degree<GLSB_REAL, 2>(p, sh, F, s);
      case 3:
        return p.packed ? q2::launch_packed_q2(p, sh, F, s) : q2::launch_degree<GLSB_REAL, 3>(p, sh, F, s);
      case 4:
        return q2::launch_float_only<4>(p, sh, F, s);
      case 5:
        return q2::launch_float_only<5>(p, sh, F, s);
      default:
        return -1;
    }
#else
  (void)p, (void)sh, (void)F, (void)s;
  return -1;
#endif
}

template <>
int Kernels<GLSB_DIM, GLSB_REAL>::matrix_columns(int n, int branch, uint32_t n_columns, const KParams<GLSB_REAL> &p,
                                                 const ShapeHost &sh, cudaStream_t s)
{
#define CALL(N) launch_matrix<GLSB_DIM, N, GLSB_REAL>(branch, n_columns, p, sh, s)
  GLSB_SWITCH_N(CALL)
#undef CALL
}

template <>
int Kernels<GLSB_DIM, GLSB_REAL>::max_u(int n, const KParams<GLSB_REAL> &p, const ShapeHost &sh, cudaStream_t s)
{
#define CALL(N) launch_maxu<GLSB_DIM, N, GLSB_REAL>(p, sh, s)
  GLSB_SWITCH_N(CALL)
#undef CALL
}

} // namespace glsb
