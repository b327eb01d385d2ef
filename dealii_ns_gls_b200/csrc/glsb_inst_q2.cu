// One translation unit per (Number, degree) of the register-tiled vmult kernel, so that the build runs the
// variants in parallel: compile with -DGLSB_REAL=double|float -DGLSB_Q2_N=2..5, or -DGLSB_Q2_PACKED for the
// packed float Q2 kernel (two cells per lane, FFMA2).
#include "glsb_q2.cuh"

namespace glsb
{
namespace q2
{
#ifdef GLSB_Q2_PACKED
int launch_packed_q2(const KParams<float> &p, const ShapeHost &sh, int F, cudaStream_t s)
{
  const auto S2 = to_packed_shape(to_shape<float, 3>(sh));
  if (p.geom == GLSB_GEOM_GENERAL)
    return launch_flags<float, F2, true>(p, S2, F, s);
  return launch_flags<float, F2, false>(p, S2, F, s);
}
#else
template <>
int launch_degree<GLSB_REAL, GLSB_Q2_N>(const KParams<GLSB_REAL> &p, const ShapeHost &sh, int F, cudaStream_t s)
{
  const auto S = to_shape<GLSB_REAL, GLSB_Q2_N>(sh);
  if (p.geom == GLSB_GEOM_GENERAL)
    return launch_flags<GLSB_REAL, GLSB_REAL, true, GLSB_Q2_N>(p, S, F, s);
  return launch_flags<GLSB_REAL, GLSB_REAL, false, GLSB_Q2_N>(p, S, F, s);
}
#endif
} // namespace q2
} // namespace glsb
