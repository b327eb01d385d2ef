// Compile, instantiate and link check of the deal.II adapter (dealii_ns_gls_b200/cpp/dealii_adapter.h)
// against tests/cpp/dealii_stub/ -- the used deal.II and reference signatures only (this image has no
// deal.II).  Every member of NavierStokesOperatorB200<dim, Number> is instantiated explicitly, so each
// `override` is checked against OperatorBase<Number> (include/operator_base.h:13-73 of the reference) and each
// C-ABI / NCCL / CUDA call against its header; the link step resolves them in libglsb200.so, libnccl, libcudart.
// Run without arguments it only constructs nothing and prints the instantiated sizes.
#include "../../dealii_ns_gls_b200/cpp/dealii_adapter.h"

#include <cstdio>

template class glsb::NavierStokesOperatorB200<2, double>;
template class glsb::NavierStokesOperatorB200<3, double>;
template class glsb::NavierStokesOperatorB200<2, float>;
template class glsb::NavierStokesOperatorB200<3, float>;

int main()
{
  static_assert(std::is_base_of<OperatorBase<double>, glsb::NavierStokesOperatorB200<3, double>>::value, "drop-in");
  static_assert(std::is_base_of<OperatorBase<float>, glsb::NavierStokesOperatorB200<3, float>>::value, "drop-in");
  static_assert(!std::is_abstract<glsb::NavierStokesOperatorB200<3, double>>::value, "all pure virtuals defined");
  static_assert(!std::is_abstract<glsb::NavierStokesOperatorB200<2, float>>::value, "all pure virtuals defined");
  std::printf("adapter instantiated: sizeof <3,double> = %zu, <3,float> = %zu\n",
              sizeof(glsb::NavierStokesOperatorB200<3, double>), sizeof(glsb::NavierStokesOperatorB200<3, float>));
  return 0;
}
