// C entry points around the C++ host mirror's time-integration classes (dealii_ns_gls_b200/cpp/operator_b200.h),
// with the interface of oracle/ref_time_integration_wrap.cc, so that tests/test_reference_time_integration.py can
// run the record of the reference's own include/time_integration.cc through them.  Host only (g++).
#include "../../dealii_ns_gls_b200/cpp/operator_b200.h"

extern "C"
{
  void *
  mirror_create(int kind, int order, double theta)
  {
    try
      {
        if (kind == 0)
          return static_cast<glsb::TimeIntegratorData *>(new glsb::TimeIntegratorDataBDF(order));
        if (kind == 1)
          return static_cast<glsb::TimeIntegratorData *>(new glsb::TimeIntegratorDataTheta(theta));
        return static_cast<glsb::TimeIntegratorData *>(new glsb::TimeIntegratorDataNone());
      }
    catch (...)
      {
        return nullptr;
      }
  }

  void
  mirror_destroy(void *p)
  {
    delete static_cast<glsb::TimeIntegratorData *>(p);
  }

  int
  mirror_update_dt(void *p, double dt)
  {
    try
      {
        static_cast<glsb::TimeIntegratorData *>(p)->update_dt(dt);
        return 0;
      }
    catch (...)
      {
        return 1;
      }
  }

  int
  mirror_query(void *p, double *weights, int cap, double *primary_weight, double *current_dt, double *theta,
               unsigned int *order)
  {
    const glsb::TimeIntegratorData *t = static_cast<glsb::TimeIntegratorData *>(p);
    const std::vector<double>      &w = t->get_weights();
    for (int i = 0; i < (int)w.size() && i < cap; ++i)
      weights[i] = w[i];
    *primary_weight = t->get_primary_weight();
    *current_dt     = t->get_current_dt();
    *theta          = t->get_theta();
    *order          = t->get_order();
    return (int)w.size();
  }
}
