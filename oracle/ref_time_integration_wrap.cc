// C entry points around the reference's OWN time-integration classes (include/time_integration.h:11-164), compiled
// from /root/reference/include/time_integration.cc as it lies there (oracle/Makefile, target _ref ->
// oracle/_ref/libref_time_integration.so).  TEST INFRASTRUCTURE: tests/test_reference_time_integration.py pins
// the restatements (oracle/gls_oracle.py OracleBDF, dealii_ns_gls_b200/time_integration.py, the C++ mirror in
// cpp/operator_b200.h) against the output of this object code.  It is the one file on the path whose only
// deal.II dependencies are a vector type and AssertThrow (stand-ins in oracle/ref_shim/).
#include "time_integration.h"

#include <cstring>

extern "C"
{
  // kind: 0 = BDF(order), 1 = Theta(theta), 2 = None
  void *
  reft_create(int kind, int order, double theta)
  {
    try
      {
        if (kind == 0)
          return new TimeIntegratorDataBDF(order);
        if (kind == 1)
          return new TimeIntegratorDataTheta(theta);
        return new TimeIntegratorDataNone();
      }
    catch (...)
      {
        return nullptr;
      }
  }

  void
  reft_destroy(void *p)
  {
    delete static_cast<TimeIntegratorData *>(p);
  }

  // returns 0, or 1 if the reference threw (e.g. "Not implemented")
  int
  reft_update_dt(void *p, double dt)
  {
    try
      {
        static_cast<TimeIntegratorData *>(p)->update_dt(dt);
        return 0;
      }
    catch (...)
      {
        return 1;
      }
  }

  // fills out[0 .. n_weights) (at most cap entries) and the four scalars; returns the number of weights
  int
  reft_query(void *p, double *weights, int cap, double *primary_weight, double *current_dt, double *theta,
             unsigned int *order)
  {
    const TimeIntegratorData *t = static_cast<TimeIntegratorData *>(p);
    const std::vector<Number> &w = t->get_weights();
    for (int i = 0; i < (int)w.size() && i < cap; ++i)
      weights[i] = w[i];
    *primary_weight = t->get_primary_weight();
    *current_dt     = t->get_current_dt();
    *theta          = t->get_theta();
    *order          = t->get_order();
    return (int)w.size();
  }

  // SolutionHistory<double>(size) with vector i holding the single value in[i]; `commits` calls of
  // commit_solution(); out[i] = value of vector i afterwards
  void
  reft_history_commit(int size, const double *in, int commits, double *out)
  {
    SolutionHistory<double> h(size);
    for (int i = 0; i < size; ++i)
      h.get_vectors()[i].values.assign(1, in[i]);
    for (int c = 0; c < commits; ++c)
      h.commit_solution();
    for (int i = 0; i < size; ++i)
      out[i] = h.get_vectors()[i].values.empty() ? 0.0 : h.get_vectors()[i].values[0];
  }
}
