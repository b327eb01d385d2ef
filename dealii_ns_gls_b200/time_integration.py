"""Host-side mirror of the reference's TimeIntegratorData hierarchy and SolutionHistory
(include/time_integration.h:10-164, include/time_integration.cc).  The operator reads
get_primary_weight()/get_current_dt()/get_weights() at call time, exactly like the
reference does through its `const TimeIntegratorData &` member (operator_ns.cc:348, :958)."""
from __future__ import annotations


class TimeIntegratorData:
    def update_dt(self, dt_new): raise NotImplementedError
    def get_primary_weight(self): raise NotImplementedError
    def get_weights(self): raise NotImplementedError
    def get_order(self): raise NotImplementedError
    def get_current_dt(self): raise NotImplementedError
    def get_theta(self): raise NotImplementedError


class TimeIntegratorDataBDF(TimeIntegratorData):
    """Variable-step BDF1-3 (time_integration.cc:4-91)."""

    def __init__(self, order):
        self.order = int(order)
        self.dt = [0.0] * self.order
        self.weights = [0.0] * (self.order + 1)

    def update_dt(self, dt_new):
        for i in range(self.order - 2, -1, -1):
            self.dt[i + 1] = self.dt[i]
        self.dt[0] = float(dt_new)
        self._update_weights()

    def _effective_order(self):
        return sum(1 for v in self.dt if v > 0)

    def _update_weights(self):
        dt, w = self.dt, [0.0] * (self.order + 1)
        eff = self._effective_order()
        if eff == 3:
            w[1] = -(dt[0] + dt[1]) * (dt[0] + dt[1] + dt[2]) / (dt[0] * dt[1] * (dt[1] + dt[2]))
            w[2] = dt[0] * (dt[0] + dt[1] + dt[2]) / (dt[1] * dt[2] * (dt[0] + dt[1]))
            w[3] = -dt[0] * (dt[0] + dt[1]) / (dt[2] * (dt[1] + dt[2]) * (dt[0] + dt[1] + dt[2]))
            w[0] = -(w[1] + w[2] + w[3])
        elif eff == 2:
            w[0] = (2 * dt[0] + dt[1]) / (dt[0] * (dt[0] + dt[1]))
            w[1] = -(dt[0] + dt[1]) / (dt[0] * dt[1])
            w[2] = dt[0] / (dt[1] * (dt[0] + dt[1]))
        elif eff == 1:
            w[0] = 1.0 / dt[0]
            w[1] = -1.0 / dt[0]
        else:
            raise RuntimeError("Not implemented")  # AssertThrow(effective_order() <= 3)
        self.weights = w

    def get_primary_weight(self): return self.weights[0]
    def get_weights(self): return self.weights
    def get_order(self): return self.order
    def get_current_dt(self): return self.dt[0]
    def get_theta(self): return 1.0


class TimeIntegratorDataTheta(TimeIntegratorData):
    """One-step theta method (time_integration.cc:95-137)."""

    def __init__(self, theta):
        self.theta = float(theta)
        self.dt = 0.0
        self.weights = [0.0, 0.0]

    def update_dt(self, dt_new):
        self.dt = float(dt_new)
        self.weights = [1.0 / self.dt, -1.0 / self.dt]

    def get_primary_weight(self): return self.weights[0]
    def get_weights(self): return self.weights
    def get_order(self): return 1
    def get_current_dt(self): return self.dt
    def get_theta(self): return self.theta


class TimeIntegratorDataNone(TimeIntegratorData):
    """Stationary (time_integration.cc:141-178): order 0, weight 0, dt 1, theta 1."""

    def update_dt(self, dt_new): pass
    def get_primary_weight(self): return 0.0
    def get_weights(self): return []
    def get_order(self): return 0
    def get_current_dt(self): return 1.0
    def get_theta(self): return 1.0


class SolutionHistory:
    """time_integration.h:145-164: solutions[0] is the current solution."""

    def __init__(self, size):
        self.solutions = [None] * size

    def get_current_solution(self): return self.solutions[0]
    def get_vectors(self): return self.solutions

    def commit_solution(self):
        for i in range(len(self.solutions) - 2, -1, -1):
            self.solutions[i + 1].copy_(self.solutions[i])
