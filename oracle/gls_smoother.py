"""CPU restatement of the multigrid smoother of the reference -- TEST INFRASTRUCTURE ONLY (see the
header of gls_oracle.py; same rules, same "parity unpinned" status).

The reference builds ``PreconditionRelaxation<OperatorBase<MGNumber>, DiagonalMatrix>``
(include/multigrid.h:67-69) with relaxation = 0, smoothing_range = 20, n_iterations = 5,
eig_cg_n_iterations = 20 and EigenvalueAlgorithm::power_iteration (include/multigrid.cc:290-304);
the arithmetic lives in deal.II (lac/precondition.h, not vendored).  Restated from deal.II's
published implementation:
  vmult : x_1 = omega D^-1 b, then x <- x + omega D^-1 (b - A x)   (n_iterations sweeps in total)
  step  : n_iterations sweeps from the given x
  estimate_eigenvalues (power_iteration): e_i = (global index i) % 11, minus the mean, zero on
          constrained dofs, normalised; n times: v = D^-1 A e, lambda = e . v, e = v / |v|;
          max = 1.2 lambda, min = max / smoothing_range, omega = 2 / (min + max).
"""
import numpy as np


class OracleRelaxation:
    def __init__(self, oracle, weight, inverse_diagonal, *, relaxation=0.0, n_iterations=5, smoothing_range=20.0,
                 eig_cg_n_iterations=20, first_local_index=0):
        self.A = lambda x: oracle.vmult(x, weight)
        self.dtype = oracle.dtype
        self.constrained = np.asarray(oracle.constrained, dtype=np.int64)
        self.d = np.asarray(inverse_diagonal, dtype=self.dtype)
        self.relaxation, self.n_iterations = float(relaxation), int(n_iterations)
        self.smoothing_range, self.eig_n, self.first = float(smoothing_range), int(eig_cg_n_iterations), first_local_index
        self.max_eigenvalue_estimate = None

    def estimate_eigenvalues(self):
        n = len(self.d)
        e = ((np.arange(n) + self.first) % 11).astype(self.dtype)
        e -= e.mean(dtype=np.float64).astype(self.dtype)
        if len(self.constrained):
            e[self.constrained] = 0
        e /= np.linalg.norm(e.astype(np.float64))
        e = e.astype(self.dtype)
        lam = 0.0
        for _ in range(self.eig_n):
            v = (self.d * self.A(e)).astype(self.dtype)
            lam = float(np.dot(e.astype(np.float64), v.astype(np.float64)))
            e = (v / np.linalg.norm(v.astype(np.float64))).astype(self.dtype)
        self.max_eigenvalue_estimate = 1.2 * abs(lam)
        alpha = self.max_eigenvalue_estimate / self.smoothing_range if self.smoothing_range > 1 \
            else 0.9 * self.max_eigenvalue_estimate
        if self.relaxation == 0.0:
            self.relaxation = 2.0 / (alpha + self.max_eigenvalue_estimate)
        return self.max_eigenvalue_estimate

    def get_relaxation(self):
        if self.relaxation == 0.0:
            self.estimate_eigenvalues()
        return self.relaxation

    def _sweep(self, x, b):
        T = self.dtype.type
        return (x + T(self.get_relaxation()) * (self.d * (b - self.A(x)))).astype(self.dtype)

    def vmult(self, b):
        b = np.asarray(b, dtype=self.dtype)
        if self.n_iterations == 0:
            return np.zeros_like(b)
        x = (self.dtype.type(self.get_relaxation()) * (self.d * b)).astype(self.dtype)
        for _ in range(1, self.n_iterations):
            x = self._sweep(x, b)
        return x

    def step(self, x, b):
        x, b = np.asarray(x, dtype=self.dtype), np.asarray(b, dtype=self.dtype)
        for _ in range(self.n_iterations):
            x = self._sweep(x, b)
        return x
