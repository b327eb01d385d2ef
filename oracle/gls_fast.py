"""gls_oracle.OracleOperator with its cell loops executed by the C restatement (oracle/gls_oracle_c.c) -- TEST
INFRASTRUCTURE ONLY; exists so that the CPU restatement of the SOLVER stack (oracle/gls_solver.py: Newton, GMRES,
V-cycle, relaxation smoothers, time loop) can be timed on the host cores next to the device's wall time per time
step (`time_step.cpu_baseline` of bench.py).  Tables, penalty parameters and max |u| stay with the numpy code of
the base class (one evaluation pass each); vmult, the residual and the unit-vector diagonal
(MatrixFreeTools::compute_diagonal, operator_ns.cc:195-225) run in C, vectorised over 8 cells and OpenMP-threaded,
in double precision whatever the level number type (results are rounded to it).  Zero-type constraint rows only
(what the channel and cylinder configurations have)."""
from __future__ import annotations

import time

import numpy as np

from . import gls_oracle as go
from .gls_oracle_c import COracle, max_threads


class FastOracleOperator(go.OracleOperator):
    c_seconds = 0.0   # wall time spent inside the C cell loops, summed over all instances (for the breakdown of a
    #                   timed run: the rest is numpy / scipy glue a compiled CPU code would not have)

    def _timed(self, fn, *a):
        t0 = time.perf_counter()
        out = fn(*a)
        FastOracleOperator.c_seconds += time.perf_counter() - t0
        return out

    def __init__(self, **kw):
        kw.setdefault("path", "sumfac")
        super().__init__(**kw)
        assert all(len(r) == 0 for r in self.constraints.values()), "zero-type constraint rows only"
        self._c = {}            # branch -> COracle (mesh and geometry interleaved once)
        self._stale = set()     # branches whose tables are out of date
        # threads in proportion to the work: the coarse levels of a hierarchy have a few dozen cells, where a
        # fork/join over all cores costs more than the loop
        self.n_threads = max(1, min(max_threads(), self.n_cells // 256))
        # Cartesian cells (constant diagonal J^-1): the C loop then reads det J and diag J^-1 per cell like
        # MatrixFree's Cartesian path, not (dim^2 + 1) numbers per quadrature point
        J = np.asarray(self.Jinv, dtype=np.float64)
        d = self.dim
        off = J.copy()
        for e in range(d):
            off[:, :, e, e] = 0
        tol = 1e-12 * np.abs(J).max()
        self._cartesian = bool(np.abs(off).max() <= tol and np.abs(J - J[:, :1]).max() <= tol)

    # ---- C handles ----
    def _handle(self, branch):
        co = self._c.get(branch)
        if co is None:
            d = self.dim
            b = self.tb.b
            if self._cartesian:
                J = np.asarray(self.Jinv, dtype=np.float64)
                ij = np.stack([J[:, 0, e, e] for e in range(d)], axis=1)
                det = np.asarray(self.JxW, dtype=np.float64)[:, 0] / float(self.tb.w[0])
                co = COracle(dim=d, degree=self.degree, cell_dofs=self.cell_dofs, n_dofs=self.n_dofs, S=b.S, D=b.D,
                             w=b.wq, cartesian=True, inv_jac=ij, jxw=det, nu=self.nu, theta=self.theta, branch=branch,
                             ctd=self.ctd, cell_wise=self.cell_wise)
            else:
                co = COracle(dim=d, degree=self.degree, cell_dofs=self.cell_dofs, n_dofs=self.n_dofs, S=b.S, D=b.D,
                             w=b.wq, cartesian=False, inv_jac=self.Jinv, jxw=self.JxW, nu=self.nu, theta=self.theta,
                             branch=branch, ctd=self.ctd, cell_wise=self.cell_wise)
            self._c[branch] = co
            self._stale.add(branch)
        if branch in self._stale:
            K, d = self.n_cells, self.dim
            d1 = self.delta1_cell.reshape(K, 1) if self.cell_wise else self.delta1_q
            d2 = self.delta2_cell.reshape(K, 1) if self.cell_wise else self.delta2_q
            co.set_tables(self.U, self.H.reshape(K, d * d, -1), self.P, self.o,
                          None if self.Gold is None else self.Gold.reshape(K, d * d, -1), self.gold_p, d1, d2)
            self._stale.discard(branch)
        return co

    def set_linearization_point(self, vec, dt):
        super().set_linearization_point(vec, dt)
        self._stale = set(self._c)

    def set_previous_solution(self, history, weights):
        super().set_previous_solution(history, weights)
        self._stale = set(self._c)

    # ---- cell loops in C ----
    def vmult(self, src, weight, edge_constrained_indices=None):
        assert edge_constrained_indices is None or len(edge_constrained_indices) == 0
        src = np.asarray(src)
        x = np.array(src, dtype=np.float64)
        if len(self.constrained):
            x[self.constrained] = 0.0
        co = self._handle(COracle.BR_NEWTON if self.increment_form else COracle.BR_FIXED_POINT)
        dst = self._timed(co.apply, x, weight, self.n_threads)
        if len(self.constrained):
            dst[self.constrained] = src[self.constrained]
        return dst.astype(self.dtype)

    def evaluate_residual(self, src_with_bc, weight):
        dst = self._timed(self._handle(COracle.BR_RESIDUAL).apply, np.asarray(src_with_bc, dtype=np.float64), weight,
                          self.n_threads)
        if len(self.constrained):
            dst[self.constrained] = 0.0
        return (-dst).astype(self.dtype)

    def compute_inverse_diagonal(self, weight, edge_constrained_indices=None):
        co = self._handle(COracle.BR_NEWTON if self.increment_form else COracle.BR_FIXED_POINT)
        diag = self._timed(co.diagonal, weight, self.n_threads)
        if len(self.constrained):
            diag[self.constrained] = 1.0
        if edge_constrained_indices is not None and len(edge_constrained_indices):
            diag[np.asarray(edge_constrained_indices, dtype=np.int64)] = 0.0
        diag = diag.astype(self.dtype)
        T = self.dtype.type
        with np.errstate(divide="ignore"):
            return np.where(np.abs(diag) > T(1e-10), T(1.0) / diag, T(1.0)).astype(self.dtype)
