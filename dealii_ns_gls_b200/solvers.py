"""Host-side mirror of the reference's linear and nonlinear solvers with DEVICE-resident vectors
(SURVEY.md section 8f, rank 3).

    LinearSolverGMRES      include/solver_l.cc:26-74   -> deal.II SolverGMRES, max_n_tmp_vectors = 30 (Arnoldi
                                                          basis of 28 vectors), right preconditioning,
                                                          tolerance max(rel * |b|, abs)
    NonLinearSolverNewton  include/solver_nl.cc:28-89  -> tolerance 1e-7, at most 30 iterations, "inexact"
                                                          reuses the preconditioner of the first iteration

The control flow is the reference's; the vector work is done by libglsb200.so on the device: the Krylov basis is
one contiguous block, the inner products of a Gram-Schmidt step are ONE pass over the data
(glsb_vec_multi_dot), the update another (glsb_vec_multi_axpy), and the only host round trip per iteration is
the read of the k + 1 Hessenberg entries.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .multigrid import DeviceVectorOps


class SolverControl:
    """deal.II SolverControl(n_max_iterations, tolerance): last_step / last_value."""

    def __init__(self, n_max_iterations, tolerance):
        self.max_steps, self.tol = int(n_max_iterations), float(tolerance)
        self._last_step, self._last_value, self.history = 0, float("nan"), []

    def check(self, step, value):
        self._last_step, self._last_value = step, value
        self.history.append(value)
        if value <= self.tol:
            return "success"
        if step >= self.max_steps or not math.isfinite(value):
            return "failure"
        return "iterate"

    def last_step(self):
        return self._last_step

    def last_value(self):
        return self._last_value


class NoConvergence(RuntimeError):
    pass


class LinearSolverGMRES:
    def __init__(self, op, preconditioner, n_max_iterations=10000, absolute_tolerance=1e-12,
                 relative_tolerance=1e-2, max_n_tmp_vectors=30):
        self.op, self.preconditioner = op, preconditioner
        self.n_max_iterations = n_max_iterations
        self.absolute_tolerance, self.relative_tolerance = absolute_tolerance, relative_tolerance
        self.basis_size = max_n_tmp_vectors - 2  # deal.II: Arnoldi basis = max_n_tmp_vectors - 2
        self._ops = DeviceVectorOps()
        self._V = None
        self.solver_control = None
        self.n_iterations = []

    def initialize(self):
        pass  # solver_l.cc:39-43

    def _norm(self, v, scratch):
        self._ops.multi_dot(scratch, v.view(1, -1), 1, v)
        return math.sqrt(float(scratch[0].item()))

    def solve(self, dst: torch.Tensor, src: torch.Tensor):
        """solver_l.cc:45-74"""
        op, ops, m = self.op, self._ops, self.basis_size
        n = src.numel()
        if self._V is None or self._V.shape[1] != n or self._V.dtype != src.dtype:
            self._V = torch.empty((m + 1, n), dtype=src.dtype, device=src.device)
            self._p = torch.empty_like(src)
            self._z = torch.empty_like(src)
            self._hd = torch.empty(m + 2, dtype=torch.float64, device=src.device)
            self._yd = torch.empty(m + 2, dtype=torch.float64, device=src.device)
        V, p, z, hd, yd = self._V, self._p, self._z, self._hd, self._yd
        tol = max(self.relative_tolerance * self._norm(src, hd), self.absolute_tolerance)
        control = SolverControl(self.n_max_iterations, tol)
        self.solver_control = control
        dst.zero_()
        accumulated = 0
        state = "iterate"
        while state == "iterate":
            # v_0 = b - A x
            op.vmult(p, dst)
            ops.axpby(V[0], 1.0, src, 0.0)
            ops.axpby(V[0], -1.0, p, 1.0)
            rho = self._norm(V[0], hd)
            state = control.check(accumulated, rho)
            if state != "iterate":
                break
            ops.axpby(V[0], 1.0 / rho, V[0], 0.0)
            H = np.zeros((m + 1, m))
            gamma = np.zeros(m + 1)
            gamma[0] = rho
            ci, si = np.zeros(m), np.zeros(m)
            dim = 0
            for inner in range(m):
                accumulated += 1
                # right preconditioning: w = A M^-1 v_inner
                self.preconditioner.vmult(z, V[inner])
                w = V[inner + 1]
                op.vmult(w, z)
                # classical Gram-Schmidt with one re-orthogonalisation, both passes batched on the device
                ops.multi_dot(hd, V, inner + 1, w)
                ops.multi_axpy(w, V, inner + 1, hd, -1.0)
                ops.multi_dot(yd, V, inner + 1, w)
                ops.multi_axpy(w, V, inner + 1, yd, -1.0)
                ops.multi_dot(hd[m + 1:], w.view(1, -1), 1, w)
                hh = hd.cpu().numpy()
                h = hh[:inner + 1] + yd[:inner + 1].cpu().numpy()
                s = math.sqrt(max(hh[m + 1], 0.0))
                if s > 0:
                    ops.axpby(w, 1.0 / s, w, 0.0)
                # Givens rotations (deal.II SolverGMRES::givens_rotation)
                for i in range(inner):
                    t = h[i]
                    h[i] = ci[i] * t + si[i] * h[i + 1]
                    h[i + 1] = -si[i] * t + ci[i] * h[i + 1]
                r = math.hypot(h[inner], s)
                ci[inner], si[inner] = h[inner] / r, s / r
                h[inner] = r
                gamma[inner + 1] = -si[inner] * gamma[inner]
                gamma[inner] *= ci[inner]
                H[:inner + 1, inner] = h
                dim = inner + 1
                rho = abs(gamma[dim])
                state = control.check(accumulated, rho)
                if state != "iterate":
                    break
            # x += M^-1 (V y), H y = gamma
            y = np.linalg.solve(np.triu(H[:dim, :dim]), gamma[:dim])
            yd[:dim] = torch.from_numpy(y).to(yd.device)
            p.zero_()
            ops.multi_axpy(p, V, dim, yd, 1.0)
            self.preconditioner.vmult(z, p)
            ops.axpby(dst, 1.0, z, 1.0)
        self.n_iterations.append(control.last_step())
        if state == "failure":
            raise NoConvergence(f"GMRES: {control.last_step()} iterations, residual {control.last_value()}")
        self.preconditioner.print_stats()


class NonLinearSolverBase:
    """include/solver_nl.h:14-35: the six hooks the driver sets (main.cc:805-869)."""

    def __init__(self):
        self.setup_jacobian = None
        self.setup_preconditioner = None
        self.evaluate_rhs = None
        self.evaluate_residual = None
        self.solve_with_jacobian = None
        self.postprocess = None


class NonLinearSolverNewton(NonLinearSolverBase):
    def __init__(self, inexact_newton=False):
        super().__init__()
        self.newton_tolerance, self.newton_max_iteration = 1.0e-7, 30
        self.inexact_newton = inexact_newton
        self.residuals = []
        self._ops = DeviceVectorOps()

    def _l2(self, v, scratch):
        self._ops.multi_dot(scratch, v.view(1, -1), 1, v)
        return math.sqrt(float(scratch[0].item()))

    def solve(self, solution: torch.Tensor):
        """solver_nl.cc:36-89"""
        rhs, inc = torch.zeros_like(solution), torch.zeros_like(solution)
        scratch = torch.zeros(1, dtype=torch.float64, device=solution.device)
        self.setup_jacobian(solution)
        self.evaluate_residual(rhs, solution)
        l2_norm, num_iteration = self._l2(rhs, scratch), 0
        self.residuals = [l2_norm]
        while l2_norm > self.newton_tolerance:
            inc.zero_()
            if num_iteration == 0 or not self.inexact_newton:
                self.setup_preconditioner(solution)
            self.solve_with_jacobian(inc, rhs)
            self._ops.axpby(solution, 1.0, inc, 1.0)
            if self.postprocess:
                self.postprocess(solution)
            self.setup_jacobian(solution)
            self.evaluate_residual(rhs, solution)
            l2_norm = self._l2(rhs, scratch)
            num_iteration += 1
            self.residuals.append(l2_norm)
            if num_iteration > self.newton_max_iteration:
                raise NoConvergence(f"Newton iteration did not converge. Final residual is {l2_norm}.")
        return num_iteration


class NonLinearSolverLinearized(NonLinearSolverBase):
    def solve(self, solution: torch.Tensor):
        """solver_nl.cc:10-24"""
        self.setup_jacobian(solution)
        rhs = torch.zeros_like(solution)
        self.evaluate_rhs(rhs)
        self.setup_preconditioner(solution)
        self.solve_with_jacobian(solution, rhs)
        return 1
