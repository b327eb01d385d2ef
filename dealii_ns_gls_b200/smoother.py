"""Host-side mirror of the multigrid smoother the reference builds around the level operators:
``PreconditionRelaxation<OperatorBase<MGNumber>, DiagonalMatrix<VectorType<MGNumber>>>``
(include/multigrid.h:67-69), configured in PreconditionerGMG::initialize
(include/multigrid.cc:282-305, :347-370): relaxation = 0 (estimate it), smoothing_range = 20,
n_iterations = 5, eig_cg_n_iterations = 20, EigenvalueAlgorithm::power_iteration.

The class has deal.II's method names (vmult / step / estimate_eigenvalues / get_relaxation); every
method forwards to libglsb200.so, where all sweeps of a call run on the device back to back
(SURVEY.md section 8f, rank 1).  With a ghost exchange attached the sweeps are driven from here:
exchange-aware vmult + the library's fused update kernel."""
from __future__ import annotations

import ctypes as C

import torch



class EigenvalueInformation:
    def __init__(self, min_ev, max_ev, n_iterations):
        self.min_eigenvalue_estimate = min_ev
        self.max_eigenvalue_estimate = max_ev
        self.cg_iterations = n_iterations


class PreconditionRelaxation:
    def __init__(self, op, inverse_diagonal: torch.Tensor | None = None, *, relaxation=0.0, n_iterations=5,
                 smoothing_range=20.0, eig_cg_n_iterations=20, first_local_index=0):
        self.op = op
        if inverse_diagonal is None:  # multigrid.cc:292-293
            inverse_diagonal = op.initialize_dof_vector()
            op.compute_inverse_diagonal(inverse_diagonal)
        self.inverse_diagonal = inverse_diagonal
        self.relaxation = float(relaxation)
        self.n_iterations = int(n_iterations)
        self.smoothing_range = float(smoothing_range)
        self.eig_cg_n_iterations = int(eig_cg_n_iterations)
        self.first_local_index = int(first_local_index)
        self._eigenvalues = None

    # ---- deal.II interface ----
    def get_relaxation(self):
        if self.relaxation == 0.0:
            self.estimate_eigenvalues()
        return self.relaxation

    def estimate_eigenvalues(self, _vec=None):
        op = self.op
        if op.exchange is not None:
            return self._estimate_eigenvalues_partitioned()
        omega, ev_max = C.c_double(0), C.c_double(0)
        w = op.time_integrator_data.get_primary_weight()
        op._chk(op._lib.glsb_estimate_relaxation(op._op, op._vec(self.inverse_diagonal, "inverse_diagonal"),
                                                 self.eig_cg_n_iterations, self.smoothing_range, w,
                                                 self.first_local_index, C.byref(omega), C.byref(ev_max),
                                                 op._stream()), "estimate_eigenvalues")
        self._eigenvalues = EigenvalueInformation(ev_max.value / self.smoothing_range, ev_max.value,
                                                  self.eig_cg_n_iterations)
        if self.relaxation == 0.0:
            self.relaxation = omega.value
        return self._eigenvalues

    def _estimate_eigenvalues_partitioned(self):
        """The power iteration of glsb_estimate_relaxation driven from the host layer for operators with a ghost
        exchange: exchange-aware vmult, reductions on the owned block + MPI::sum.  Same start vector (global
        index % 11, mean-free, zero on constrained dofs), same 1.2 safety factor."""
        op, ex = self.op, self.op.exchange
        no, dev = op.n_owned, op.device
        first = op.mesh.partition.owned_offset if op.mesh.partition is not None else self.first_local_index
        n_global = op.mesh.n_global_dofs
        e = op.initialize_dof_vector()
        e[:no] = ((torch.arange(no, device=dev, dtype=torch.int64) + first) % 11).to(op.dtype)
        (total,) = ex.allreduce_sum([float(e[:no].double().sum())])
        e[:no] -= total / n_global
        op.get_constraints().set_zero(e)
        e[no:] = 0
        (nrm2,) = ex.allreduce_sum([float(torch.dot(e[:no].double(), e[:no].double()))])
        e[:no] /= nrm2 ** 0.5
        v = op.initialize_dof_vector()
        lam = 0.0
        for _ in range(self.eig_cg_n_iterations):
            op.vmult(v, e)
            v[:no] *= self.inverse_diagonal[:no]
            lam, nrm2 = ex.allreduce_sum([float(torch.dot(e[:no].double(), v[:no].double())),
                                          float(torch.dot(v[:no].double(), v[:no].double()))])
            e[:no] = v[:no] / nrm2 ** 0.5
        ev_max = 1.2 * abs(lam)  # deal.II's power_iteration returns std::abs(eigenvalue_estimate)
        alpha = ev_max / self.smoothing_range if self.smoothing_range > 1.0 else 0.9 * ev_max
        self._eigenvalues = EigenvalueInformation(ev_max / self.smoothing_range, ev_max, self.eig_cg_n_iterations)
        if self.relaxation == 0.0:
            self.relaxation = 2.0 / (alpha + ev_max)
        return self._eigenvalues

    def vmult(self, dst: torch.Tensor, src: torch.Tensor):
        """n_iterations sweeps from a zero initial guess."""
        self._sweeps(dst, src, True)

    def step(self, dst: torch.Tensor, src: torch.Tensor):
        """n_iterations sweeps from the current dst."""
        self._sweeps(dst, src, False)

    # ---- implementation ----
    def _sweeps(self, dst, src, from_zero):
        op, omega = self.op, self.get_relaxation()
        w = op.time_integrator_data.get_primary_weight()
        d = op._vec(self.inverse_diagonal, "inverse_diagonal")
        if op.exchange is None:
            fn = op._lib.glsb_relaxation_vmult if from_zero else op._lib.glsb_relaxation_step
            op._chk(fn(op._op, op._vec(dst, "dst"), op._vec(src, "src"), d, omega, self.n_iterations, w, op._stream()),
                    "relaxation")
            return
        it = 0
        if from_zero:
            dst.zero_()
            if self.n_iterations == 0:
                return
            dst[:op.n_owned] = omega * (self.inverse_diagonal[:op.n_owned] * src[:op.n_owned])
            it = 1
        tmp = torch.empty_like(dst)
        for _ in range(it, self.n_iterations):
            op.vmult(tmp, dst)
            op._chk(op._lib.glsb_relaxation_update(op._op, op._vec(dst, "dst"), op._vec(tmp, "tmp"),
                                                   op._vec(src, "src"), d, omega, op._stream()), "relaxation")
