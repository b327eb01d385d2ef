"""CPU restatement of the solver stack the reference wraps around its operator -- TEST INFRASTRUCTURE ONLY
(same rules and the same "parity unpinned" status as gls_oracle.py: the arithmetic of these pieces lives in
deal.II, which is not vendored; restated from deal.II's published implementation and anchored on the
reference's call sites).

  two-level transfer   main.cc:540-563 -> deal.II MGTwoLevelTransfer: prolongation by the FE_Q embedding of
                       every coarse cell into its 2^dim children, weighted by 1 / (number of fine cells touching
                       a dof) and 0 on constrained fine dofs; coarse constraints resolved on read / transposed
                       on write; restriction = transpose; interpolate = fine function at the coarse nodes
  V-cycle              include/multigrid.cc:534-548 -> deal.II Multigrid::level_v_step with
                       MGSmootherPrecondition (pre: smoother.vmult from zero; post: u += P (rhs - A u))
  coarse solver        "direct" (multigrid.cc:419-425): dense solve with the level-0 matrix
  GMRES                include/solver_l.cc:45-74 -> deal.II SolverGMRES, 28 basis vectors, right preconditioning
  Newton               include/solver_nl.cc:36-89
  time loop            main.cc:908-990 with the hooks of main.cc:772-869

Everything is explicit (scipy sparse matrices, modified Gram-Schmidt) and written independently of the device
code in dealii_ns_gls_b200/: the tests compare iteration counts and solution vectors of the two.
"""
import math

import numpy as np
import scipy.sparse as sp

from . import gls_oracle as go
from .gls_smoother import OracleRelaxation


# ---------------------------------------------------------------------------------------------------------
# transfer
# ---------------------------------------------------------------------------------------------------------
def _embedding_1d(degree):
    """P[a][l, j] = phi_j^coarse((x_l + a) / 2) for the two children a = 0, 1"""
    nodes = go.gauss_lobatto_points(degree)
    return [go.lagrange_tables(nodes, (nodes + a) / 2.0)[0] for a in (0, 1)]


def _kron_lex(mats):
    """tensor product with x fastest: mats = [M_x, M_y(, M_z)]"""
    out = mats[0]
    for m in mats[1:]:
        out = np.kron(m, out)
    return out


def constraint_matrix(n, constraints):
    """C with x_resolved = C x: identity on free dofs, the constraint row on constrained ones"""
    rows, cols, vals = [], [], []
    for i in range(n):
        if i in constraints:
            for m, w in constraints[i]:
                rows.append(i), cols.append(m), vals.append(w)
        else:
            rows.append(i), cols.append(i), vals.append(1.0)
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))


def prolongation_matrix(dim, degree, fine_cell_dofs, coarse_cell_dofs, children, n_fine, n_coarse,
                        fine_constraints=None, coarse_constraints=None):
    """The matrix of MGTwoLevelTransfer::prolongate_and_add (restrict_and_add is its transpose)."""
    C = dim + 1
    P1 = _embedding_1d(degree)
    n_loc = (degree + 1) ** dim
    rows, cols, vals = [], [], []
    for ch in range(2 ** dim):
        Pc = _kron_lex([P1[(ch >> e) & 1] for e in range(dim)])  # [n_loc fine, n_loc coarse]
        li, lj = np.nonzero(np.abs(Pc) > 1e-15)
        for c in range(C):
            fd = fine_cell_dofs[children[:, ch]][:, c * n_loc:(c + 1) * n_loc]
            cd = coarse_cell_dofs[:, c * n_loc:(c + 1) * n_loc]
            rows.append(fd[:, li].reshape(-1))
            cols.append(cd[:, lj].reshape(-1))
            vals.append(np.broadcast_to(Pc[li, lj], (fd.shape[0], len(li))).reshape(-1))
    P = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows).astype(np.int64),
                                              np.concatenate(cols).astype(np.int64))), shape=(n_fine, n_coarse))
    touch = np.bincount(np.asarray(fine_cell_dofs, dtype=np.int64).reshape(-1), minlength=n_fine)
    w = np.where(touch > 0, 1.0 / np.maximum(touch, 1), 0.0)
    if fine_constraints:
        w[np.fromiter(fine_constraints.keys(), dtype=np.int64)] = 0.0
    P = sp.diags(w) @ P
    if coarse_constraints:
        P = P @ constraint_matrix(n_coarse, coarse_constraints)
    return P.tocsr()


def interpolation_matrix(dim, degree, fine_cell_dofs, coarse_cell_dofs, children, n_fine, n_coarse):
    """coarse = R fine: the fine finite-element function evaluated at the coarse support points"""
    C = dim + 1
    n = degree + 1
    n_loc = n ** dim
    nodes = go.gauss_lobatto_points(degree)
    child_1d = (nodes > 0.5).astype(int)
    R1 = np.stack([go.lagrange_tables(nodes, np.array([2 * nodes[j] - child_1d[j]]))[0][0] for j in range(n)])
    fine_cell_dofs = np.asarray(fine_cell_dofs, dtype=np.int64)
    coarse_cell_dofs = np.asarray(coarse_cell_dofs, dtype=np.int64)
    rows, cols, vals = [], [], []
    for j in range(n_loc):
        jj = [(j // n ** e) % n for e in range(dim)]
        ch = sum(child_1d[jj[e]] << e for e in range(dim))
        row = _kron_lex([R1[jj[e]][None, :] for e in range(dim)])[0]
        nz = np.nonzero(np.abs(row) > 1e-15)[0]
        for c in range(C):
            r = coarse_cell_dofs[:, c * n_loc + j]
            f = fine_cell_dofs[children[:, ch]][:, c * n_loc + nz]          # [n_coarse_cells, len(nz)]
            rows.append(np.repeat(r, len(nz)))
            cols.append(f.reshape(-1))
            vals.append(np.tile(row[nz], len(r)))
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    # a coarse dof shared by several cells gets the same row from each of them: keep one copy (an assignment,
    # not a sum)
    key = rows * np.int64(n_fine) + cols
    _, first = np.unique(key, return_index=True)
    return sp.csr_matrix((vals[first], (rows[first], cols[first])), shape=(n_coarse, n_fine))


# ---------------------------------------------------------------------------------------------------------
# multigrid
# ---------------------------------------------------------------------------------------------------------
class OracleGMG:
    """PreconditionerGMG::initialize + vmult (include/multigrid.cc:202-220, :248-590), coarse solver "direct"."""

    def __init__(self, level_ops, prolongations, get_weight, *, smoothing_range=20.0, n_iterations=5,
                 eig_n_iterations=20):
        self.ops, self.P, self.get_weight = level_ops, prolongations, get_weight
        self.lo, self.hi = min(level_ops), max(level_ops)
        self.kw = dict(smoothing_range=smoothing_range, n_iterations=n_iterations, eig_cg_n_iterations=eig_n_iterations)

    def initialize(self):
        w = self.get_weight()
        self.smoothers = {}
        for l in range(self.lo, self.hi + 1):
            self.smoothers[l] = OracleRelaxation(self.ops[l], w, self.ops[l].compute_inverse_diagonal(w), **self.kw)
        for l in range(self.lo + 1, self.hi + 1):
            self.smoothers[l].estimate_eigenvalues()
        self.coarse_matrix = self.ops[self.lo].dense_matrix(w).astype(np.float64)

    def _v(self, l, rhs):
        op, w = self.ops[l], self.get_weight()
        if l == self.lo:
            return np.linalg.solve(self.coarse_matrix, rhs.astype(np.float64)).astype(rhs.dtype)
        sm = self.smoothers[l]
        sol = sm.vmult(rhs)
        t = rhs - op.vmult(sol, w)
        sol = sol + self.P[l] @ self._v(l - 1, self.P[l].T @ t)
        t = rhs - op.vmult(sol, w)
        return sol + sm.vmult(t)

    def vmult(self, src):
        dt = self.ops[self.hi].dtype
        return self._v(self.hi, np.asarray(src).astype(dt)).astype(np.float64)


# ---------------------------------------------------------------------------------------------------------
# Krylov / Newton
# ---------------------------------------------------------------------------------------------------------
def gmres(A, M, b, *, rel_tol=1e-2, abs_tol=1e-12, max_iter=10000, basis=28):
    """right-preconditioned restarted GMRES, x0 = 0; returns (x, iterations, residual history)"""
    tol = max(rel_tol * np.linalg.norm(b), abs_tol)
    x = np.zeros_like(b)
    it, hist = 0, []
    while True:
        r = b - A(x)
        rho = np.linalg.norm(r)
        hist.append(rho)
        if rho <= tol or it >= max_iter:
            return x, it, hist
        V = [r / rho]
        H = np.zeros((basis + 1, basis))
        g = np.zeros(basis + 1)
        g[0] = rho
        cs, sn = np.zeros(basis), np.zeros(basis)
        k = 0
        done = False
        for j in range(basis):
            it += 1
            w = A(M(V[j]))
            for i in range(j + 1):
                H[i, j] = np.dot(V[i], w)
                w = w - H[i, j] * V[i]
            for i in range(j + 1):  # second pass (re-orthogonalisation)
                c = np.dot(V[i], w)
                H[i, j] += c
                w = w - c * V[i]
            H[j + 1, j] = np.linalg.norm(w)
            if H[j + 1, j] > 0:
                V.append(w / H[j + 1, j])
            else:
                V.append(w)
            for i in range(j):
                t = H[i, j]
                H[i, j] = cs[i] * t + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * t + cs[i] * H[i + 1, j]
            r_ = math.hypot(H[j, j], H[j + 1, j])
            cs[j], sn[j] = H[j, j] / r_, H[j + 1, j] / r_
            H[j, j] = r_
            g[j + 1] = -sn[j] * g[j]
            g[j] *= cs[j]
            k = j + 1
            hist.append(abs(g[k]))
            if abs(g[k]) <= tol or it >= max_iter:
                done = True
                break
        y = np.linalg.solve(np.triu(H[:k, :k]), g[:k])
        p = sum(y[i] * V[i] for i in range(k))
        x = x + M(p)
        if done:
            return x, it, hist


class _TimeNone:
    """TimeIntegratorDataNone (include/time_integration.cc:141-178): order 0, weight 0, dt = 1"""
    weights = [0.0]

    def update_dt(self, dt):
        pass


class OracleChannelDriver:
    """Driver<dim>::run for the channel (main.cc:220-1000), all on the CPU oracle."""

    def __init__(self, *, dim, degree, meshes, children, constraints_inhomogeneous, inhomogeneities, min_dx,
                 nu, c1, c2, cfl, bdf_order, consider_time_derivative=True, cell_wise_stabilization=True,
                 rel_tol=1e-2, abs_tol=1e-12, newton_inexact=False, level_dtype=np.float64,
                 operator_class=go.OracleOperator):
        """operator_class: gls_oracle.OracleOperator (numpy cell loops, the checker of the tests) or
        gls_fast.FastOracleOperator (the same operator with its cell loops in C: the timed CPU baseline)"""
        self.dim, self.cfl, self.min_dx = dim, cfl, min_dx
        self.lo, self.hi = min(meshes), max(meshes)
        self.bdf = go.OracleBDF(bdf_order) if bdf_order > 0 else _TimeNone()
        self.rel_tol, self.abs_tol, self.inexact = rel_tol, abs_tol, newton_inexact

        def make(mesh, dtype):
            return operator_class(dim=dim, degree=degree, cell_dofs=mesh.cell_dofs, n_dofs=mesh.n_dofs,
                                     cell_points=mesh.cell_points, mapping_degree=mesh.mapping_degree,
                                     constraints=mesh.constraints, nu=nu, c1=c1, c2=c2, theta=1.0, order=bdf_order,
                                     consider_time_derivative=consider_time_derivative, increment_form=True,
                                     cell_wise_stabilization=cell_wise_stabilization, dtype=dtype, path="sumfac")

        self.op = make(meshes[self.hi], np.float64)
        self.level_ops = {l: make(meshes[l], level_dtype) for l in meshes}
        self.P, self.R = {}, {}
        for l in range(self.lo + 1, self.hi + 1):
            mf, mc = meshes[l], meshes[l - 1]
            fd, cd = np.asarray(mf.cell_dofs, dtype=np.int64), np.asarray(mc.cell_dofs, dtype=np.int64)
            self.P[l] = prolongation_matrix(dim, degree, fd, cd, children[l], mf.n_dofs, mc.n_dofs,
                                            mf.constraints, mc.constraints).astype(level_dtype)
            self.R[l] = interpolation_matrix(dim, degree, fd, cd, children[l], mf.n_dofs, mc.n_dofs)
        self.gmg = OracleGMG(self.level_ops, self.P, lambda: self.bdf.weights[0])
        self.cdofs = np.array(sorted(constraints_inhomogeneous), dtype=np.int64)
        self.cvals = np.array([inhomogeneities.get(int(d), 0.0) for d in self.cdofs])
        n = meshes[self.hi].n_dofs
        self.history = [np.zeros(n) for _ in range(bdf_order + 1)]
        self._distribute(self.history[0])
        self.t, self.log = 0.0, []

    def _distribute(self, v):
        v[self.cdofs] = self.cvals

    def _to_levels(self, v):
        out = {self.hi: v.astype(self.level_ops[self.hi].dtype)}
        for l in range(self.hi, self.lo, -1):
            out[l - 1] = (self.R[l] @ out[l]).astype(self.level_ops[l - 1].dtype)
        return out

    def _setup_preconditioner(self, sol, dt):
        for l, v in self._to_levels(sol).items():
            self.level_ops[l].set_linearization_point(v, dt)
        self.gmg.initialize()

    def _residual(self, sol):
        tmp = sol.copy()
        self._distribute(tmp)
        return self.op.evaluate_residual(tmp, self.bdf.weights[0])

    def step(self):
        cur = self.history[0]
        u_max = self.op.get_max_u(cur)
        dt = self.min_dx * self.cfl / max(u_max, 1.0)
        self.bdf.update_dt(dt)
        w = self.bdf.weights
        dt_loop = dt
        if isinstance(self.bdf, _TimeNone):
            dt = 1.0  # get_current_dt() of the "none" scheme; the loop body runs once (main.cc:982-987)
        for i in range(len(self.history) - 2, -1, -1):
            self.history[i + 1] = self.history[i].copy()
        order = len(self.history) - 1
        self.op.set_previous_solution(self.history, w)
        levels = [self._to_levels(h) for h in self.history]
        for l, op in self.level_ops.items():
            op.set_previous_solution([lv[l] for lv in levels], w)
        sol = self.history[0]
        # Newton (solver_nl.cc:36-89)
        self.op.set_linearization_point(sol, dt)
        rhs = self._residual(sol)
        res = [np.linalg.norm(rhs)]
        lin = []
        it = 0
        while res[-1] > 1e-7:
            if it == 0 or not self.inexact:
                self._setup_preconditioner(sol, dt)
            rhs[self.op.constrained] = 0
            inc, n_it, _ = gmres(lambda x: self.op.vmult(x, w[0]), self.gmg.vmult, rhs, rel_tol=self.rel_tol,
                                 abs_tol=self.abs_tol)
            inc[self.op.constrained] = 0
            lin.append(n_it)
            sol += inc
            self.op.set_linearization_point(sol, dt)
            rhs = self._residual(sol)
            res.append(np.linalg.norm(rhs))
            it += 1
            if it > 30:
                raise RuntimeError("Newton iteration did not converge")
        self._distribute(sol)
        self.t += dt_loop
        rec = dict(t=self.t, dt=dt_loop, u_max=u_max, newton_iterations=it, newton_residuals=res, linear_iterations=lin)
        self.log.append(rec)
        return rec
