"""CPU oracle for the GLS/SUPG-PSPG Navier-Stokes matrix-free operator.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
``dealii_ns_gls_b200`` / ``libglsb200.so``) may import, link or call this file;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker.

PARITY: PINNED FOR THE REFERENCE'S OWN ARITHMETIC, UNPINNED FOR DEAL.II'S.  The reference
(peterrum/dealii-ns-gls) has no tests, golden vectors or fixtures, and the operator as a whole cannot be built
here (deal.II >= 9.6 with p4est / Trilinos / MPI, not vendored, not installable).  What does compile without
deal.II is the code the reference itself contributes to the path, and it is run here as object code
(oracle/_ref/, built from /root/reference by `make -C oracle _ref`, recorded in tests/golden/reference*):
  * include/time_integration.cc, the whole file       -> OracleBDF below is bit-equal to it
    (tests/test_reference_time_integration.py);
  * include/operator_ns.cc:880-1182 (do_vmult_cell, both branches, symm_scalar_product_add), :1195-1301
    (do_vmult_boundary), :348-421 (the cell loop of compute_penalty_parameters) and :428-457 (beta of the
    outflow faces), compiled unmodified on
    stand-in Tensor / VectorizedArray / FEEvaluation types  ->  _cell_newton, _cell_fixed_point, _face_qpoint
    and _penalty below agree with them to <= 4e-15 in every branch / flag combination
    (tests/test_reference_qpoint.py).
Everything deal.II computes around those kernels -- FE_Q / QGauss tables, sum-factorised evaluate / integrate,
MappingQ geometry, constraint resolution in read_dof_values / distribute_local_to_global, compute_diagonal --
is a restatement from the published deal.II conventions listed in SURVEY.md section 8c and stays unpinned; it is
checked against closed-form answers (tests/test_oracle.py: Q1 element matrices, polynomial exactness,
finite-difference Jacobian consistency, naive-full-tensor vs sum-factorised evaluation, unit-vector matrix vs
vmult, 1/diag vs the assembled diagonal), against an independent FEValues-style restatement of the reference's
matrix-based operator (oracle/gls_matrix_based.py) and against published benchmark numbers (oracle/gls_turek.py).

Reference files restated (all relative to /root/reference):
  include/operator_ns.cc:899-916    symm_scalar_product_add
  include/operator_ns.cc:949-1066   do_vmult_cell, fixed-point / residual branch
  include/operator_ns.cc:1067-1182  do_vmult_cell, Newton (increment) branch
  include/operator_ns.cc:322-420    compute_penalty_parameters (delta_1, delta_2)
  include/operator_ns.cc:234-320    set_previous_solution
  include/operator_ns.cc:570-620    set_linearization_point
  include/operator_ns.cc:684-732    vmult (constrained rows = identity)
  include/operator_ns.cc:622-682    evaluate_rhs / evaluate_residual
  include/operator_ns.cc:195-225    compute_inverse_diagonal (+ 1e-10 guard)
  include/operator_ns.cc:530-568    get_max_u
  include/operator_ns.cc:1195-1301  do_vmult_boundary (outflow "cut" and Nitsche faces), :423-521 its tables
  include/time_integration.cc:61-91,100-107,141-178  BDF / theta / none weights

Two independent evaluation paths are provided: ``path="naive"`` uses full
(n_q x n_dofs) tensor shape tables (the way NavierStokesOperatorMatrixBased,
operator_ns.cc:1660-1744, works with FEValues), ``path="sumfac"`` uses 1-D
sweeps with the collocation derivative (the way FEEvaluation works).
"""
from __future__ import annotations

import math
import numpy as np

# --------------------------------------------------------------------------- #
# 1-D basis: FE_Q(p) = Lagrange on Gauss-Lobatto points of [0,1]; QGauss(n)
# --------------------------------------------------------------------------- #


def gauss_lobatto_points(p: int) -> np.ndarray:
    """p+1 Gauss-Lobatto points on [0,1] (support points of deal.II FE_Q(p))."""
    if p == 0:
        return np.array([0.5])
    if p == 1:
        return np.array([0.0, 1.0])
    # interior points: roots of P'_p on [-1,1]
    c = np.zeros(p + 1)
    c[p] = 1.0
    dc = np.polynomial.legendre.legder(c)
    x = np.sort(np.real(np.polynomial.legendre.legroots(dc)))
    # Newton polish in extended precision
    x = x.astype(np.longdouble)
    ddc = np.polynomial.legendre.legder(dc)
    for _ in range(3):
        f = np.polynomial.legendre.legval(x, dc)
        df = np.polynomial.legendre.legval(x, ddc)
        x = x - f / df
    x = np.concatenate([[-1.0], np.asarray(x, dtype=np.float64), [1.0]])
    x = 0.5 * (x - x[::-1])  # symmetrise
    return 0.5 * (x + 1.0)


def gauss_points_weights(n: int):
    """n-point Gauss-Legendre rule on [0,1] (deal.II QGauss<1>(n))."""
    x, w = np.polynomial.legendre.leggauss(n)
    x = 0.5 * (x - x[::-1])
    w = 0.5 * (w + w[::-1])
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_tables(nodes: np.ndarray, x: np.ndarray):
    """values[q,i] = phi_i(x_q), derivs[q,i] = phi_i'(x_q) for the Lagrange basis on nodes."""
    n = len(nodes)
    vals = np.ones((len(x), n))
    ders = np.zeros((len(x), n))
    for i in range(n):
        for m in range(n):
            if m != i:
                vals[:, i] *= (x - nodes[m]) / (nodes[i] - nodes[m])
        for l in range(n):
            if l == i:
                continue
            t = np.ones(len(x)) / (nodes[i] - nodes[l])
            for m in range(n):
                if m != i and m != l:
                    t *= (x - nodes[m]) / (nodes[i] - nodes[m])
            ders[:, i] += t
    return vals, ders


class Basis1D:
    def __init__(self, degree: int, n_q: int | None = None):
        self.degree = degree
        self.n = degree + 1
        self.n_q = self.n if n_q is None else n_q
        self.nodes = gauss_lobatto_points(degree)
        self.xq, self.wq = gauss_points_weights(self.n_q)
        self.S, self.G = lagrange_tables(self.nodes, self.xq)  # [q, i]
        _, self.D = lagrange_tables(self.xq, self.xq)  # collocation derivative [q, q']


def _kron_all(mats):
    """mats given fastest direction first (x, y, z); returns kron(z, y, x)."""
    out = mats[0]
    for m in mats[1:]:
        out = np.kron(m, out)
    return out


class TensorBasis:
    """dim-dimensional tensor-product tables, lexicographic (x fastest)."""

    def __init__(self, dim: int, degree: int, n_q: int | None = None):
        self.dim = dim
        self.b = Basis1D(degree, n_q)
        b = self.b
        self.n_dofs = b.n ** dim
        self.n_q = b.n_q ** dim
        self.N = _kron_all([b.S] * dim)  # [q, i]
        self.dN = []
        for e in range(dim):
            mats = [b.S] * dim
            mats[e] = b.G
            self.dN.append(_kron_all(mats))
        self.dN = np.stack(self.dN)  # [e, q, i]
        self.w = _kron_all([b.wq.reshape(-1, 1)] * dim).reshape(-1)  # [q]
        # reference coordinates of q points [q, dim]
        grids = np.meshgrid(*([b.xq] * dim), indexing="ij")  # index order (x, y, z)
        # lexicographic with x fastest: flatten in Fortran order
        self.xq = np.stack([g.reshape(-1, order="F") for g in grids], axis=1)


# --------------------------------------------------------------------------- #
# time integrator weights (include/time_integration.cc)
# --------------------------------------------------------------------------- #


class OracleBDF:
    """include/time_integration.cc:4-91."""

    def __init__(self, order):
        self.order = order
        self.dt = [0.0] * order
        self.weights = [0.0] * (order + 1)

    def update_dt(self, dt_new):
        for i in range(self.order - 2, -1, -1):
            self.dt[i + 1] = self.dt[i]
        self.dt[0] = dt_new
        dt = self.dt
        eff = sum(1 for v in dt if v > 0)
        w = [0.0] * (self.order + 1)
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
            raise RuntimeError("Not implemented")
        self.weights = w

    primary_weight = property(lambda s: s.weights[0])
    current_dt = property(lambda s: s.dt[0])
    theta = 1.0


# --------------------------------------------------------------------------- #
# geometry
# --------------------------------------------------------------------------- #


def mapping_jacobians(cell_points: np.ndarray, mapping_degree: int, tb: TensorBasis):
    """J[k,q,i,e] = d x_i / d xi_e at the quadrature points of each cell.

    cell_points[k, m, i]: mapping support points (lexicographic, Gauss-Lobatto
    placed, (mapping_degree+1)^dim of them) -- what MappingQ(k) interpolates.
    """
    dim = tb.dim
    mb = Basis1D(mapping_degree, tb.b.n_q)
    dM = []
    for e in range(dim):
        mats = [mb.S] * dim
        mats[e] = mb.G
        dM.append(_kron_all(mats))
    dM = np.stack(dM)  # [e, q, m]
    # J[k, q, i, e] = sum_m dM[e, q, m] X[k, m, i]: one matrix product per direction (BLAS) instead of an einsum
    K, nq = cell_points.shape[0], dM.shape[1]
    J = np.empty((K, nq, dim, dim))
    for e in range(dim):
        J[:, :, :, e] = np.matmul(dM[e][None, :, :], cell_points)
    return J


def small_inverse_and_determinant(J):
    """inverse and determinant of a stack of 2 x 2 or 3 x 3 matrices [..., d, d] in closed form (adjugate / det);
    np.linalg for other sizes"""
    d = J.shape[-1]
    if d == 2:
        det = J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]
        inv = np.empty_like(J)
        inv[..., 0, 0], inv[..., 0, 1] = J[..., 1, 1], -J[..., 0, 1]
        inv[..., 1, 0], inv[..., 1, 1] = -J[..., 1, 0], J[..., 0, 0]
        return inv / det[..., None, None], det
    if d == 3:
        a = J
        c00 = a[..., 1, 1] * a[..., 2, 2] - a[..., 1, 2] * a[..., 2, 1]
        c01 = a[..., 1, 2] * a[..., 2, 0] - a[..., 1, 0] * a[..., 2, 2]
        c02 = a[..., 1, 0] * a[..., 2, 1] - a[..., 1, 1] * a[..., 2, 0]
        det = a[..., 0, 0] * c00 + a[..., 0, 1] * c01 + a[..., 0, 2] * c02
        inv = np.empty_like(J)
        inv[..., 0, 0], inv[..., 1, 0], inv[..., 2, 0] = c00, c01, c02
        inv[..., 0, 1] = a[..., 0, 2] * a[..., 2, 1] - a[..., 0, 1] * a[..., 2, 2]
        inv[..., 0, 2] = a[..., 0, 1] * a[..., 1, 2] - a[..., 0, 2] * a[..., 1, 1]
        inv[..., 1, 1] = a[..., 0, 0] * a[..., 2, 2] - a[..., 0, 2] * a[..., 2, 0]
        inv[..., 1, 2] = a[..., 0, 2] * a[..., 1, 0] - a[..., 0, 0] * a[..., 1, 2]
        inv[..., 2, 1] = a[..., 0, 1] * a[..., 2, 0] - a[..., 0, 0] * a[..., 2, 1]
        inv[..., 2, 2] = a[..., 0, 0] * a[..., 1, 1] - a[..., 0, 1] * a[..., 1, 0]
        return inv / det[..., None, None], det
    return np.linalg.inv(J), np.linalg.det(J)


def cell_vertices(cell_points: np.ndarray, mapping_degree: int, dim: int):
    n = mapping_degree + 1
    idx = []
    for v in range(2 ** dim):
        l = 0
        for e in range(dim):
            if (v >> e) & 1:
                l += (n - 1) * n ** e
        idx.append(l)
    return cell_points[:, idx, :]


def minimum_vertex_distance(verts: np.ndarray):
    nv = verts.shape[1]
    h = np.full(verts.shape[0], np.inf)
    for a in range(nv):
        for b in range(a + 1, nv):
            h = np.minimum(h, np.linalg.norm(verts[:, a] - verts[:, b], axis=1))
    return h


def vertex_measure(verts: np.ndarray, dim: int):
    """cell->measure(): volume of the (multi)linear cell spanned by the vertices."""
    tb = TensorBasis(dim, 1, 2)
    J = mapping_jacobians(verts, 1, tb)
    return np.einsum("kq,q->k", np.linalg.det(J), tb.w)


# --------------------------------------------------------------------------- #
# the operator
# --------------------------------------------------------------------------- #


class OracleOperator:
    """Restatement of NavierStokesOperator<dim, Number> (include/operator_ns.cc).

    mesh inputs (plain arrays):
      dim, degree
      cell_dofs[k, C*n^dim]   plain vector indices, component-blocked lexicographic
      n_dofs                  vector length
      cell_points[k, m, dim]  mapping support points, mapping_degree
      constraints             dict {dof: [(master, weight), ...]} (homogeneous)
    """

    def __init__(self, *, dim, degree, cell_dofs, n_dofs, cell_points, mapping_degree,
                 constraints=None, nu, c1, c2, theta=1.0, order=1,
                 consider_time_derivative, increment_form, cell_wise_stabilization,
                 dtype=np.float64, path="naive"):
        self.dim = dim
        self.degree = degree
        self.C = dim + 1
        self.dtype = np.dtype(dtype)
        self.tb = TensorBasis(dim, degree)
        self.n_loc = self.tb.n_dofs
        self.cell_dofs = np.asarray(cell_dofs, dtype=np.int64)
        self.n_cells = self.cell_dofs.shape[0]
        self.n_dofs = int(n_dofs)
        self.nu = nu
        self.c1 = c1
        self.c2 = c2
        self.theta = theta
        self.order = order
        # operator_ns.cc:97-98
        self.ctd = bool(consider_time_derivative and order > 0)
        self.increment_form = increment_form
        self.cell_wise = cell_wise_stabilization
        self.path = path
        self.constraints = dict(constraints or {})
        self.constrained = np.array(sorted(self.constraints.keys()), dtype=np.int64)

        self._cell_points = np.asarray(cell_points, dtype=np.float64)
        self._mapping_degree = mapping_degree
        self.faces = None  # boundary faces with outflow terms (set_outflow_faces)
        J = mapping_jacobians(np.asarray(cell_points, dtype=np.float64), mapping_degree, self.tb)
        Jinv, detJ = small_inverse_and_determinant(J)
        self.Jinv = Jinv.astype(self.dtype)  # [k,q,e,j] = (J^-1)_{e j}
        self.JxW = (detJ * self.tb.w[None, :]).astype(self.dtype)
        verts = cell_vertices(np.asarray(cell_points, dtype=np.float64), mapping_degree, dim)
        self.h_min = minimum_vertex_distance(verts)
        self.measure = vertex_measure(verts, dim)

        self.N = self.tb.N.astype(self.dtype)
        self.dN = self.tb.dN.astype(self.dtype)
        b = self.tb.b
        self.S1 = b.S.astype(self.dtype)
        self.D1 = b.D.astype(self.dtype)

        # q-point tables
        self.U = self.H = self.P = None
        self.o = None  # u_time_derivative_old
        self.Gold = self.gold_p = None
        self.delta1 = self.delta2 = None

        self._Cmat = None

    # ---------------- evaluation / integration ---------------- #

    def _gather(self, vec):
        return vec[self.cell_dofs].reshape(self.n_cells, self.C, self.n_loc)

    def _evaluate(self, dofs):
        """dofs[k,c,i] -> values[k,c,q], grads[k,c,j,q] (physical)."""
        if self.path == "naive":
            val = np.einsum("qi,kci->kcq", self.N, dofs)
            rg = np.einsum("eqi,kci->kceq", self.dN, dofs)
        else:
            val, rg = self._evaluate_sumfac(dofs)
        grad = np.einsum("kqej,kceq->kcjq", self.Jinv, rg)
        return val, grad

    def _integrate(self, vout, gout):
        """vout[k,c,q], gout[k,c,j,q] (physical) -> local residuals[k,c,i]."""
        v = vout * self.JxW[:, None, :]
        rg = np.einsum("kqej,kcjq->kceq", self.Jinv, gout) * self.JxW[:, None, None, :]
        if self.path == "naive":
            return np.einsum("qi,kcq->kci", self.N, v) + np.einsum("eqi,kceq->kci", self.dN, rg)
        return self._integrate_sumfac(v, rg)

    def _evaluate_sumfac(self, dofs):
        d, n = self.dim, self.tb.b.n
        K, C = dofs.shape[0], dofs.shape[1]
        x = dofs.reshape((K, C) + (n,) * d)  # axes: k, c, (z), y, x
        # interpolate to quadrature points direction by direction
        for e in range(d):
            ax = x.ndim - 1 - e
            x = np.moveaxis(np.tensordot(x, self.S1, axes=([ax], [1])), -1, ax)
        val = x.reshape(K, C, -1)
        rgs = []
        for e in range(d):
            ax = x.ndim - 1 - e
            g = np.moveaxis(np.tensordot(x, self.D1, axes=([ax], [1])), -1, ax)
            rgs.append(g.reshape(K, C, -1))
        return val, np.stack(rgs, axis=2)

    def _integrate_sumfac(self, v, rg):
        d, n = self.dim, self.tb.b.n
        K, C = v.shape[0], v.shape[1]
        x = v.reshape((K, C) + (n,) * d).copy()
        for e in range(d):
            ax = x.ndim - 1 - e
            g = rg[:, :, e, :].reshape((K, C) + (n,) * d)
            x += np.moveaxis(np.tensordot(g, self.D1, axes=([ax], [0])), -1, ax)
        for e in range(d):
            ax = x.ndim - 1 - e
            x = np.moveaxis(np.tensordot(x, self.S1, axes=([ax], [0])), -1, ax)
        return x.reshape(K, C, -1)

    def _scatter(self, loc):
        dst = np.zeros(self.n_dofs, dtype=self.dtype)
        np.add.at(dst, self.cell_dofs.reshape(-1), loc.reshape(-1))
        return dst

    # ---------------- constraints ---------------- #

    def _constraint_matrix(self):
        if self._Cmat is None:
            import scipy.sparse as sp
            rows, cols, vals = [], [], []
            is_c = np.zeros(self.n_dofs, dtype=bool)
            is_c[self.constrained] = True
            free = np.nonzero(~is_c)[0]
            rows.extend(free.tolist())
            cols.extend(free.tolist())
            vals.extend([1.0] * len(free))
            for dof, entries in self.constraints.items():
                for m, w in entries:
                    rows.append(dof)
                    cols.append(m)
                    vals.append(w)
            self._Cmat = sp.csr_matrix((np.array(vals, dtype=self.dtype), (rows, cols)),
                                       shape=(self.n_dofs, self.n_dofs))
        return self._Cmat

    def _resolve(self, src):
        """read_dof_values: constrained entries replaced by sum w*master (0 if none)."""
        if len(self.constrained) == 0:
            return src
        return self._constraint_matrix() @ src

    def _distribute_transpose(self, dst):
        """distribute_local_to_global: C^T, nothing lands on constrained rows."""
        if len(self.constrained) == 0:
            return dst
        return self._constraint_matrix().T @ dst

    # ---------------- tables ---------------- #

    def set_previous_solution(self, history, weights):
        """operator_ns.cc:234-320. history[0] is the current solution slot."""
        if self.order == 0:
            return
        vec_old = np.zeros(self.n_dofs, dtype=self.dtype)
        for i in range(1, self.order + 1):
            vec_old = vec_old + self.dtype.type(weights[i]) * np.asarray(history[i], dtype=self.dtype)
        val, _ = self._evaluate(self._gather(vec_old))
        self.o = val[:, : self.dim, :].copy()
        if self.theta != 1.0:
            _, grad = self._evaluate(self._gather(np.asarray(history[1], dtype=self.dtype)))
            self.Gold = grad[:, : self.dim].copy()
            self.gold_p = grad[:, self.dim].copy()

    def set_linearization_point(self, vec, dt):
        """operator_ns.cc:570-620 followed by compute_penalty_parameters :322-420."""
        vec = np.asarray(vec, dtype=self.dtype)
        val, grad = self._evaluate(self._gather(vec))
        d = self.dim
        self.U = val[:, :d, :].copy()
        self.H = grad[:, :d].copy()
        self.P = grad[:, d].copy()
        self._penalty(val[:, :d, :], dt)
        if self.faces is not None:
            self._set_face_velocity(vec)

    def _penalty(self, u, dt):
        d = self.dim
        stau = 0.0 if dt == 0.0 else 1.0 / dt
        umag = np.sqrt(np.sum(u.astype(self.dtype) ** 2, axis=1))  # [k,q], Number precision
        u_max = umag.max(axis=1).astype(np.float64)
        h = self.h_min
        d1 = np.where(self.nu < h,
                      self.c1 / np.sqrt(stau ** 2 + u_max * u_max / (h * h)),
                      self.c1 * h * h)
        d2 = np.where(self.nu < h, self.c2 * h, self.c2 * h * h)
        self.delta1_cell = d1.astype(self.dtype)
        self.delta2_cell = d2.astype(self.dtype)
        # q-point-wise, after Lethe (operator_ns.cc:390-420); arithmetic in Number
        T = self.dtype.type
        if d == 2:
            hq = np.sqrt(4.0 * self.measure / math.pi) / self.degree
        else:
            hq = np.power(6.0 * self.measure / math.pi, 1.0 / 3.0) / self.degree
        hq = hq.astype(self.dtype)[:, None]
        u2 = T(1e-12) + np.sum(u.astype(self.dtype) ** 2, axis=1)
        nu = T(self.nu)
        self.delta1_q = (T(1.0) / np.sqrt(T(stau ** 2) + T(4.0) * u2 / hq / hq
                                          + T(9.0) * (T(4.0) * nu / (hq * hq)) ** 2)).astype(self.dtype)
        self.delta2_q = (np.sqrt(u2) * hq * T(0.5)).astype(self.dtype)

    def _deltas(self):
        if self.cell_wise:
            return self.delta1_cell[:, None, None], self.delta2_cell[:, None, None]
        return self.delta1_q[:, None, :], self.delta2_q[:, None, :]

    # ---------------- q-point physics ---------------- #

    def _symm_add(self, gout, B, factor):
        """operator_ns.cc:899-916."""
        d = self.dim
        for a in range(d):
            gout[:, a, a] += B[:, a, a] * factor
        for e in range(d):
            for a in range(e + 1, d):
                tmp = (B[:, a, e] + B[:, e, a]) * (factor * 0.5)
                gout[:, a, e] += tmp
                gout[:, e, a] += tmp

    def _cell_newton(self, val, grad, weight):
        """operator_ns.cc:1067-1182."""
        d = self.dim
        T = self.dtype.type
        w = T(weight)
        nu = T(self.nu)
        d1, d2 = self._deltas()
        u, p = val[:, :d], val[:, d]
        G, g = grad[:, :d], grad[:, d]
        U, H, P = self.U, self.H, self.P
        td = u * w
        div = sum(G[:, a, a] for a in range(d))
        sgu = np.einsum("kcjq,kjq->kcq", G, U)
        ugs = np.einsum("kcjq,kjq->kcq", H, u)
        sgs = np.einsum("kcjq,kjq->kcq", H, U)
        vout = np.zeros_like(val)
        gout = np.zeros_like(grad)
        vout[:, :d] = td + sgu + ugs
        for a in range(d):
            gout[:, a, a] -= p
        self._symm_add(gout, G, nu * T(2.0))
        r0 = g + sgu + ugs
        r1 = P + sgs
        if self.ctd:
            r0 = td + r0
            r1 = (U * w + self.o) + r1
        r0 = d1 * r0
        r1 = d1 * r1
        for a in range(d):
            for b in range(d):
                gout[:, a, b] += U[:, b] * r0[:, a] + u[:, b] * r1[:, a]
        d2div = d2[:, 0] * div if d2.ndim == 3 else d2 * div
        for a in range(d):
            gout[:, a, a] += d2div
        vout[:, d] = div
        gout[:, d] = r0
        return vout, gout

    def _cell_fixed_point(self, val, grad, weight, residual):
        """operator_ns.cc:955-1066."""
        d = self.dim
        T = self.dtype.type
        w = T(weight)
        nu = T(self.nu)
        th = T(self.theta)
        d1, d2 = self._deltas()
        u, p = val[:, :d], val[:, d]
        g = grad[:, d]
        U = self.U
        pbar = th * g
        td = u * w
        B = th * grad[:, :d]
        if residual and self.o is not None:
            td = td + self.o
        if residual and self.theta != 1.0:
            B = B + (T(1.0) - th) * self.Gold
            pbar = pbar + (T(1.0) - th) * self.gold_p
        divb = sum(B[:, a, a] for a in range(d))
        sgb = np.einsum("kcjq,kjq->kcq", B, U)
        vout = np.zeros_like(val)
        gout = np.zeros_like(grad)
        vout[:, :d] = td + sgb
        for a in range(d):
            gout[:, a, a] -= p
        self._symm_add(gout, B, nu * T(2.0))
        tdc = td if self.ctd else 0
        r0 = d1 * (tdc + pbar + sgb)
        for a in range(d):
            for b in range(d):
                gout[:, a, b] += U[:, b] * r0[:, a]
        d2div = d2[:, 0] * divb if d2.ndim == 3 else d2 * divb
        for a in range(d):
            gout[:, a, a] += d2div
        vout[:, d] = divb
        gout[:, d] = d1 * (tdc + g + sgb)
        return vout, gout

    def _apply_cells(self, loc_in, weight, residual):
        val, grad = self._evaluate(loc_in)
        if residual or not self.increment_form:
            vout, gout = self._cell_fixed_point(val, grad, weight, residual)
        else:
            vout, gout = self._cell_newton(val, grad, weight)
        out = self._integrate(vout, gout)
        if self.faces is not None:
            self._apply_faces(loc_in, out, residual)
        return out

    # ---------------- boundary faces with outflow terms ---------------- #

    def _face_tables(self, face_no, degree, n_q):
        """basis values / reference gradients at the quadrature points of face `face_no` = 2 * direction + side
        (QGauss(n_q) in the tangential directions, ascending direction fastest): N[q, i], dN[e, q, i], w[q]"""
        d = self.dim
        direction, side = face_no // 2, face_no % 2
        nodes = gauss_lobatto_points(degree)
        xq, wq = gauss_points_weights(n_q)
        St, Gt = lagrange_tables(nodes, xq)
        Sn, Gn = lagrange_tables(nodes, np.array([float(side)]))
        vals = [Sn if e == direction else St for e in range(d)]
        N = _kron_all(vals)
        dN = []
        for e in range(d):
            mats = list(vals)
            mats[e] = Gn if e == direction else Gt
            dN.append(_kron_all(mats))
        w = _kron_all([np.ones((1, 1)) if e == direction else wq.reshape(-1, 1) for e in range(d)]).reshape(-1)
        return N, np.stack(dN), w

    def set_outflow_faces(self, face_cell, face_no, face_kind, target_velocity=None):
        """Boundary faces carrying the outflow terms of do_vmult_boundary (operator_ns.cc:1195-1301):
        face_kind 1 = all_outflow_bcs_cut, 2 = all_outflow_bcs_nitsche; target_velocity[f, q, dim] for the
        Nitsche residual (:495-521).  Geometry from the mapping support points: n = J^-T n_ref / |.|,
        JxW = |det J| |J^-T n_ref| w_q; beta = 1 / h^(p+1), h after Lethe (:423-458)."""
        d, p = self.dim, self.degree
        nq1 = self.tb.b.n_q
        face_cell = np.asarray(face_cell, dtype=np.int64)
        face_no = np.asarray(face_no, dtype=np.int64)
        nf = len(face_cell)
        nqf = nq1 ** (d - 1)
        F = dict(cell=face_cell, no=face_no, kind=np.asarray(face_kind, dtype=np.int64),
                 N=np.zeros((nf, nqf, self.n_loc)), dN=np.zeros((nf, d, nqf, self.n_loc)),
                 normal=np.zeros((nf, nqf, d)), jxw=np.zeros((nf, nqf)), Jinv=np.zeros((nf, nqf, d, d)))
        for fn in np.unique(face_no):
            sel = np.nonzero(face_no == fn)[0]
            N, dN, w = self._face_tables(int(fn), p, nq1)
            _, dM, _ = self._face_tables(int(fn), self._mapping_degree, nq1)
            J = np.einsum("eqm,kmi->kqie", dM, self._cell_points[face_cell[sel]])
            Jinv = np.linalg.inv(J)  # [k, q, e, j]
            nref = np.zeros(d)
            nref[fn // 2] = 1.0 if fn % 2 else -1.0
            nn = np.einsum("kqej,e->kqj", Jinv, nref)  # J^-T n_ref
            ln = np.linalg.norm(nn, axis=2)
            F["N"][sel], F["dN"][sel] = N, dN
            F["normal"][sel] = nn / ln[:, :, None]
            F["jxw"][sel] = np.abs(np.linalg.det(J)) * ln * w[None, :]
            F["Jinv"][sel] = Jinv
        if d == 2:
            h = np.sqrt(4.0 * self.measure[face_cell] / math.pi) / p
        else:
            h = np.power(6.0 * self.measure[face_cell] / math.pi, 1.0 / 3.0) / p
        F["beta"] = (1.0 / np.power(h.astype(self.dtype), self.dtype.type(p + 1))).astype(self.dtype)
        F["target"] = None if target_velocity is None else np.asarray(target_velocity, dtype=self.dtype)
        F["velocity"] = np.zeros((nf, nqf, d), dtype=self.dtype)
        for k in ("N", "dN", "normal", "jxw", "Jinv"):
            F[k] = F[k].astype(self.dtype)
        self.faces = F

    def _set_face_velocity(self, vec):
        """face_velocity of compute_penalty_parameters (operator_ns.cc:460-476)"""
        F = self.faces
        u = self._gather(vec)[F["cell"]][:, : self.dim]  # [f, d, i]
        F["velocity"] = np.einsum("fqi,fdi->fqd", F["N"], u).astype(self.dtype)

    def _face_qpoint(self, kind, val, grad, n, beta, velocity, target, residual):
        """do_vmult_boundary at the face quadrature points (operator_ns.cc:1195-1301): what the reference hands to
        submit_value / submit_gradient.  kind[f] 1 = cut, 2 = Nitsche; val[f, d, q], grad[f, d, j, q] velocity
        values and physical gradients, n[f, q, j] outward normals, beta[f], velocity[f, q, d] (face_velocity of the
        linearization point), target[f, q, d] or None."""
        T = self.dtype.type
        beta = beta[:, None, None]
        vr = np.zeros_like(val)
        gr = np.zeros_like(grad)
        cut = kind == 1
        nit = kind == 2
        if cut.any():
            sv = val if residual else np.moveaxis(velocity, 2, 1)  # [f, d, q]
            no = np.minimum(T(0), np.einsum("fdq,fqd->fq", sv, n))
            vr[cut] = (beta * no[:, None, :] * val)[cut]
        if nit.any():
            v = val - np.moveaxis(target, 2, 1) if (residual and target is not None) else val
            gn = np.einsum("fcjq,fqj->fcq", grad, n)
            vr[nit] = (beta * v - T(self.nu) * gn)[nit]
            gr[nit] = (-T(self.nu) * v[:, :, None, :] * np.moveaxis(n, 2, 1)[:, None, :, :])[nit]
        return vr, gr

    def _apply_faces(self, loc_in, out, residual):
        F, d = self.faces, self.dim
        T = self.dtype.type
        u = loc_in[F["cell"]][:, :d]  # [f, d, i]
        val = np.einsum("fqi,fci->fcq", F["N"], u)
        rg = np.einsum("feqi,fci->fceq", F["dN"], u)
        grad = np.einsum("fqej,fceq->fcjq", F["Jinv"], rg)
        vr, gr = self._face_qpoint(F["kind"], val, grad, F["normal"], F["beta"], F["velocity"], F["target"], residual)
        vq = vr * F["jxw"][:, None, :]
        rgq = np.einsum("fqej,fcjq->fceq", F["Jinv"], gr) * F["jxw"][:, None, None, :]
        loc = np.einsum("fqi,fcq->fci", F["N"], vq) + np.einsum("feqi,fceq->fci", F["dN"], rgq)
        np.add.at(out, (F["cell"][:, None, None], np.arange(d)[None, :, None], np.arange(self.n_loc)[None, None, :]), loc)

    # ---------------- public API ---------------- #

    def vmult(self, src, weight, edge_constrained_indices=None):
        """operator_ns.cc:684-732 (no face integrals).  With edge_constrained_indices (GMG-LS,
        :692-700, :724-731): src is zeroed there for the loop and dst gets the saved src value."""
        src = np.asarray(src, dtype=self.dtype)
        if edge_constrained_indices is not None and len(edge_constrained_indices):
            e = np.asarray(edge_constrained_indices, dtype=np.int64)
            masked = src.copy()
            masked[e] = 0
            dst = self.vmult(masked, weight)
            dst[e] = src[e]
            return dst
        x = self._resolve(src)
        loc = self._apply_cells(self._gather(x), weight, residual=False)
        dst = self._distribute_transpose(self._scatter(loc))
        dst = np.asarray(dst, dtype=self.dtype)
        if len(self.constrained):
            dst[self.constrained] = src[self.constrained]
        return dst

    def vmult_interface_down(self, src, weight):
        """operator_ns.cc:734-752: the plain cell loop + identity on constrained rows."""
        return self.vmult(src, weight)

    def vmult_interface_up(self, src, weight, edge_constrained_indices, has_edge=None):
        """operator_ns.cc:754-787: A applied to src restricted to the edge indices, no identity
        on constrained rows; zero if no rank has edge indices."""
        src = np.asarray(src, dtype=self.dtype)
        e = np.asarray(edge_constrained_indices, dtype=np.int64)
        if has_edge is None:
            has_edge = len(e) > 0
        if not has_edge:
            return np.zeros_like(src)
        cpy = np.zeros_like(src)
        cpy[e] = src[e]
        loc = self._apply_cells(self._gather(self._resolve(cpy)), weight, residual=False)
        return np.asarray(self._distribute_transpose(self._scatter(loc)), dtype=self.dtype)

    def evaluate_residual(self, src_with_bc, weight):
        """operator_ns.cc:648-682; src must already carry the inhomogeneous BCs."""
        src = np.asarray(src_with_bc, dtype=self.dtype)
        loc = self._apply_cells(self._gather(src), weight, residual=True)
        dst = np.asarray(self._distribute_transpose(self._scatter(loc)), dtype=self.dtype)
        if len(self.constrained):
            dst[self.constrained] = 0
        return -dst

    def get_max_u(self, vec):
        """operator_ns.cc:530-568."""
        val, _ = self._evaluate(self._gather(np.asarray(vec, dtype=self.dtype)))
        return float(np.sqrt(np.sum(val[:, : self.dim] ** 2, axis=1)).max())

    def cell_matrices(self, weight):
        """A_cell[k, i, j] by applying do_vmult_cell<false> to unit vectors
        (operator_ns.cc:1407-1430 builds the assembled matrix the same way)."""
        nl = self.C * self.n_loc
        A = np.zeros((self.n_cells, nl, nl), dtype=self.dtype)
        for j in range(nl):
            e = np.zeros((self.n_cells, nl), dtype=self.dtype)
            e[:, j] = 1
            out = self._apply_cells(e.reshape(self.n_cells, self.C, self.n_loc), weight, False)
            A[:, :, j] = out.reshape(self.n_cells, nl)
        return A

    def dense_matrix(self, weight):
        """Global matrix of vmult by unit vectors (small meshes only)."""
        A = np.zeros((self.n_dofs, self.n_dofs), dtype=self.dtype)
        for j in range(self.n_dofs):
            e = np.zeros(self.n_dofs, dtype=self.dtype)
            e[j] = 1
            A[:, j] = self.vmult(e, weight)
        return A

    def compute_inverse_diagonal(self, weight, edge_constrained_indices=None):
        """operator_ns.cc:195-225: diag(C^T A C), 1 on constrained rows, 0 on the refinement-edge dofs of a
        GMG-LS level (:219-220), then x -> |x| > 1e-10 ? 1/x : 1."""
        A = self.cell_matrices(weight)
        nl = self.C * self.n_loc
        diag = np.zeros(self.n_dofs, dtype=self.dtype)
        if len(self.constrained) == 0:
            np.add.at(diag, self.cell_dofs.reshape(-1),
                      np.einsum("kii->ki", A).reshape(-1))
        else:
            Cm = self._constraint_matrix().tocsr()
            for k in range(self.n_cells):
                Ck = Cm[self.cell_dofs[k]].toarray()  # [nl, n_dofs] rows of C for the local dofs
                cols = np.nonzero(np.abs(Ck).sum(axis=0))[0]
                Cc = Ck[:, cols]
                diag[cols] += np.einsum("ig,ij,jg->g", Cc, A[k], Cc)
            diag[self.constrained] = 1
        if edge_constrained_indices is not None and len(edge_constrained_indices):
            diag[np.asarray(edge_constrained_indices, dtype=np.int64)] = 0
        T = self.dtype.type
        with np.errstate(divide="ignore"):
            return np.where(np.abs(diag) > T(1e-10), T(1.0) / diag, T(1.0)).astype(self.dtype)
