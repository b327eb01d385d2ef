"""Second, independently written restatement: NavierStokesOperatorMatrixBased
(/root/reference/include/operator_ns.cc:1600-1756), the FEValues-style assembly of the one-step-theta
fixed-point system the reference itself keeps as its cross-check for the matrix-free operator
(SURVEY.md section 4, row 2).

TEST INFRASTRUCTURE ONLY (see oracle/gls_oracle.py).  PARITY UNPINNED like the rest of oracle/: the
reference holds no golden vectors.  What this file adds is an anchor that shares NO code with
gls_oracle.py: its own Gauss / Gauss-Lobatto points (Golub-Welsch eigenvalues, Newton on (1-x^2)P'_p),
its own Lagrange polynomials (explicit products), its own mapping Jacobians, vector-valued shape
functions held as explicit tensors (value, gradient, divergence, symmetric gradient per local dof, the
way FEValuesViews::Vector / ::Scalar hand them out), and the weak form typed term by term from
operator_ns.cc:1706-1740.  tests/test_matrix_based.py asserts, on the same mesh and vectors,

    diag(tau I_u, I_p) . A_mf  ==  A_mb            (fixed-point branch, theta scheme, weight 1/tau)
    -diag(tau I_u, I_p) . residual_mf(u)  ==  A_mb u - rhs_mb

where A_mf / residual_mf come from gls_oracle.OracleOperator (operator_ns.cc:955-1066).
"""
from __future__ import annotations

import numpy as np


# ---- 1-D ingredients, written without looking at gls_oracle.py ------------------------------------------

def _legendre(n, x):
    """P_n(x) and P'_n(x) on [-1, 1] by the three-term recurrence."""
    x = np.asarray(x, dtype=np.float64)
    p0, p1 = np.ones_like(x), x.copy()
    if n == 0:
        return p0, np.zeros_like(x)
    for k in range(2, n + 1):
        p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
    dp = n * (x * p1 - p0) / (x * x - 1.0) if n > 0 else np.zeros_like(x)
    return p1, dp


def gauss_unit(n):
    """n-point Gauss-Legendre rule on [0, 1] (Golub-Welsch)."""
    k = np.arange(1, n)
    beta = k / np.sqrt(4.0 * k * k - 1.0)
    T = np.diag(beta, 1) + np.diag(beta, -1)
    ev, evec = np.linalg.eigh(T)
    w = 2.0 * evec[0, :] ** 2
    # one Newton step polishes the eigenvalues to round-off
    for _ in range(2):
        p, dp = _legendre(n, ev)
        ev = ev - p / dp
    _, dp = _legendre(n, ev)
    w = 2.0 / ((1.0 - ev * ev) * dp * dp)
    return 0.5 * (ev + 1.0), 0.5 * w


def lobatto_unit(p):
    """p+1 Gauss-Lobatto points on [0, 1]: end points and the roots of P'_p."""
    if p == 1:
        return np.array([0.0, 1.0])
    x = -np.cos(np.pi * np.arange(p + 1) / p)  # Chebyshev-Lobatto start
    xi = x[1:-1].copy()
    for _ in range(100):
        # f = P'_p, f' from the Legendre ODE: (1-x^2) P'' = 2x P' - p(p+1) P
        P, dP = _legendre(p, xi)
        ddP = (2.0 * xi * dP - p * (p + 1) * P) / (1.0 - xi * xi)
        step = dP / ddP
        xi = xi - step
        if np.max(np.abs(step)) < 1e-16:
            break
    return 0.5 * (np.concatenate([[-1.0], xi, [1.0]]) + 1.0)


def lagrange_1d(nodes, x):
    """values L[i, q] = l_i(x_q) and derivatives dL[i, q] of the Lagrange polynomials on `nodes`."""
    n, m = len(nodes), len(x)
    L, dL = np.zeros((n, m)), np.zeros((n, m))
    for i in range(n):
        for q in range(m):
            v = 1.0
            for j in range(n):
                if j != i:
                    v *= (x[q] - nodes[j]) / (nodes[i] - nodes[j])
            L[i, q] = v
            d = 0.0
            for s in range(n):
                if s == i:
                    continue
                t = 1.0 / (nodes[i] - nodes[s])
                for j in range(n):
                    if j != i and j != s:
                        t *= (x[q] - nodes[j]) / (nodes[i] - nodes[j])
                d += t
            dL[i, q] = d
    return L, dL


def tensor_shapes(dim, degree, xq):
    """scalar tensor-product shape functions at the tensor quadrature points, lexicographic (x fastest)
    in both the dof and the point index: N[a, q], dN[a, q, e] (reference derivatives)."""
    L, dL = lagrange_1d(lobatto_unit(degree), xq)
    n, m = L.shape
    N = np.zeros((n ** dim, m ** dim))
    dN = np.zeros((n ** dim, m ** dim, dim))
    for a in range(n ** dim):
        ia = [(a // n ** e) % n for e in range(dim)]
        for q in range(m ** dim):
            iq = [(q // m ** e) % m for e in range(dim)]
            v = 1.0
            for e in range(dim):
                v *= L[ia[e], iq[e]]
            N[a, q] = v
            for e in range(dim):
                g = 1.0
                for f in range(dim):
                    g *= dL[ia[f], iq[f]] if f == e else L[ia[f], iq[f]]
                dN[a, q, e] = g
    return N, dN


class MatrixBasedOperator:
    """NavierStokesOperatorMatrixBased: system matrix and right-hand side of the theta scheme
    (operator_ns.cc:1600-1756).  Mesh arrays as in gls_oracle.OracleOperator (cell_dofs component-blocked
    lexicographic, cell_points = MappingQ support points)."""

    def __init__(self, *, dim, degree, cell_dofs, n_dofs, cell_points, mapping_degree, nu, c1, c2, theta):
        self.dim, self.degree, self.n_dofs = dim, degree, int(n_dofs)
        self.cell_dofs = np.asarray(cell_dofs, dtype=np.int64)
        self.nu, self.c1, self.c2, self.theta = nu, c1, c2, theta
        xq, wq = gauss_unit(degree + 1)
        self.N, self.dN = tensor_shapes(dim, degree, xq)
        nq = len(xq) ** dim
        self.w = np.array([np.prod([wq[(q // len(xq) ** e) % len(xq)] for e in range(dim)]) for q in range(nq)])
        self.Nm, self.dNm = tensor_shapes(dim, mapping_degree, xq)
        self.points = np.asarray(cell_points, dtype=np.float64)
        n1 = mapping_degree + 1
        self.vertex_ids = [sum(((v >> e) & 1) * (n1 - 1) * n1 ** e for e in range(dim)) for v in range(2 ** dim)]

    def _fe_values(self, k):
        """what FEValues::reinit(cell) provides: physical gradients of the scalar shapes and JxW."""
        X = self.points[k]                                  # [m, dim]
        J = np.einsum("mi,mqe->qie", X, self.dNm)           # J[q, i, e] = d x_i / d xi_e
        JxW = np.linalg.det(J) * self.w
        nq = J.shape[0]
        gphys = np.zeros((self.N.shape[0], nq, self.dim))
        for q in range(nq):
            # grad_x phi = J^-T grad_xi phi
            gphys[:, q, :] = np.linalg.solve(J[q].T, self.dN[:, q, :].T).T
        return gphys, JxW

    def _views(self, gphys):
        """vector-valued shape functions of FESystem(FE_Q^dim, FE_Q), local dof i = c * n_loc + a"""
        d, (nl, nq) = self.dim, self.N.shape
        nd = (d + 1) * nl
        V = np.zeros((nd, nq, d))
        GV = np.zeros((nd, nq, d, d))
        Q = np.zeros((nd, nq))
        GQ = np.zeros((nd, nq, d))
        for c in range(d):
            V[c * nl:(c + 1) * nl, :, c] = self.N
            GV[c * nl:(c + 1) * nl, :, c, :] = gphys
        Q[d * nl:] = self.N
        GQ[d * nl:] = gphys
        DIV = np.einsum("iqcc->iq", GV)
        EPS = 0.5 * (GV + GV.transpose(0, 1, 3, 2))
        return V, GV, DIV, EPS, Q, GQ

    def assemble(self, u_0, u_star, tau):
        """returns (A, rhs): dense system matrix and vector WITHOUT constraints applied
        (AffineConstraints::distribute_local_to_global is left to the caller)."""
        d, theta, nu = self.dim, self.theta, self.nu
        A = np.zeros((self.n_dofs, self.n_dofs))
        rhs = np.zeros(self.n_dofs)
        for k in range(self.cell_dofs.shape[0]):
            gphys, JxW = self._fe_values(k)
            V, GV, DIV, EPS, Q, GQ = self._views(gphys)
            ids = self.cell_dofs[k]
            l0, ls = u_0[ids], u_star[ids]
            # get_function_values / gradients / divergences of the two vectors
            u0 = np.einsum("i,iqc->qc", l0, V)
            us = np.einsum("i,iqc->qc", ls, V)
            g_u0 = np.einsum("i,iqcd->qcd", l0, GV)
            div_u0 = np.einsum("i,iq->q", l0, DIV)
            g_p0 = np.einsum("i,iqd->qd", l0, GQ)
            # cell-wise stabilization parameters (operator_ns.cc:1670-1690)
            verts = self.points[k][self.vertex_ids]
            h = min(np.linalg.norm(verts[a] - verts[b]) for a in range(len(verts)) for b in range(a + 1, len(verts)))
            u_max = max(np.linalg.norm(u0[q]) for q in range(u0.shape[0]))
            if nu < h:
                delta_1 = self.c1 / np.sqrt(1.0 / (tau * tau) + u_max * u_max / (h * h))
                delta_2 = self.c2 * h
            else:
                delta_1 = self.c1 * h * h
                delta_2 = self.c2 * h * h
            gv_us = np.einsum("iqcd,qd->iqc", GV, us)       # grad_u_j * u_star  (and grad_v_i * u_star)
            lhs = np.einsum("jqc,iqc,q->ij", V, V, JxW)                                              # a
            lhs += theta * tau * np.einsum("jqc,iqc,q->ij", gv_us, V, JxW)                            # b
            lhs -= tau * np.einsum("jq,iq,q->ij", Q, DIV, JxW)                                        # c
            lhs += theta * tau * 2.0 * nu * np.einsum("jqcd,iqcd,q->ij", EPS, EPS, JxW)               # d
            lhs += theta * tau * delta_1 * np.einsum("jqc,iqc,q->ij", gv_us + GQ, gv_us, JxW)         # e
            lhs += theta * tau * delta_2 * np.einsum("jq,iq,q->ij", DIV, DIV, JxW)                    # f
            lhs += theta * np.einsum("jq,iq,q->ij", DIV, Q, JxW)                                      # pressure a
            lhs += delta_1 * np.einsum("jqc,iqc,q->ij", GQ + theta * gv_us, GQ, JxW)                  # pressure b
            gu0_us = np.einsum("qcd,qd->qc", g_u0, us)
            sym0 = 0.5 * (g_u0 + g_u0.transpose(0, 2, 1))
            r = np.einsum("qc,iqc,q->i", u0, V, JxW)                                                  # a
            r -= (1.0 - theta) * tau * np.einsum("qc,iqc,q->i", gu0_us, V, JxW)                       # b
            r -= (1.0 - theta) * tau * 2.0 * nu * np.einsum("qcd,iqcd,q->i", sym0, EPS, JxW)          # d
            r -= (1.0 - theta) * tau * delta_1 * np.einsum("qc,iqc,q->i", gu0_us + g_p0, gv_us, JxW)  # e
            r -= (1.0 - theta) * tau * delta_2 * np.einsum("q,iq,q->i", div_u0, DIV, JxW)             # f
            r -= (1.0 - theta) * np.einsum("q,iq,q->i", div_u0, Q, JxW)                               # pressure a
            r -= delta_1 * (1.0 - theta) * np.einsum("qc,iqc,q->i", gu0_us, GQ, JxW)                  # pressure b
            A[np.ix_(ids, ids)] += lhs
            np.add.at(rhs, ids, r)
        return A, rhs
