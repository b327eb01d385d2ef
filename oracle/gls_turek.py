"""Schaefer-Turek benchmark 2D-1 (stationary flow around a cylinder, Re = 20) solved with the CPU oracle -- a
PHYSICS known-answer test for the restatement.  TEST INFRASTRUCTURE ONLY (see oracle/gls_oracle.py).

Why it exists.  The reference holds no golden vectors and cannot be built here as a whole; its own
quadrature-point code is run as object code (tests/test_reference_qpoint.py), but everything deal.II does around
it -- basis, quadrature, evaluate / integrate, mapping, constraints -- is restated and nothing ties THAT to the
reference's output.  What the reference does hold is the input of
a benchmark with published answers: input/input_turek_2D_Re20_stat.json is the DFG benchmark 2D-1 (nu = 0.001,
parabolic inflow with u_max = 0.3 on a 2.2 x 0.41 channel, cylinder of diameter 0.1, Q2 elements, Newton,
"time intration": "none", q-point-wise stabilisation), and SimulationCylinder::postprocess
(include/simulation.cc:434-548) writes drag, lift and pressure difference for exactly this comparison.  This file
runs that configuration through gls_oracle.OracleOperator -- residual branch for the right-hand side, Newton
branch for the Jacobian (assembled from its cell matrices like operator_ns.cc:1407-1430), both with the GLS terms
switched on -- and evaluates the three functionals the way the reference does.  Published values (Schaefer &
Turek 1996 give the intervals; the digits are from John & Matthies 2001 / Nabh 1998):

    c_D = 5.57953523384     c_L = 0.010618948146     delta p = 0.11752016697

A sign error, a wrong factor in 2 nu eps(u), a wrong convective or pressure coupling, a constraint that is not
applied, a curved-cell Jacobian that is transposed, or stabilisation terms that are not consistent with the
strong residual all move these numbers far outside the tolerance of tests/test_turek_benchmark.py.  It is not a
bit-level pin against deal.II (the mesh is this file's own multi-block mesh, not GridGenerator's), it pins the
physics the restated operator discretises.

Geometry as in include/grid_cylinder.h:22-85 with the origin at the cylinder centre: channel
[-0.2, 2.0] x [-0.2, 0.21] (cylinder shift 0.005), hole of radius 0.05, boundary ids like
SimulationCylinder::get_boundary_descriptor (simulation.cc:364-420): inflow x = -0.2 (inhomogeneous Dirichlet,
InflowBoundaryValues::Channel, simulation.cc:41-67), outflow x = 2.0 ("homogeneous nbc": pressure rows zero,
main.cc:279-283), no-slip walls and cylinder.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import gls_oracle as go

LITERATURE = {"drag": 5.57953523384, "lift": 0.010618948146, "p_diff": 0.11752016697}

R_CYL, HALF, X_IN, X_OUT, Y_LO, Y_HI = 0.05, 0.1, -0.2, 2.0, -0.2, 0.21
U_MAX, NU = 0.3, 0.001


# ---------------------------------------------------------------------------------------------------------
# multi-block Q2 mesh
# ---------------------------------------------------------------------------------------------------------
def _stretch(xi, ratio):
    """monotone map of [0, 1] onto itself whose cell size grows geometrically by `ratio` from start to end"""
    if abs(ratio - 1.0) < 1e-12:
        return xi
    b = math.log(ratio)
    return np.expm1(b * xi) / math.expm1(b)


def _patches(level):
    """(map, n_xi, n_eta) of every block; maps take reference coordinates in [0, 1]^2 to the plane and are
    right-handed.  Shared edges carry the same 1-D node distribution on both sides."""
    nt = 2 * 2 ** level            # cells per side of the inner square (= per 90-degree sector)
    ns = 3 * 2 ** level // 2 + 1   # radial layers in the ring
    n_top, n_bot, n_left = max(1, nt // 2 + 1), max(1, nt // 2), max(1, nt // 2)
    n_right = 7 * 2 ** level
    patches = []

    def sector(k):
        th0 = -0.25 * math.pi + 0.5 * math.pi * k
        c0 = HALF * math.sqrt(2.0) * np.array([math.cos(th0), math.sin(th0)])
        c1 = HALF * math.sqrt(2.0) * np.array([math.cos(th0 + 0.5 * math.pi), math.sin(th0 + 0.5 * math.pi)])

        def f(s, t):
            th = th0 + 0.5 * math.pi * t
            inner = R_CYL * np.stack([np.cos(th), np.sin(th)], axis=-1)
            outer = c0[None, :] * (1.0 - t)[:, None] + c1[None, :] * t[:, None]
            g = _stretch(s, 4.0)[:, None]
            return inner * (1.0 - g) + outer * g
        return f

    for k in range(4):
        patches.append((sector(k), ns, nt))

    def rect(x0, x1, y0, y1, ratio_x=1.0):
        def f(s, t):
            return np.stack([x0 + (x1 - x0) * _stretch(s, ratio_x), y0 + (y1 - y0) * t], axis=-1)
        return f

    grow = 6.0
    patches.append((rect(-HALF, HALF, HALF, Y_HI), nt, n_top))
    patches.append((rect(-HALF, HALF, Y_LO, -HALF), nt, n_bot))
    for (y0, y1, n) in ((Y_LO, -HALF, n_bot), (-HALF, HALF, nt), (HALF, Y_HI, n_top)):
        patches.append((rect(X_IN, -HALF, y0, y1), n_left, n))
        patches.append((rect(HALF, X_OUT, y0, y1, grow), n_right, n))
    return patches


class TurekMesh:
    """cell_dofs / cell_points in the layout of gls_oracle.OracleOperator (Q2, Q2 mapping, node-major dofs)"""

    def __init__(self, level=2):
        gl = np.array([0.0, 0.5, 1.0])  # Gauss-Lobatto points of FE_Q(2) = support points of MappingQ(2)
        pts = []
        self.on_cylinder_face = []      # cells whose xi = 0 face lies on the cylinder (ring blocks, first layer)
        for b, (f, n0, n1) in enumerate(_patches(level)):
            i, j = np.meshgrid(np.arange(n0), np.arange(n1), indexing="ij")
            i, j = i.reshape(-1), j.reshape(-1)
            s = (i[:, None] + np.tile(gl, 3)[None, :]) / n0           # local node l = a + 3 b: a along xi
            t = (j[:, None] + np.repeat(gl, 3)[None, :]) / n1
            xy = f(s.reshape(-1), t.reshape(-1)).reshape(len(i), 9, 2)
            self.on_cylinder_face.append((b < 4) & (i == 0))
            pts.append(xy)
        self.cell_points = np.concatenate(pts)
        self.on_cylinder_face = np.concatenate(self.on_cylinder_face)
        self.n_cells = self.cell_points.shape[0]
        key = np.round(self.cell_points.reshape(-1, 2) * 1e9).astype(np.int64)
        _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
        self.node_xy = self.cell_points.reshape(-1, 2)[first]
        self.cell_nodes = inv.reshape(self.n_cells, 9)
        self.n_nodes = len(first)
        self.n_dofs = 3 * self.n_nodes
        self.cell_dofs = np.concatenate([3 * self.cell_nodes + c for c in range(3)], axis=1)
        x, y = self.node_xy[:, 0], self.node_xy[:, 1]
        eps = 1e-9
        self.is_inflow = np.abs(x - X_IN) < eps
        self.is_outflow = np.abs(x - X_OUT) < eps
        self.is_wall = (np.abs(y - Y_LO) < eps) | (np.abs(y - Y_HI) < eps)
        self.is_cylinder = np.abs(np.hypot(x, y) - R_CYL) < eps
        vel = self.is_inflow | self.is_wall | self.is_cylinder
        nodes_v, nodes_p = np.nonzero(vel)[0], np.nonzero(self.is_outflow)[0]
        cons = np.concatenate([3 * nodes_v, 3 * nodes_v + 1, 3 * nodes_p + 2])
        self.constraints = {int(d): [] for d in np.sort(cons)}
        # InflowBoundaryValues::Channel with no-slip walls: parabola over the channel height H = 0.41
        H = Y_HI - Y_LO
        inl = np.nonzero(self.is_inflow & ~self.is_wall)[0]
        yy = y[inl] - Y_LO
        self.inhomogeneities = {int(3 * n): float(U_MAX * 4.0 * v * (H - v) / H / H) for n, v in zip(inl, yy)}


# ---------------------------------------------------------------------------------------------------------
# stationary Newton solve (solver_nl.cc:36-89 with a sparse direct solve in place of GMRES + GMG)
# ---------------------------------------------------------------------------------------------------------
def make_operator(mesh: TurekMesh, nu=NU):
    """the flags of input_turek_2D_Re20_stat.json: Newton (increment form), time integration "none" (order 0,
    weight 0, dt = 1), q-point-wise stabilisation"""
    return go.OracleOperator(dim=2, degree=2, cell_dofs=mesh.cell_dofs, n_dofs=mesh.n_dofs,
                             cell_points=mesh.cell_points, mapping_degree=2, constraints=mesh.constraints, nu=nu,
                             c1=1.0, c2=1.0, theta=1.0, order=0, consider_time_derivative=True, increment_form=True,
                             cell_wise_stabilization=False, path="sumfac")


def system_matrix(op, weight=0.0):
    """get_system_matrix (operator_ns.cc:1303-1434): cell matrices of the Newton branch assembled with the
    zero constraints (rows and columns dropped, 1 on the diagonal)"""
    A = op.cell_matrices(weight)
    nl = A.shape[1]
    rows = np.repeat(op.cell_dofs[:, :, None], nl, axis=2).reshape(-1)
    cols = np.repeat(op.cell_dofs[:, None, :], nl, axis=1).reshape(-1)
    free = np.ones(op.n_dofs, dtype=bool)
    free[op.constrained] = False
    keep = free[rows] & free[cols]
    J = sp.coo_matrix((A.reshape(-1)[keep], (rows[keep], cols[keep])), shape=(op.n_dofs, op.n_dofs)).tocsr()
    return (J + sp.diags((~free).astype(np.float64))).tocsc()


def solve_stationary(mesh: TurekMesh, nu=NU, tol=1e-10, max_it=25, log=None):
    op = make_operator(mesh, nu)
    x = np.zeros(mesh.n_dofs)
    for d, v in mesh.inhomogeneities.items():
        x[d] = v
    history = []
    for it in range(max_it + 1):
        op.set_linearization_point(x, 1.0)
        rhs = op.evaluate_residual(x, 0.0)
        history.append(float(np.linalg.norm(rhs)))
        if log is not None:
            log(f"    [N] step {it} ; residual = {history[-1]:.6e}")
        if history[-1] < tol:
            return x, history
        inc = spla.spsolve(system_matrix(op), rhs)
        inc[op.constrained] = 0.0
        x = x + inc
    raise RuntimeError(f"Newton iteration did not converge: {history}")


# ---------------------------------------------------------------------------------------------------------
# SimulationCylinder::postprocess (simulation.cc:434-548)
# ---------------------------------------------------------------------------------------------------------
def drag_lift_pressure(mesh: TurekMesh, x, nu=NU):
    gl = np.array([0.0, 0.5, 1.0])
    xq, wq = go.gauss_points_weights(3)                     # QGauss<dim - 1>(3) on the face
    N0, D0 = go.lagrange_tables(gl, np.array([0.0]))        # xi = 0: the face on the cylinder
    N1, D1 = go.lagrange_tables(gl, xq)
    drag = lift = 0.0
    for k in np.nonzero(mesh.on_cylinder_face)[0]:
        X = mesh.cell_points[k]                             # [9, 2], local node a + 3 b
        u = x[mesh.cell_dofs[k]].reshape(3, 9)
        for q in range(len(xq)):
            n = np.outer(N1[q], N0[0]).reshape(-1)          # [b, a] -> a + 3 b
            dxi = np.outer(N1[q], D0[0]).reshape(-1)
            deta = np.outer(D1[q], N0[0]).reshape(-1)
            J = np.stack([dxi @ X, deta @ X], axis=1)       # J[i, e] = d x_i / d xi_e
            Jinv = np.linalg.inv(J)
            gphys = np.stack([dxi, deta], axis=1) @ Jinv    # [9, 2]: d N / d x_j
            grad_u = u[:2] @ gphys                          # [c, j]
            p = u[2] @ n
            stress = -p * np.eye(2) + nu * (grad_u + grad_u.T)
            normal = Jinv[0] / np.linalg.norm(Jinv[0])      # grad xi: from the cylinder into the fluid
            jxw = np.linalg.norm(J[:, 1]) * wq[q]
            f = stress @ normal
            drag += f[0] * jxw
            lift += f[1] * jxw
    u_bar = U_MAX * 2.0 / 3.0
    scaling = 2.0 / (2.0 * R_CYL) / u_bar ** 2

    def p_at(px, py):
        n = np.nonzero((np.abs(mesh.node_xy[:, 0] - px) < 1e-9) & (np.abs(mesh.node_xy[:, 1] - py) < 1e-9))[0]
        assert len(n) == 1
        return x[3 * n[0] + 2]
    return {"drag": drag * scaling, "lift": lift * scaling, "p_diff": p_at(-R_CYL, 0.0) - p_at(R_CYL, 0.0)}


def consistent_forces(mesh: TurekMesh, x, nu=NU):
    """Drag and lift from the rows of the weak form that belong to the cylinder's velocity dofs (the "consistent
    nodal forces": the residual branch tested with the discrete function that is 1 on the cylinder and 0 on all
    other nodes; integration by parts turns it into the traction integral).  Not what the reference prints -- it is
    here because it converges faster than the boundary integral and so shows that the discrete SOLUTION, not
    only its boundary gradients, approaches the published values."""
    op = make_operator(mesh, nu)
    op.set_linearization_point(x, 1.0)
    rows = np.asarray(op._scatter(op._apply_cells(op._gather(x), 0.0, residual=True)))
    cyl = np.nonzero(mesh.is_cylinder)[0]
    scaling = 2.0 / (2.0 * R_CYL) / (U_MAX * 2.0 / 3.0) ** 2
    return {"drag": float(-rows[3 * cyl].sum() * scaling), "lift": float(-rows[3 * cyl + 1].sum() * scaling)}


def run(level=2, log=None):
    mesh = TurekMesh(level)
    x, history = solve_stationary(mesh, log=log)
    out = {k: float(v) for k, v in drag_lift_pressure(mesh, x).items()}
    cf = consistent_forces(mesh, x)
    out.update(drag_consistent=cf["drag"], lift_consistent=cf["lift"])
    out.update(n_cells=int(mesh.n_cells), n_dofs=int(mesh.n_dofs), newton_residuals=history, level=level)
    return out


if __name__ == "__main__":
    import sys
    r = run(int(sys.argv[1]) if len(sys.argv) > 1 else 2, log=print)
    for k in ("drag", "lift", "p_diff", "drag_consistent", "lift_consistent"):
        lit = LITERATURE[k.split("_consistent")[0]]
        print(f"{k:16s} {r[k]: .8f}   literature {lit: .8f}   rel. dev. {r[k] / lit - 1: .2e}")
    print(r["n_cells"], "cells,", r["n_dofs"], "dofs,", len(r["newton_residuals"]) - 1, "Newton steps")


# ---------------------------------------------------------------------------------------------------------
# DFG benchmark 2D-2 (periodic vortex shedding at Re = 100): the time-dependent terms
# ---------------------------------------------------------------------------------------------------------
LITERATURE_2D2 = {"drag_max": (3.22, 3.24), "lift_max": (0.99, 1.01), "strouhal": (0.295, 0.305)}


class UnsteadyTurek:
    """input/input_turek_2D_Re100.json on this file's mesh: u_max = 1.5 (mean velocity 1, Re = 100), BDF2
    (include/time_integration.cc:61-91 through gls_oracle.OracleBDF), time-derivative terms in the Galerkin and in
    the stabilisation part, q-point-wise delta with the 1/dt^2 term, inflow ramp over t_init = 0.01
    (simulation.cc:46-49).  The time loop is main.cc:908-990; the nonlinear solve is Newton on the residual branch
    with a matrix assembled from the Newton branch that is kept for a few iterations / steps (the converged
    iterate does not depend on how fresh the matrix is; solver_nl.cc:36-89 with newton inexact)."""

    def __init__(self, level=2, dt=1.0 / 300.0, u_max=1.5, t_init=0.01, nu=NU, refresh=4):
        self.mesh = TurekMesh(level)
        m = self.mesh
        self.dt, self.u_max, self.t_init, self.nu, self.refresh = dt, u_max, t_init, nu, refresh
        self.op = go.OracleOperator(dim=2, degree=2, cell_dofs=m.cell_dofs, n_dofs=m.n_dofs,
                                    cell_points=m.cell_points, mapping_degree=2, constraints=m.constraints, nu=nu,
                                    c1=2.0, c2=1.0, theta=1.0, order=2, consider_time_derivative=True,
                                    increment_form=True, cell_wise_stabilization=False, path="sumfac")
        self.bdf = go.OracleBDF(2)
        self.history = [np.zeros(m.n_dofs) for _ in range(3)]
        self.t, self.n_steps, self._lu, self._age = 0.0, 0, None, 0
        self.profile = {d: v / U_MAX for d, v in m.inhomogeneities.items()}   # parabola with peak 1
        self.records = []

    def _distribute(self, x, t):
        f = self.u_max * min(t / self.t_init, 1.0) if self.t_init > 0 else self.u_max
        for d, v in self.profile.items():
            x[d] = f * v

    def step(self):
        op = self.op
        self.bdf.update_dt(self.dt)
        w = self.bdf.weights
        self.history[2], self.history[1] = self.history[1], self.history[0].copy()
        x = self.history[0]
        self._distribute(x, self.t)        # main.cc:926-942: boundary values at the old time level t
        op.set_previous_solution(self.history, w)
        if self._lu is not None and (self._w0 != w[0]):
            self._lu = None
        n_it = 0
        while True:
            op.set_linearization_point(x, self.dt)
            rhs = op.evaluate_residual(x, w[0])
            res = float(np.linalg.norm(rhs))
            if res < 1e-7:
                break
            if self._lu is None or self._age >= self.refresh or n_it >= 6:
                self._lu, self._age, self._w0 = spla.splu(system_matrix(op, w[0])), 0, w[0]
            inc = self._lu.solve(rhs)
            inc[op.constrained] = 0.0
            x += inc
            n_it += 1
            if n_it > 40:
                raise RuntimeError(f"Newton iteration did not converge at t = {self.t}: {res}")
        self._age += 1
        self.t += self.dt
        self.n_steps += 1
        f = drag_lift_pressure(self.mesh, x, self.nu)
        u_bar = self.u_max * 2.0 / 3.0
        rescale = (U_MAX * 2.0 / 3.0) ** 2 / u_bar ** 2      # drag_lift_pressure scales with the 2D-1 mean velocity
        rec = {"t": self.t, "drag": float(f["drag"] * rescale), "lift": float(f["lift"] * rescale),
               "p_diff": float(f["p_diff"]), "newton_iterations": n_it}
        self.records.append(rec)
        return rec


def shedding_statistics(records, t_from):
    """maxima of drag and lift and the Strouhal number (D / (u_mean T), u_mean = 1) from the upward zero crossings
    of the lift's oscillating part after t_from"""
    t = np.array([r["t"] for r in records])
    cl = np.array([r["lift"] for r in records])
    cd = np.array([r["drag"] for r in records])
    sel = t >= t_from
    t, cl, cd = t[sel], cl[sel], cd[sel]
    osc = cl - 0.5 * (cl.max() + cl.min())
    up = np.nonzero((osc[:-1] < 0) & (osc[1:] >= 0))[0]
    tc = t[up] - osc[up] * (t[up + 1] - t[up]) / (osc[up + 1] - osc[up])
    period = float(np.mean(np.diff(tc))) if len(tc) > 1 else float("nan")
    return {"drag_max": float(cd.max()), "drag_min": float(cd.min()), "lift_max": float(cl.max()),
            "lift_min": float(cl.min()), "period": period, "strouhal": 2.0 * R_CYL / period,
            "n_periods": int(len(tc) - 1)}
