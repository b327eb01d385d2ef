"""An analytic Navier-Stokes solution as known answer for the restated operator in 2-D AND 3-D, on straight and on
curved cells: Kovasznay flow (L. Kovasznay, Proc. Camb. Phil. Soc. 44, 1948), an exact stationary solution of the
incompressible Navier-Stokes equations WITHOUT body force,

    u = 1 - e^{lambda x} cos(2 pi y),  v = lambda / (2 pi) e^{lambda x} sin(2 pi y),  p = (1 - e^{2 lambda x}) / 2,
    lambda = Re / 2 - sqrt(Re^2 / 4 + 4 pi^2),   nu = 1 / Re,

carried into 3-D by a rotation of the coordinate system (the equations are rotation-invariant, so
u'(x) = R u(R^T x), p'(x) = p(R^T x) is again exact, now with all three components and all nine derivatives
non-zero).  TEST INFRASTRUCTURE ONLY (see oracle/gls_oracle.py).

The stationary problem of the reference's configuration (increment form, "time intration": "none", GLS terms on,
cell-wise or q-point-wise delta) is solved on a block with the exact velocity on the whole boundary and the exact
pressure in one node (the constant the velocity-Dirichlet problem leaves open), by Newton's method on the residual
branch with the matrix assembled from the Newton branch (gls_turek.solve machinery).  Since the stabilisation is
residual-based, the discrete solution has to converge to the analytic one at the order of the element: the nodal
velocity error of Q2 falls by ~ 8 per halving of h, the pressure error by >= 4 -- in 2-D, in 3-D, and on
smoothly deformed (curved, "general geometry") cells.  A wrong term, sign, factor, Jacobian or constraint row
destroys the order or the convergence; tests/test_exact_solution.py holds the thresholds.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse.linalg as spla

from dealii_ns_gls_b200 import mesh as gm

from . import gls_oracle as go
from .gls_turek import system_matrix

RE = 40.0
LAM = RE / 2.0 - math.sqrt(RE * RE / 4.0 + 4.0 * math.pi ** 2)


def rotation(dim):
    """a fixed rotation with no axis left in place (3-D); the identity in 2-D"""
    if dim == 2:
        return np.eye(2)
    a, b, c = 0.4, -0.7, 0.3
    Rx = np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]])
    Ry = np.array([[math.cos(b), 0, math.sin(b)], [0, 1, 0], [-math.sin(b), 0, math.cos(b)]])
    Rz = np.array([[math.cos(c), -math.sin(c), 0], [math.sin(c), math.cos(c), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def exact(x):
    """(velocity [n, dim], pressure [n]) of the (rotated) Kovasznay flow at the points x [n, dim]"""
    dim = x.shape[1]
    R = rotation(dim)
    y = x @ R                       # coordinates in the frame of the 2-D solution: y = R^T x
    e = np.exp(LAM * y[:, 0])
    u2 = np.zeros_like(y)
    u2[:, 0] = 1.0 - e * np.cos(2.0 * math.pi * y[:, 1])
    u2[:, 1] = LAM / (2.0 * math.pi) * e * np.sin(2.0 * math.pi * y[:, 1])
    return u2 @ R.T, 0.5 * (1.0 - e * e)


def make_mesh(dim, n, degree=2, curved=False):
    """block [-0.5, 0.5]^dim of n^dim cells, optionally deformed smoothly in its interior (curved cells, Q2 mapping);
    velocity rows of the whole boundary and one pressure row constrained"""
    eps = 1e-12

    def boundary(ref, c):
        on = np.zeros(len(ref), dtype=bool)
        for e in range(dim):
            on |= (np.abs(ref[:, e] + 0.5) < eps) | (np.abs(ref[:, e] - 0.5) < eps)
        if c == dim:
            return np.all(np.abs(ref + 0.5) < eps, axis=1)      # the pressure node in the corner (-0.5, ..)
        return on

    def deform(x):
        out = x.copy()
        bump = np.prod(np.sin(math.pi * (x + 0.5)), axis=-1)     # vanishes on the boundary
        for e in range(dim):
            out[..., e] += 0.06 * bump * math.cos(1.0 + e)
        return out

    return gm.structured_mesh(dim, (n,) * dim, degree, extent=np.ones(dim), origin=-0.5 * np.ones(dim),
                              deform=deform if curved else None, mapping_degree=degree, dirichlet=boundary)


def node_coordinates(mesh):
    """physical coordinates of the support point of every dof (Q_p nodes = mapping support points for
    mapping_degree = degree), and its component"""
    n_loc, C = mesh.n_loc, mesh.dim + 1
    xyz = np.zeros((mesh.n_dofs, mesh.dim))
    comp = np.zeros(mesh.n_dofs, dtype=np.int64)
    if mesh.cell_points.shape[1] == n_loc:
        pts = mesh.cell_points
    else:   # Cartesian block generated with the 2^dim vertices only
        ref = gm.dof_coordinates(mesh)
        for c in range(C):
            comp[mesh.cell_dofs[:, c * n_loc:(c + 1) * n_loc].astype(np.int64).reshape(-1)] = c
        return ref, comp
    for c in range(C):
        idx = mesh.cell_dofs[:, c * n_loc:(c + 1) * n_loc].astype(np.int64).reshape(-1)
        xyz[idx] = pts.reshape(-1, mesh.dim)
        comp[idx] = c
    return xyz, comp


def solve(dim, n, *, curved=False, cell_wise=False, tol=1e-11, log=None):
    """returns dict(err_u, err_p, h, n_dofs, newton_residuals): maximal nodal errors against the analytic solution"""
    mesh = make_mesh(dim, n, curved=curved)
    xyz, comp = node_coordinates(mesh)
    u_ex, p_ex = exact(xyz)
    x_exact = np.where(comp == dim, p_ex, np.take_along_axis(u_ex, np.minimum(comp, dim - 1)[:, None], axis=1)[:, 0])
    op = go.OracleOperator(dim=dim, degree=2, cell_dofs=mesh.cell_dofs, n_dofs=mesh.n_dofs,
                           cell_points=mesh.cell_points, mapping_degree=mesh.mapping_degree,
                           constraints=mesh.constraints, nu=1.0 / RE, c1=1.0, c2=1.0, theta=1.0, order=0,
                           consider_time_derivative=True, increment_form=True, cell_wise_stabilization=cell_wise,
                           path="sumfac")
    cons = op.constrained
    x = np.zeros(mesh.n_dofs)
    x[cons] = x_exact[cons]                       # constraints_inhomogeneous.distribute
    history = []
    for it in range(30):
        op.set_linearization_point(x, 1.0)
        rhs = op.evaluate_residual(x, 0.0)
        history.append(float(np.linalg.norm(rhs)))
        if log is not None:
            log(f"    [N] step {it} ; residual = {history[-1]:.6e}")
        if history[-1] < tol:
            break
        inc = spla.spsolve(system_matrix(op), rhs)
        inc[cons] = 0.0
        x = x + inc
    else:
        raise RuntimeError(f"Newton iteration did not converge: {history}")
    err = np.abs(x - x_exact)
    return {"err_u": float(err[comp < dim].max()), "err_p": float(err[comp == dim].max()), "h": 1.0 / n,
            "n_dofs": int(mesh.n_dofs), "newton_residuals": history}


if __name__ == "__main__":
    import sys
    dim = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    curved = len(sys.argv) > 2 and sys.argv[2] == "curved"
    prev = None
    for n in ((8, 16, 32) if dim == 2 else (3, 6)):
        r = solve(dim, n, curved=curved, log=print)
        rate = "" if prev is None else f"   ratios u {prev['err_u'] / r['err_u']:.2f}  p {prev['err_p'] / r['err_p']:.2f}"
        print(f"dim {dim} n {n:3d} dofs {r['n_dofs']:7d}  err_u {r['err_u']:.3e}  err_p {r['err_p']:.3e}{rate}")
        prev = r
