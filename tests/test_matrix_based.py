"""Independent anchor for the oracle: the FEValues-style assembly of
NavierStokesOperatorMatrixBased (operator_ns.cc:1600-1756), restated in oracle/gls_matrix_based.py without
sharing code with oracle/gls_oracle.py, must equal the matrix-free fixed-point operator:

    diag(tau I_u, I_p) A_mf == A_mb      and      -diag(tau I_u, I_p) residual_mf(u) == A_mb u - rhs_mb

(theta scheme, weight 1/tau, no time derivative in the stabilization, cell-wise delta with the previous
solution as the linearization point so that both sides see the same u_max)."""
import numpy as np
import pytest

from dealii_ns_gls_b200 import mesh as gm
from oracle.gls_matrix_based import MatrixBasedOperator, gauss_unit, lobatto_unit
from oracle import gls_oracle as go
from tests.util import TI, make_oracle

TAU = 0.1


def _mesh(kind, dim, degree):
    if kind == "cube":
        return gm.hypercube(dim, 3 if dim == 2 else 2, degree)
    if kind == "cube_hanging":
        # hanging-node-like rows: masters of the SAME component (diag(tau I_u, I_p) must commute with C)
        m = gm.hypercube(dim, 3 if dim == 2 else 2, degree)
        comp, rng = _components(m), np.random.default_rng(3)
        chosen = rng.choice(m.n_dofs, size=10, replace=False)
        for t, dof in enumerate(chosen):
            same = np.array([i for i in np.nonzero(comp == comp[dof])[0] if i not in set(chosen.tolist())])
            masters = rng.choice(same, size=int(rng.integers(2, 5)), replace=False)
            m.constraints[int(dof)] = [] if t >= 6 else [(int(a), float(w)) for a, w in
                                                         zip(masters, rng.uniform(-0.5, 1.0, len(masters)))]
        return m
    if kind == "hanging":  # a true 2:1 mesh with deal.II's hanging-node rows (mesh.hypercube_hanging)
        return gm.hypercube_hanging(dim, 2, degree)
    return gm.cylinder_shell((2, 5) if dim == 2 else (1, 4, 1), degree)


def _components(mesh):
    comp = np.zeros(mesh.n_dofs, dtype=np.int64)
    nl = (mesh.degree + 1) ** mesh.dim
    for c in range(mesh.dim + 1):
        comp[mesh.cell_dofs[:, c * nl:(c + 1) * nl].reshape(-1)] = c
    return comp


def _constraint_matrix(mesh):
    C = np.eye(mesh.n_dofs)
    for dof, row in mesh.constraints.items():
        C[dof, dof] = 0.0
        for m, w in row:
            C[dof, m] += w
    return C


def test_points_agree_with_the_oracles():
    for p in range(1, 5):
        assert np.max(np.abs(lobatto_unit(p) - go.gauss_lobatto_points(p))) < 1e-15
        x, w = gauss_unit(p + 1)
        xo, wo = go.gauss_points_weights(p + 1)
        assert np.max(np.abs(x - xo)) < 1e-15 and np.max(np.abs(w - wo)) < 1e-15


@pytest.mark.parametrize("theta", [1.0, 0.5])
@pytest.mark.parametrize("kind,dim,degree", [("cube", 2, 1), ("shell", 2, 2), ("cube_hanging", 2, 2), ("cube", 3, 1),
                                             ("shell", 3, 2), ("cube", 2, 3), ("hanging", 2, 2), ("hanging", 3, 1)])
def test_matrix_free_equals_matrix_based(kind, dim, degree, theta):
    mesh = _mesh(kind, dim, degree)
    rng = np.random.default_rng(42)
    u_0 = rng.uniform(-1, 1, mesh.n_dofs)
    nu, c1, c2 = 0.05, 4.0, 2.0
    ti = TI(1, [1.0 / TAU, -1.0 / TAU], TAU, theta=theta)
    mf = make_oracle(mesh, ti, nu=nu, c1=c1, c2=c2, ctd=False, increment_form=False, cell_wise=True)
    mf.set_previous_solution([u_0, u_0], ti.get_weights())
    mf.set_linearization_point(u_0, TAU)
    mb = MatrixBasedOperator(dim=dim, degree=degree, cell_dofs=mesh.cell_dofs, n_dofs=mesh.n_dofs,
                             cell_points=mesh.cell_points, mapping_degree=mesh.mapping_degree,
                             nu=nu, c1=c1, c2=c2, theta=theta)
    A_mb, rhs_mb = mb.assemble(u_0, u_0, TAU)
    C = _constraint_matrix(mesh)
    scale = np.where(_components(mesh) < dim, TAU, 1.0)
    free = np.array([i for i in range(mesh.n_dofs) if i not in mesh.constraints])
    A_mf = mf.dense_matrix(1.0 / TAU)
    lhs = (scale[:, None] * A_mf)[np.ix_(free, free)]
    ref = (C.T @ A_mb @ C)[np.ix_(free, free)]
    assert np.linalg.norm(lhs - ref) / np.linalg.norm(ref) < 1e-12
    # constrained rows of the matrix-free operator are the identity (operator_ns.cc:719-721)
    cons = np.array(sorted(mesh.constraints.keys()), dtype=np.int64)
    if len(cons):
        assert np.array_equal(A_mf[np.ix_(cons, cons)], np.eye(len(cons)))
    # residual branch (u_time_derivative_old, u_old_gradient, p_old_gradient tables)
    u = rng.uniform(-1, 1, mesh.n_dofs)
    u[cons] = 0.0
    u = C @ u  # hanging-node rows carry the interpolated value, like a distributed solution vector
    r_mf = mf.evaluate_residual(u, 1.0 / TAU)
    ref_r = (C.T @ (A_mb @ u - rhs_mb))[free]
    assert np.linalg.norm(-(scale * r_mf)[free] - ref_r) / np.linalg.norm(ref_r) < 1e-12
    assert np.all(r_mf[cons] == 0.0) if len(cons) else True
