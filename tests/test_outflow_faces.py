"""Boundary-face outflow terms (do_vmult_boundary, operator_ns.cc:1195-1301; face tables :423-521):
known-answer tests of the oracle restatement (CPU) and parity of the CUDA path with it (GPU)."""
import numpy as np
import pytest

from dealii_ns_gls_b200 import mesh as M
from tests.util import TI, make_gpu, make_oracle, rel_l2


def _mesh(dim, degree, curved, kinds):
    shape = (3, 2, 2)[:dim]
    deform = None
    if curved:
        def deform(x):
            y = x.copy()
            y[..., 0] += 0.08 * np.sin(2.0 * x[..., 1]) * (1 + x[..., 0])
            y[..., 1] += 0.05 * x[..., 0] * x[..., 0]
            return y
    eps = 1e-12
    m = M.structured_mesh(dim, shape, degree, deform=deform, mapping_degree=degree,
                          dirichlet=lambda x, c: (np.abs(x[:, 0]) < eps) & (c < dim))
    m.outflow_faces = M.boundary_faces(m, kinds, target_velocity=lambda x: 0.3 + 0.5 * x ** 2)
    return m


def _oracle(m, ti, **kw):
    o = make_oracle(m, ti, path="naive", **kw)
    f = m.outflow_faces
    o.set_outflow_faces(f["face_cell"], f["face_no"], f["face_kind"], f["target"])
    return o


def test_face_geometry_two_ways():
    """mesh.boundary_faces (what the adapter takes from MatrixFree) against the oracle's own face geometry"""
    m = _mesh(3, 2, True, {1: 1, 3: 2})
    o = _oracle(m, TI(1, [10.0, -10.0], 0.1))
    f = m.outflow_faces
    assert np.allclose(o.faces["normal"], f["normal"], atol=1e-13)
    assert np.allclose(o.faces["jxw"], f["jxw"], atol=1e-13)
    assert np.allclose(o.faces["Jinv"], f["inv_jac"], atol=1e-12)
    assert np.allclose(np.linalg.norm(f["normal"], axis=2), 1.0)


def test_nitsche_is_symmetric_and_cut_is_mass_like():
    ti = TI(1, [10.0, -10.0], 0.1)
    m = _mesh(2, 2, False, {1: 2})
    o = _oracle(m, ti)
    rng = np.random.default_rng(0)
    o.set_linearization_point(rng.standard_normal(m.n_dofs), 0.1)
    A = o.dense_matrix(10.0)
    o.faces = None
    F = A - o.dense_matrix(10.0)
    assert np.abs(F).max() > 1 and np.abs(F - F.T).max() < 1e-12
    comp = M.dof_components(m)
    x = (comp == 0).astype(float)
    mcells = m.outflow_faces["face_cell"]
    h = np.sqrt(4.0 * m.cell_measure[mcells[0]] / np.pi) / 2
    assert abs(x @ F @ x - 1.0 / h ** 3) < 1e-9  # beta * |face|, |face| = 1, gradient of a constant = 0
    # "cut" faces: (v, beta min(0, U.n) u) vanishes for outflow (U.n > 0) and is -beta |U.n| mass for inflow
    m = _mesh(2, 1, False, {1: 1})
    o = _oracle(m, ti)
    U = np.zeros(m.n_dofs)
    U[M.dof_components(m) == 0] = 2.0
    o.set_linearization_point(U, 0.1)
    A = o.dense_matrix(10.0)
    o.faces = None
    assert np.abs(A - o.dense_matrix(10.0)).max() < 1e-13
    o = _oracle(m, ti)
    o.set_linearization_point(-U, 0.1)
    A = o.dense_matrix(10.0)
    o.faces = None
    F = A - o.dense_matrix(10.0)
    x = (M.dof_components(m) == 1).astype(float)
    h = np.sqrt(4.0 * m.cell_measure[0] / np.pi)
    assert abs(x @ F @ x - (-2.0 / h ** 2)) < 1e-10


CASES = [(2, 1, False, {1: 1}), (2, 2, True, {1: 2, 3: 1}), (3, 2, False, {1: 1, 5: 2}), (3, 1, True, {1: 2}),
         (3, 3, True, {1: 1, 2: 2}), (2, 4, False, {1: 2})]


@pytest.mark.gpu
@pytest.mark.parametrize("number,tol", [("double", 1e-12), ("float", 3e-5)])
@pytest.mark.parametrize("dim,degree,curved,kinds", CASES)
def test_gpu_outflow_faces_match_oracle(dim, degree, curved, kinds, number, tol):
    import torch
    ti = TI(1, [10.0, -10.0], 0.1)
    m = _mesh(dim, degree, curved, kinds)
    kw = dict(nu=0.05, ctd=True, cell_wise=False)
    o = _oracle(m, ti, **kw)
    g = make_gpu(m, ti, number=number, **kw)
    dt = torch.float64 if number == "double" else torch.float32
    rng = np.random.default_rng(7)
    U, x, old = (rng.uniform(-1, 1, m.n_dofs) for _ in range(3))

    def dev(a):
        return torch.from_numpy(a).to("cuda", dtype=dt)

    o.set_previous_solution([old, old], ti.get_weights())
    o.set_linearization_point(U, 0.1)
    g.set_previous_solution([dev(old), dev(old)])
    g.set_linearization_point(dev(U))
    # the face part alone must be visible: compare against the oracle with and without faces
    y = torch.zeros(m.n_dofs, dtype=dt, device="cuda")
    g.vmult(y, dev(x))
    ref = o.vmult(x, 10.0)
    assert rel_l2(y.cpu().numpy(), ref, mesh=m) < tol
    faces, o.faces = o.faces, None
    assert rel_l2(o.vmult(x, 10.0), ref) > 1e-3
    o.faces = faces
    # residual (plain reads, u - u_target on Nitsche faces, u itself as the transport velocity on cut faces)
    r = torch.zeros_like(y)
    g.evaluate_residual(r, dev(x))
    xb = x.copy()
    assert rel_l2(r.cpu().numpy(), o.evaluate_residual(xb, 10.0), mesh=m) < tol
    # inverse diagonal
    d = torch.zeros_like(y)
    g.compute_inverse_diagonal(d)
    assert rel_l2(d.cpu().numpy(), o.compute_inverse_diagonal(10.0)) < max(tol, 1e-11)
    # system matrix column by column (coarse-level path)
    if m.n_dofs < 400:
        A = g.get_system_matrix().cpu().numpy()
        assert rel_l2(A, o.dense_matrix(10.0)) < tol
