"""Known-answer checks that pin the CPU oracle (parity is otherwise unpinned: the
reference has no tests/golden vectors)."""
import numpy as np
import pytest

from oracle import gls_oracle as go
from dealii_ns_gls_b200 import mesh as gm


def make_op(mesh, **kw):
    args = dict(dim=mesh.dim, degree=mesh.degree, cell_dofs=mesh.cell_dofs, n_dofs=mesh.n_dofs,
                cell_points=mesh.cell_points, mapping_degree=mesh.mapping_degree,
                constraints=mesh.constraints, nu=0.1, c1=4.0, c2=2.0, theta=1.0, order=2,
                consider_time_derivative=False, increment_form=True,
                cell_wise_stabilization=True)
    args.update(kw)
    return go.OracleOperator(**args)


def test_basis_known_values():
    assert np.allclose(go.gauss_lobatto_points(2), [0, 0.5, 1])
    s5 = 1 / np.sqrt(5.0)
    assert np.allclose(go.gauss_lobatto_points(3), [0, 0.5 - 0.5 * s5, 0.5 + 0.5 * s5, 1], atol=1e-15)
    s37 = np.sqrt(3.0 / 7.0)
    assert np.allclose(go.gauss_lobatto_points(4), [0, 0.5 - 0.5 * s37, 0.5, 0.5 + 0.5 * s37, 1], atol=1e-15)
    x, w = go.gauss_points_weights(3)
    assert np.allclose(x, [0.5 - 0.5 * np.sqrt(0.6), 0.5, 0.5 + 0.5 * np.sqrt(0.6)], atol=1e-15)
    assert np.allclose(w, [5 / 18, 8 / 18, 5 / 18], atol=1e-15)
    for p in range(1, 5):
        b = go.Basis1D(p)
        assert np.allclose(b.S.sum(axis=1), 1, atol=1e-14)      # partition of unity
        assert np.allclose(b.G.sum(axis=1), 0, atol=1e-12)
        # collocation derivative differentiates the interpolant: G = D S
        assert np.allclose(b.D @ b.S, b.G, atol=1e-12)
        # exact for monomials up to degree p
        for k in range(p + 1):
            assert np.allclose(b.S @ b.nodes ** k, b.xq ** k, atol=1e-14)
            assert np.allclose(b.G @ b.nodes ** k, k * b.xq ** max(k - 1, 0) if k else 0, atol=1e-12)


def test_q1_closed_form_element_matrices():
    """U = 0, nu = 0, c2 = 0: velocity block = w*M, pressure block = delta1*K (Q1, 2-D square)."""
    m = gm.hypercube(2, 1, 1, order="lex")
    h = 1.0
    op = make_op(m, nu=0.0, c1=4.0, c2=0.0)
    op.set_linearization_point(np.zeros(m.n_dofs), dt=0.1)
    w = 10.0
    A = op.cell_matrices(w)[0]
    M = h * h / 36.0 * np.array([[4, 2, 2, 1], [2, 4, 1, 2], [2, 1, 4, 2], [1, 2, 2, 4]])
    K = 1.0 / 6.0 * np.array([[4, -1, -1, -2], [-1, 4, -2, -1], [-1, -2, 4, -1], [-2, -1, -1, 4]])
    # nu(=0) < h: delta1 = c1/sqrt(1/dt^2 + 0) = c1*dt
    d1 = 4.0 * 0.1
    assert np.allclose(A[0:4, 0:4], w * M, atol=1e-13)
    assert np.allclose(A[4:8, 4:8], w * M, atol=1e-13)
    assert np.allclose(A[0:4, 4:8], 0, atol=1e-13)
    assert np.allclose(A[8:12, 8:12], d1 * K, atol=1e-13)
    # (q, div u) block vs -(div v, p) block: B and -B^T
    assert np.allclose(A[8:12, 0:8], -A[0:8, 8:12].T, atol=1e-13)
    # divergence block, closed form: int phi_i d_x phi_j
    Bx = np.array([[-2, 2, -1, 1], [-2, 2, -1, 1], [-1, 1, -2, 2], [-1, 1, -2, 2]]) * h / 12.0
    assert np.allclose(A[8:12, 0:4], Bx, atol=1e-13)


def test_viscous_block_is_symmetric_gradient():
    """With U=0, w=0, c1=c2=0 the velocity block is (eps(v), 2 nu eps(u)); check a
    rigid rotation lies in its kernel and a shear gives the analytic energy."""
    m = gm.hypercube(3, 2, 2)
    op = make_op(m, nu=0.3, c1=0.0, c2=0.0)
    op.set_linearization_point(np.zeros(m.n_dofs), dt=0.1)
    # node coordinates per dof
    tb = op.tb
    gl = go.gauss_lobatto_points(2)
    X = np.zeros((m.n_dofs, 3))
    comp = np.zeros(m.n_dofs, dtype=int)
    n = 3
    for k in range(m.n_cells):
        c0 = m.cell_points[k, 0]
        hh = m.cell_points[k, -1] - c0
        for c in range(4):
            for l in range(27):
                i, j, kk = l % n, (l // n) % n, l // (n * n)
                d = m.cell_dofs[k, c * 27 + l]
                X[d] = c0 + hh * np.array([gl[i], gl[j], gl[kk]])
                comp[d] = c
    rot = np.zeros(m.n_dofs)
    rot[comp == 0] = -X[comp == 0, 1]
    rot[comp == 1] = X[comp == 1, 0]
    assert np.abs(op.vmult(rot, 0.0)).max() < 1e-13
    shear = np.zeros(m.n_dofs)
    shear[comp == 0] = X[comp == 0, 1]  # u = (y,0,0): eps_xy = 1/2, 2 nu eps:eps = nu
    assert np.isclose(shear @ op.vmult(shear, 0.0), 0.3, atol=1e-13)


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 3), (3, 2), (3, 4)])
def test_naive_vs_sumfac(dim, degree):
    m = gm.cylinder_shell((3, 4) if dim == 2 else (2, 4, 2), degree)
    rng = np.random.default_rng(1)
    outs = []
    for path in ("naive", "sumfac"):
        op = make_op(m, path=path, consider_time_derivative=True, cell_wise_stabilization=False)
        hist = [rng.standard_normal(m.n_dofs) for _ in range(3)] if path == "naive" else hist
        op.set_previous_solution(hist, [15.0, -20.0, 5.0])
        lin = rng.standard_normal(m.n_dofs) if path == "naive" else lin
        op.set_linearization_point(lin, 0.1)
        src = rng.standard_normal(m.n_dofs) if path == "naive" else src
        outs.append((op.vmult(src, 15.0), op.evaluate_residual(src, 15.0)))
    for a, b in zip(outs[0], outs[1]):
        assert np.linalg.norm(a - b) <= 1e-13 * np.linalg.norm(a)


def test_polynomial_exactness_of_evaluation():
    m = gm.hypercube(3, 2, 2)
    op = make_op(m)
    # u_c = polynomial of degree 2 per direction, exact on Q2
    def f(x, c):
        return (1 + c) * x[..., 0] ** 2 * x[..., 1] - x[..., 2] ** 2 + 0.5 * x[..., 0] * x[..., 2]
    def df(x, c):
        return np.stack([(1 + c) * 2 * x[..., 0] * x[..., 1] + 0.5 * x[..., 2],
                         (1 + c) * x[..., 0] ** 2,
                         -2 * x[..., 2] + 0.5 * x[..., 0]], axis=-1)
    gl = go.gauss_lobatto_points(2)
    n = 3
    vec = np.zeros(m.n_dofs)
    for k in range(m.n_cells):
        c0 = m.cell_points[k, 0]
        hh = m.cell_points[k, -1] - c0
        for c in range(4):
            for l in range(27):
                i, j, kk = l % n, (l // n) % n, l // (n * n)
                vec[m.cell_dofs[k, c * 27 + l]] = f(c0 + hh * np.array([gl[i], gl[j], gl[kk]]), c)
    val, grad = op._evaluate(op._gather(vec))
    for k in range(m.n_cells):
        c0 = m.cell_points[k, 0]
        hh = m.cell_points[k, -1] - c0
        xq = c0 + hh * op.tb.xq
        for c in range(4):
            assert np.allclose(val[k, c], f(xq, c), atol=1e-13)
            assert np.allclose(grad[k, c].T, df(xq, c), atol=1e-12)


@pytest.mark.parametrize("ctd", [False, True])
def test_newton_branch_is_derivative_of_residual(ctd):
    """operator_ns.cc:919-948: the linearized system is the Gateaux derivative of the
    fixed-point system at S = B = U with delta frozen (nu >= h => delta = c h^2)."""
    m = gm.cylinder_shell((2, 5), 2, r_inner=0.5, r_outer=1.5)
    rng = np.random.default_rng(7)
    kw = dict(nu=5.0, consider_time_derivative=ctd, cell_wise_stabilization=True, order=2)
    hist = [rng.standard_normal(m.n_dofs) for _ in range(3)]
    wts = [15.0, -20.0, 5.0]
    u = rng.standard_normal(m.n_dofs)
    v = rng.standard_normal(m.n_dofs)
    free = np.ones(m.n_dofs, dtype=bool)
    free[list(m.constraints.keys())] = False
    v[~free] = 0
    u[~free] = 0

    def F(x):
        op = make_op(m, **kw)
        op.set_previous_solution(hist, wts)
        op.set_linearization_point(x, 0.1)
        return -op.evaluate_residual(x, wts[0])

    eps = 1e-4
    fd = (F(u + eps * v) - F(u - eps * v)) / (2 * eps)
    op = make_op(m, **kw)
    op.set_previous_solution(hist, wts)
    op.set_linearization_point(u, 0.1)
    jv = op.vmult(v, wts[0])
    assert np.linalg.norm(fd[free] - jv[free]) <= 1e-6 * np.linalg.norm(jv[free])


def test_diagonal_and_constraints_against_dense_matrix():
    m = gm.hypercube(2, 3, 2)
    gm.add_random_constraints(m, n_weighted=6, n_zero=5, seed=3)
    rng = np.random.default_rng(2)
    op = make_op(m, nu=0.05)
    op.set_linearization_point(rng.standard_normal(m.n_dofs), 0.1)
    A = op.dense_matrix(10.0)
    x = rng.standard_normal(m.n_dofs)
    assert np.allclose(A @ x, op.vmult(x, 10.0), atol=1e-11)
    cons = np.array(sorted(m.constraints.keys()))
    # identity on constrained rows/cols
    assert np.allclose(A[cons][:, cons], np.eye(len(cons)))
    d = np.diag(A).copy()
    inv = np.where(np.abs(d) > 1e-10, 1 / d, 1.0)
    assert np.allclose(op.compute_inverse_diagonal(10.0), inv, rtol=1e-11, atol=1e-13)


def test_penalty_parameters_closed_form():
    m = gm.hypercube(3, 4, 2)
    op = make_op(m, nu=0.1)
    vec = np.zeros(m.n_dofs)
    op.set_linearization_point(vec, 0.1)
    h = 0.25
    # nu < h: delta1 = c1/sqrt(1/dt^2) = 0.4, delta2 = c2 h
    assert np.allclose(op.delta1_cell, 0.4) and np.allclose(op.delta2_cell, 0.5)
    op2 = make_op(gm.hypercube(3, 16, 1), nu=0.1)
    op2.set_linearization_point(np.zeros(op2.n_dofs), 0.1)
    hh = 1 / 16
    assert np.allclose(op2.delta1_cell, 4 * hh * hh) and np.allclose(op2.delta2_cell, 2 * hh * hh)
    # q-wise (Lethe): h = (6V/pi)^(1/3)/p
    hq = (6 * h ** 3 / np.pi) ** (1 / 3) / 2
    d1 = 1 / np.sqrt(100 + 4 * 1e-12 / hq ** 2 + 9 * (4 * 0.1 / hq ** 2) ** 2)
    assert np.allclose(op.delta1_q, d1, rtol=1e-13)
    assert np.allclose(op.delta2_q, 0.5 * hq * 1e-6, rtol=1e-13)


def test_bdf_weights():
    b = go.OracleBDF(2)
    b.update_dt(0.1)
    assert np.allclose(b.weights, [10, -10, 0])      # performance.cc:44-46
    b.update_dt(0.1)
    assert np.allclose(b.weights, [15, -20, 5])
    b3 = go.OracleBDF(3)
    for _ in range(3):
        b3.update_dt(0.2)
    assert np.allclose(b3.weights, np.array([11 / 6, -3, 1.5, -1 / 3]) / 0.2)


def test_rank_count_invariance():
    """Partitioned meshes (owner = lowest rank, ghosts appended) reproduce the 1-rank vmult."""
    rng = np.random.default_rng(5)
    full = gm.cylinder_shell((2, 4, 2), 2)
    ng = full.n_global_dofs
    lin_g = rng.standard_normal(ng)
    src_g = rng.standard_normal(ng)
    op = make_op(full, cell_wise_stabilization=False)
    op.set_linearization_point(lin_g, 0.1)
    ref = op.vmult(src_g, 10.0)
    for R in (2, 3):
        acc = np.zeros(ng)
        for r in range(R):
            m = gm.cylinder_shell((2, 4, 2), 2, n_ranks=R, rank=r)
            part = m.partition
            l2g = np.concatenate([part.owned_offset + np.arange(m.n_owned), part.ghost_global])
            # constraints were generated on local indices; the oracle sees a local problem
            o = make_op(m, cell_wise_stabilization=False)
            o.set_linearization_point(lin_g[l2g], 0.1)
            loc_src = src_g[l2g].copy()
            out = o._distribute_transpose(o._scatter(o._apply_cells(o._gather(o._resolve(loc_src)), 10.0, False)))
            np.add.at(acc, l2g, out)                      # compress(add)
        cons = np.array(sorted(full.constraints.keys()))
        acc[cons] = src_g[cons]
        assert np.linalg.norm(acc - ref) <= 1e-13 * np.linalg.norm(ref)


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 4), (3, 2), (3, 3)])
@pytest.mark.parametrize("branch", [0, 1, 2])
def test_c_restatement_matches_numpy_oracle(dim, degree, branch):
    """Third, independently written path: C, SIMD over cell batches, OpenMP, atomics on shared dofs."""
    from oracle.gls_oracle_c import COracle
    m = gm.cylinder_shell((3, 5) if dim == 2 else (2, 5, 3), degree, no_slip=False)
    rng = np.random.default_rng(3)
    theta = 0.5 if branch else 1.0
    op = make_op(m, theta=theta, order=1 if branch else 2, consider_time_derivative=(branch == 0),
                 increment_form=(branch == 0), cell_wise_stabilization=(branch == 1), path="naive")
    hist = [rng.standard_normal(m.n_dofs) for _ in range(3)]
    op.set_previous_solution(hist, [15.0, -20.0, 5.0] if branch == 0 else [10.0, -10.0])
    op.set_linearization_point(rng.standard_normal(m.n_dofs), 0.1)
    src = rng.standard_normal(m.n_dofs)
    w = 15.0 if branch == 0 else 10.0
    ref = op._scatter(op._apply_cells(op._gather(src), w, residual=(branch == 2)))
    co = COracle.from_numpy_oracle(op, branch)
    for nt in (1, 3):
        got = co.apply(src, w, n_threads=nt)
        assert np.linalg.norm(got - ref) <= 1e-13 * np.linalg.norm(ref)


def test_relaxation_restatement_smooths():
    """oracle/gls_smoother.py (PreconditionRelaxation as multigrid.cc:290-304 configures it): omega from
    the power iteration is positive and below 2 / lambda_max, vmult equals n steps from zero, and the
    sweeps damp the high-frequency start vector of deal.II's eigenvalue estimate."""
    from dealii_ns_gls_b200 import mesh as gm
    from oracle.gls_smoother import OracleRelaxation
    from tests.util import TI, make_oracle
    mesh = gm.hypercube(2, 4, 2)
    ti = TI(1, [10.0, -10.0], 0.1)
    ora = make_oracle(mesh, ti, nu=0.1)
    rng = np.random.default_rng(3)
    ora.set_linearization_point(0.1 * rng.uniform(-1, 1, mesh.n_dofs), 0.1)
    d = ora.compute_inverse_diagonal(10.0)
    sm = OracleRelaxation(ora, 10.0, d)
    omega = sm.get_relaxation()
    assert 0 < omega < 2.0 / (sm.max_eigenvalue_estimate / 1.2)
    b = rng.uniform(-1, 1, mesh.n_dofs)
    x = sm.vmult(b)
    assert np.allclose(x, sm.step(np.zeros_like(b), b), rtol=1e-12, atol=1e-14)
    e = ((np.arange(mesh.n_dofs)) % 11).astype(float)
    e -= e.mean()
    r0 = np.linalg.norm(e)
    r5 = np.linalg.norm(sm.step(e, np.zeros_like(e)))  # error propagation of 5 sweeps
    assert r5 < 0.9 * r0


@pytest.mark.parametrize("kind,order,cell_wise", [("channel3", 1, True), ("shell2", 2, False), ("channel2", 0, False)])
def test_c_backed_operator_matches_numpy_oracle(kind, order, cell_wise):
    """oracle/gls_fast.py (vmult, residual and unit-vector diagonal in C, Cartesian cells detected) against the
    numpy oracle it derives from, before and after a change of the linearization point"""
    from dealii_ns_gls_b200 import mesh as gm
    from dealii_ns_gls_b200.driver import ChannelParameters, channel_level_mesh
    from oracle.gls_fast import FastOracleOperator
    rng = np.random.default_rng(0)
    if kind == "channel3":
        m = channel_level_mesh(ChannelParameters(dim=3, fe_degree=2, n_global_refinements=0), 1)
    elif kind == "channel2":
        m = channel_level_mesh(ChannelParameters(dim=2, fe_degree=1, n_global_refinements=1), 2)
    else:
        m = gm.cylinder_shell((3, 8), 2)
    kw = dict(dim=m.dim, degree=m.degree, cell_dofs=m.cell_dofs, n_dofs=m.n_dofs, cell_points=m.cell_points,
              mapping_degree=m.mapping_degree, constraints=m.constraints, nu=0.01, c1=2.0, c2=1.0, theta=1.0, order=order,
              consider_time_derivative=True, increment_form=True, cell_wise_stabilization=cell_wise)
    a, b = go.OracleOperator(path="sumfac", **kw), FastOracleOperator(**kw)
    assert b._cartesian == (kind != "shell2")
    hist = [rng.uniform(-1, 1, m.n_dofs) for _ in range(order + 1)]
    w = [3.0, -4.0, 1.0][:order + 1] if order else [0.0]
    src = rng.uniform(-1, 1, m.n_dofs)

    def err(x, y):
        return np.linalg.norm(x - y) / np.linalg.norm(y)

    for _ in range(2):
        lin = rng.uniform(-1, 1, m.n_dofs)
        for o in (a, b):
            o.set_previous_solution(hist, w)
            o.set_linearization_point(lin, 0.1)
        assert err(b.vmult(src, w[0]), a.vmult(src, w[0])) < 1e-14
        assert err(b.evaluate_residual(src, w[0]), a.evaluate_residual(src, w[0])) < 1e-14
        assert err(b.compute_inverse_diagonal(w[0]), a.compute_inverse_diagonal(w[0])) < 1e-13
