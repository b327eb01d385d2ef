"""GPU parity of the device multigrid transfer, the Krylov vector kernels and the solver stack built on them
 against the CPU restatement in oracle/gls_solver.py: same vectors to
round-off, identical GMRES / Newton iteration counts."""
import numpy as np
import pytest

from dealii_ns_gls_b200 import mesh as M
from dealii_ns_gls_b200.driver import ChannelParameters, Driver
from oracle import gls_solver as gs
from tests.test_solver_oracle import _oracle_driver
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype=dtype)


@pytest.mark.parametrize("number,tol", [("double", 1e-13), ("float", 2e-6)])
@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (2, 3), (2, 4), (3, 1), (3, 2), (3, 3), (3, 4)])
def test_transfer_matches_explicit_matrices(dim, degree, number, tol):
    from dealii_ns_gls_b200.multigrid import MGTwoLevelTransfer
    from dealii_ns_gls_b200.operator import AffineConstraints
    shape = (3, 2, 2)[:dim]
    eps = 1e-12

    def bc(x, c):  # zero constraints on two faces for the velocity, one for the pressure
        return (np.abs(x[:, 0]) < eps) | ((c < dim) & (np.abs(x[:, 1] - 1.0) < eps))

    mc = M.structured_mesh(dim, shape, degree, dirichlet=bc)
    mf = M.structured_mesh(dim, tuple(2 * s for s in shape), degree, dirichlet=bc)
    # one weighted (hanging-node-like) coarse constraint: dof a = 0.5 b + 0.5 c
    free = [d for d in range(mc.n_dofs) if d not in mc.constraints]
    mc.constraints[free[5]] = [(free[7], 0.5), (free[11], 0.5)]
    ch = M.child_cells(mc, mf)
    fd, cd = mf.cell_dofs.astype(np.int64), mc.cell_dofs.astype(np.int64)
    P = gs.prolongation_matrix(dim, degree, fd, cd, ch, mf.n_dofs, mc.n_dofs, mf.constraints, mc.constraints)
    R = gs.interpolation_matrix(dim, degree, fd, cd, ch, mf.n_dofs, mc.n_dofs)
    dt = torch.float64 if number == "double" else torch.float32
    t = MGTwoLevelTransfer().reinit(mf, mc, AffineConstraints(mf.constraints), AffineConstraints(mc.constraints),
                                    number=number)
    t0 = MGTwoLevelTransfer().reinit(mf, mc, None, None, number=number)
    rng = np.random.default_rng(3)
    xc, xf = rng.standard_normal(mc.n_dofs), rng.standard_normal(mf.n_dofs)
    yf0, yc0 = rng.standard_normal(mf.n_dofs), rng.standard_normal(mc.n_dofs)
    yf = _dev(yf0, dt)
    t.prolongate_and_add(yf, _dev(xc, dt))
    assert rel_l2(yf.cpu().numpy(), yf0 + P @ xc) < tol
    yc = _dev(yc0, dt)
    t.restrict_and_add(yc, _dev(xf, dt))
    assert rel_l2(yc.cpu().numpy(), yc0 + P.T @ xf) < tol
    zc = torch.zeros(mc.n_dofs, dtype=dt, device="cuda")
    t0.interpolate(zc, _dev(xf, dt))
    assert rel_l2(zc.cpu().numpy(), R @ xf) < tol


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_vector_kernels(dt):
    from dealii_ns_gls_b200.multigrid import DeviceVectorOps
    ops = DeviceVectorOps()
    rng = np.random.default_rng(5)
    n, k = 100_003, 19
    V, w = rng.standard_normal((k, n)), rng.standard_normal(n)
    Vd, wd = _dev(V, dt), _dev(w, dt)
    tol = 1e-13 if dt == torch.float64 else 1e-5
    out = torch.zeros(k, dtype=torch.float64, device="cuda")
    ops.multi_dot(out, Vd, k, wd)
    ref = Vd.double().cpu().numpy() @ wd.double().cpu().numpy()
    assert np.allclose(out.cpu().numpy(), ref, rtol=1e-12, atol=1e-9)
    out2 = torch.zeros(k, dtype=torch.float64, device="cuda")
    ops.multi_dot(out2, Vd, k, wd)
    assert torch.equal(out, out2)  # deterministic reduction
    coef = _dev(rng.standard_normal(k), torch.float64)
    w2 = wd.clone()
    ops.multi_axpy(w2, Vd, k, coef, -1.0)
    assert rel_l2(w2.cpu().numpy(), w - coef.cpu().numpy() @ V) < 10 * tol
    y = _dev(V[0], dt)
    ops.axpby(y, 2.0, wd, -0.5)
    assert rel_l2(y.cpu().numpy(), 2 * w - 0.5 * V[0]) < 10 * tol
    ops.axpby(y, 3.0, wd, 0.0)
    assert rel_l2(y.cpu().numpy(), 3 * w) < tol
    other = torch.float32 if dt == torch.float64 else torch.float64
    z = torch.empty(n, dtype=other, device="cuda")
    ops.convert(z, wd)
    assert torch.equal(z, wd.to(other))
    idx = torch.tensor([0, 5, n - 1], dtype=torch.int32, device="cuda")
    ops.set_zero_indexed(z, idx)
    assert z[0] == 0 and z[5] == 0 and z[n - 1] == 0 and z[1] == wd[1].to(other)
    A = rng.standard_normal((37, 53))
    x = rng.standard_normal(53)
    yy = torch.empty(37, dtype=dt, device="cuda")
    ops.dense_apply(yy, _dev(A, torch.float64), _dev(x, dt))
    assert rel_l2(yy.cpu().numpy(), A @ x) < 10 * tol


def _drivers(**kw):
    p = ChannelParameters(**kw)
    lv = np.float64 if p.mg_number == "double" else np.float32
    return Driver(p), _oracle_driver(p, level_dtype=lv)


@pytest.mark.parametrize("mg_number,tol", [("double", 1e-9), ("float", 2e-4)])
def test_vcycle_matches_oracle(mg_number, tol):
    """PreconditionerGMG::vmult (one V-cycle with relaxation smoothers, device transfers, dense coarse solve)"""
    dev, ora = _drivers(n_global_refinements=0, mg_number=mg_number)
    rec_o = ora.step()
    # same linearization point and time-step data on both sides
    sol = torch.from_numpy(ora.history[0]).cuda()
    dev.time_integrator_data.update_dt(rec_o["dt"])
    dev.solution.solutions[1].copy_(torch.from_numpy(ora.history[1]).cuda())
    dev.set_previous_solution(dev.solution)
    dev.nonlinear_solver.setup_jacobian(sol)
    dev.nonlinear_solver.setup_preconditioner(sol)
    ora._setup_preconditioner(ora.history[0], rec_o["dt"])
    for l in dev.preconditioner.smoothers:
        if l > 0:
            assert abs(dev.preconditioner.smoothers[l].relaxation / ora.gmg.smoothers[l].relaxation - 1) < tol
    b = np.random.default_rng(11).standard_normal(sol.numel())
    b[ora.op.constrained] = 0
    dst = torch.zeros_like(sol)
    dev.preconditioner.vmult(dst, torch.from_numpy(b).cuda())
    assert rel_l2(dst.cpu().numpy(), ora.gmg.vmult(b)) < tol


@pytest.mark.parametrize("kw", [dict(n_global_refinements=0), dict(n_global_refinements=1),
                                dict(n_global_refinements=0, fe_degree=2),
                                dict(dim=3, n_global_refinements=0, fe_degree=1),
                                dict(n_global_refinements=0, bdf_order=2),
                                dict(n_global_refinements=0, cell_wise_stabilization=False, nu=0.01),
                                dict(n_global_refinements=0, time_integration="none", fe_degree=2,
                                     cell_wise_stabilization=False, nu=0.05)],
                         ids=["q1", "q1_r1", "q2", "3d_q1", "bdf2", "qwise", "stationary_q2"])
@pytest.mark.parametrize("mg_number", ["double", "float"])
def test_channel_time_steps_identical_iteration_counts(kw, mg_number):
    """the north-star's solver-level criterion, as far as it can be checked without deal.II: the Newton and
    GMRES iteration counts of the time loop with the device operators / transfers / Krylov kernels equal those
    of the CPU restatement, step by step, and the solutions agree"""
    dev, ora = _drivers(mg_number=mg_number, **kw)
    for _ in range(3):
        rd, ro = dev.step(), ora.step()
        assert rd["newton_iterations"] == ro["newton_iterations"]
        assert rd["linear_iterations"] == ro["linear_iterations"]
        # FP32 level operators (config.h:7) change the V-cycle, and so the inexact (1e-2) linear solves, at 1e-5
        tol = 1e-6 if mg_number == "double" else 1e-3
        assert abs(rd["dt"] / ro["dt"] - 1) < tol
        assert np.allclose(rd["newton_residuals"][:2], ro["newton_residuals"][:2], rtol=tol)
        assert rel_l2(dev.solution.get_current_solution().cpu().numpy(), ora.history[0]) < tol


@pytest.mark.parametrize("name", ["q1", "q1_r1", "q2", "3d_q1", "bdf2", "qwise", "stationary_q2"])
def test_channel_time_steps_match_golden_record(name):
    """the device time loop (float level operators, config.h:7) against the committed record of the oracle's
    solver stack, tests/golden/solver_channel.json: iteration counts identical, norms to the float-level tolerance"""
    from tests.test_solver_oracle import _golden
    g = _golden()[name]
    dev = Driver(ChannelParameters(mg_number="float", **g["parameters"]))
    for ref in g["steps"]:
        r = dev.step()
        assert r["newton_iterations"] == ref["newton_iterations"]
        assert r["linear_iterations"] == ref["linear_iterations"]
        assert abs(r["dt"] / ref["dt"] - 1) < 1e-3
        assert np.allclose(r["newton_residuals"][:2], ref["first_residuals"], rtol=1e-3)
        l2 = float(torch.linalg.vector_norm(dev.solution.get_current_solution()))
        assert abs(l2 / ref["solution_l2"] - 1) < 1e-3


def test_vcycle_as_cuda_graph_matches_eager():
    """the V-cycle captured into a CUDA graph (PreconditionerGMGAdditionalData.use_cuda_graph) gives the same
    vectors as the eager launch sequence and the same iteration counts in the time loop"""
    from dealii_ns_gls_b200.multigrid import PreconditionerGMGAdditionalData
    kw = dict(dim=3, fe_degree=2, n_global_refinements=0, newton_inexact=True)
    eager = Driver(ChannelParameters(**kw))
    graph = Driver(ChannelParameters(gmg=PreconditionerGMGAdditionalData(use_cuda_graph=True), **kw))
    for _ in range(2):
        re_, rg = eager.step(), graph.step()
        assert re_["newton_iterations"] == rg["newton_iterations"]
        assert re_["linear_iterations"] == rg["linear_iterations"]
    a, b = eager.solution.get_current_solution(), graph.solution.get_current_solution()
    assert rel_l2(b.cpu().numpy(), a.cpu().numpy()) < 1e-6
    src = torch.randn_like(a)
    y0, y1 = torch.zeros_like(a), torch.zeros_like(a)
    graph.preconditioner.use_cuda_graph = False
    graph.preconditioner.vmult(y0, src)
    graph.preconditioner.use_cuda_graph = True
    graph.preconditioner.vmult(y1, src)
    graph.preconditioner.vmult(y1, src)
    assert rel_l2(y1.cpu().numpy(), y0.cpu().numpy()) < 1e-5  # float levels, atomics: not bit-identical


# ---- the BASELINE configs that are not the channel: curved O-grid, no-slip cylinder rows, q-point-wise delta ----
def _cylinder_golden():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solver_cylinder.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("kw,n_steps", [
    (dict(dim=2, time_integration="none", c_1=1.0, u_max=0.3, n_global_refinements=2), 1),
    (dict(dim=2, n_global_refinements=2, newton_inexact=True, u_max=0.3), 2)], ids=["turek2d_stat", "turek2d_bdf2"])
@pytest.mark.parametrize("mg_number", ["double", "float"])
def test_cylinder_time_steps_identical_iteration_counts(kw, n_steps, mg_number):
    """input_turek_2D_Re20_stat.json (stationary, exact Newton) and its BDF2 / inexact-Newton variant on the
    synthetic O-grid: device loop against the CPU restatement run side by side, counts identical step by step"""
    from dealii_ns_gls_b200.driver import CylinderParameters
    p = CylinderParameters(mg_number=mg_number, **kw)
    dev = Driver(p)
    ora = _oracle_driver(p, level_dtype=np.float64 if mg_number == "double" else np.float32)
    for _ in range(n_steps):
        rd, ro = dev.step(), ora.step()
        assert rd["newton_iterations"] == ro["newton_iterations"]
        assert rd["linear_iterations"] == ro["linear_iterations"]
        tol = 1e-6 if mg_number == "double" else 1e-3
        assert abs(rd["dt"] / ro["dt"] - 1) < tol
        assert np.allclose(rd["newton_residuals"][:2], ro["newton_residuals"][:2], rtol=tol)
        assert rel_l2(dev.solution.get_current_solution().cpu().numpy(), ora.history[0]) < tol


@pytest.mark.parametrize("name", ["turek2d_stat", "turek2d_bdf2", "turek3d_bdf2", "hoffmann3d_slip"])
def test_cylinder_time_steps_match_golden_record(name):
    """the device time loop (float level operators) against the committed record of the oracle's solver stack for
    the Turek / Hoffmann-like configurations, tests/golden/solver_cylinder.json (3-D BDF2 + inexact Newton +
    q-point-wise delta with no-slip and with slip walls included): iteration counts identical"""
    from dealii_ns_gls_b200.driver import CylinderParameters
    g = _cylinder_golden()[name]
    kw = dict(g["parameters"])
    if "base_shape" in kw:
        kw["base_shape"] = tuple(kw["base_shape"])
    kw.setdefault("mg_number", "float")  # the record says which level number type it was made with
    dev = Driver(CylinderParameters(**kw))
    for ref in g["steps"]:
        r = dev.step()
        assert r["newton_iterations"] == ref["newton_iterations"]
        assert r["linear_iterations"] == ref["linear_iterations"]
        assert abs(r["dt"] / ref["dt"] - 1) < 1e-3
        assert np.allclose(r["newton_residuals"][:2], ref["first_residuals"], rtol=1e-3)
        l2 = float(torch.linalg.vector_norm(dev.solution.get_current_solution()))
        assert abs(l2 / ref["solution_l2"] - 1) < 1e-3
