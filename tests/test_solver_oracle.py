"""Known-answer tests of the CPU restatement of the solver stack (oracle/gls_solver.py): two-level
transfer, V-cycle, GMRES, Newton, channel time loop.  No GPU."""
import numpy as np
import pytest

from dealii_ns_gls_b200 import mesh as M
from dealii_ns_gls_b200.driver import (ChannelParameters, channel_inhomogeneous_constraints,
                                       channel_level_mesh)
from oracle import gls_solver as gs


def _pair(dim, degree, shape):
    mc = M.structured_mesh(dim, shape, degree)
    mf = M.structured_mesh(dim, tuple(2 * s for s in shape), degree)
    return mc, mf, M.child_cells(mc, mf)


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (3, 2), (2, 3), (3, 1)])
def test_prolongation_reproduces_polynomials(dim, degree):
    """the FE_Q(p) embedding is exact for polynomials of degree <= p per direction"""
    mc, mf, ch = _pair(dim, degree, (3, 2, 2)[:dim])
    P = gs.prolongation_matrix(dim, degree, mf.cell_dofs.astype(np.int64), mc.cell_dofs.astype(np.int64), ch,
                               mf.n_dofs, mc.n_dofs)
    xc, xf = M.dof_coordinates(mc), M.dof_coordinates(mf)
    cc, cf = M.dof_components(mc), M.dof_components(mf)

    def f(x, c):
        return (1 + c) * np.prod(1.0 + x ** degree - 0.3 * x, axis=1)

    assert np.allclose(P @ f(xc, cc), f(xf, cf), atol=1e-12)
    R = gs.interpolation_matrix(dim, degree, mf.cell_dofs, mc.cell_dofs, ch, mf.n_dofs, mc.n_dofs)
    assert np.allclose(R @ f(xf, cf), f(xc, cc), atol=1e-12)
    assert abs(R @ P - np.eye(mc.n_dofs)).max() < 1e-12


def test_prolongation_constraints_and_weights():
    """zero rows on constrained fine dofs, constrained coarse dofs read as their constraint row"""
    p = ChannelParameters()
    mc, mf = channel_level_mesh(p, 1), channel_level_mesh(p, 2)
    ch = M.child_cells(mc, mf)
    P = gs.prolongation_matrix(2, 1, mf.cell_dofs.astype(np.int64), mc.cell_dofs.astype(np.int64), ch, mf.n_dofs,
                               mc.n_dofs, mf.constraints, mc.constraints)
    fc = np.array(sorted(mf.constraints))
    cc = np.array(sorted(mc.constraints))
    assert abs(P[fc]).sum() == 0
    assert abs(P[:, cc]).sum() == 0
    free = np.setdiff1d(np.arange(mf.n_dofs), fc)
    # a coarse vector that satisfies the (zero) constraints is prolongated like without constraints
    P0 = gs.prolongation_matrix(2, 1, mf.cell_dofs.astype(np.int64), mc.cell_dofs.astype(np.int64), ch, mf.n_dofs,
                                mc.n_dofs)
    v = np.random.default_rng(0).standard_normal(mc.n_dofs)
    v[cc] = 0
    assert np.allclose((P @ v)[free], (P0 @ v)[free])


def test_gmres_matches_dense_solve():
    rng = np.random.default_rng(1)
    n = 60
    A = np.eye(n) * 4 + rng.standard_normal((n, n)) * 0.3
    b = rng.standard_normal(n)
    x, it, hist = gs.gmres(lambda v: A @ v, lambda v: v / 4, b, rel_tol=1e-10, basis=28)
    assert np.allclose(x, np.linalg.solve(A, b), atol=1e-8)
    assert hist[-1] <= 1e-10 * np.linalg.norm(b) and it == len(hist) - 1 - (it // 28)


def _oracle_driver(p, **kw):
    n_levels = p.n_levels()
    meshes = {l: p.level_mesh(l) for l in range(n_levels + 1)}
    children = {l: M.child_cells(meshes[l - 1], meshes[l]) for l in range(1, n_levels + 1)}
    ci = p.inhomogeneous_constraints(meshes[n_levels])
    kw.setdefault("newton_inexact", p.newton_inexact)
    return gs.OracleChannelDriver(dim=p.dim, degree=p.fe_degree, meshes=meshes, children=children,
                                  constraints_inhomogeneous=ci.rows, inhomogeneities=ci.inhomogeneities,
                                  min_dx=p.minimal_cell_diameter(meshes[n_levels]), nu=p.nu, c1=p.c_1, c2=p.c_2, cfl=p.cfl,
                                  bdf_order=p.bdf_order if p.time_integration == "bdf" else 0,
                                  consider_time_derivative=p.consider_time_derivative,
                                  cell_wise_stabilization=p.cell_wise_stabilization,
                                  rel_tol=p.lin_relative_tolerance, abs_tol=p.lin_absolute_tolerance, **kw)


def test_channel_time_steps_converge():
    """input_channel.json at n_global_refinements = 0: Newton converges in a few steps, GMRES + GMG in a
    handful of iterations, and the V-cycle is a contraction"""
    p = ChannelParameters(n_global_refinements=0)
    d = _oracle_driver(p)
    for _ in range(2):
        rec = d.step()
        assert rec["newton_residuals"][-1] <= 1e-7
        assert 1 <= rec["newton_iterations"] <= 8
        assert all(1 <= k <= 15 for k in rec["linear_iterations"])
    # inflow profile is kept, walls stay at rest
    sol = d.history[0]
    assert np.allclose(sol[d.cdofs], d.cvals)


def _golden():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solver_channel.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["q1", "qwise", "stationary_q2"])
def test_solver_stack_matches_golden_record(name):
    """tests/golden/solver_channel.json pins the oracle's Newton / GMRES iteration counts, time-step sizes and
    solution norms (a cheap subset here; the GPU suite checks the device path against every case)"""
    g = _golden()[name]
    d = _oracle_driver(ChannelParameters(**g["parameters"]))
    for ref in g["steps"]:
        r = d.step()
        assert r["newton_iterations"] == ref["newton_iterations"]
        assert r["linear_iterations"] == ref["linear_iterations"]
        assert abs(r["dt"] / ref["dt"] - 1) < 1e-10
        assert np.allclose(r["newton_residuals"][:2], ref["first_residuals"], rtol=1e-8)
        assert abs(np.linalg.norm(d.history[0]) / ref["solution_l2"] - 1) < 1e-8


@pytest.mark.parametrize("name", ["turek2d_stat", "turek2d_bdf2"])
def test_cylinder_solver_stack_matches_golden_record(name):
    """tests/golden/solver_cylinder.json pins the oracle's counts for the Turek-like configurations on the synthetic
    O-grid (curved cells, no-slip cylinder rows, q-point-wise delta, float level operators); the 2-D cases here, the
    3-D ones (minutes on the CPU) are checked through the device path in the GPU suite"""
    import json
    import os
    from dealii_ns_gls_b200.driver import CylinderParameters
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solver_cylinder.json")) as f:
        g = json.load(f)[name]
    d = _oracle_driver(CylinderParameters(**g["parameters"]), level_dtype=np.float32)
    for ref in g["steps"]:
        r = d.step()
        assert r["newton_iterations"] == ref["newton_iterations"]
        assert r["linear_iterations"] == ref["linear_iterations"]
        assert abs(r["dt"] / ref["dt"] - 1) < 1e-10
        assert np.allclose(r["newton_residuals"][:2], ref["first_residuals"], rtol=1e-6)
    # the cylinder stays at rest and the inflow value is kept
    assert np.allclose(d.history[0][d.cdofs], d.cvals)


def test_reference_log_parser_round_trip():
    """reflog.py reads the console lines of the reference (main.cc:921-923, :971; solver_nl.cc:53, :79, :88;
    solver_l.cc:70) -- here a transcript typed from those format strings -- and the lines the device driver prints,
    and diffs the iteration counts"""
    from dealii_ns_gls_b200 import reflog
    ref = ("    [I] Number of active cells:    1024\n    [I] Global degrees of freedom: 3267\n"
           "\ncycle\t1 at time t = 0 with delta_t = 0.00441942 and u_max = 1\n"
           "    [N] step 0; residual = 0.0721688\n    [L] solved in 3 iterations.\n"
           "    [N] step 1 ; residual = 0.00123\n    [L] solved in 4 iterations.\n"
           "    [N] step 2 ; residual = 3.1e-08\n    [N] solved in 2 iterations.\n"
           "    [S] l2-norm of solution: 25.7391\n"
           "\ncycle\t2 at time t = 0.00441942 with delta_t = 0.00441942 and u_max = 1\n"
           "    [N] step 0; residual = 0.01\n    [L] solved in 2 iterations.\n"
           "    [N] step 1 ; residual = 9e-09\n    [N] solved in 1 iterations.\n"
           "    [S] l2-norm of solution: 25.74\n")
    a = reflog.parse(ref)
    assert [s["newton_iterations"] for s in a] == [2, 1]
    assert [s["linear_iterations"] for s in a] == [[3, 4], [2]]
    assert a[0]["dt"] == 0.00441942 and a[1]["solution_l2"] == 25.74 and len(a[0]["newton_residuals"]) == 3
    # what the driver prints parses back to the same record
    b = reflog.parse("\n".join(reflog.format_step(s) for s in a))
    assert reflog.compare(a, b) == []
    b[1]["linear_iterations"] = [3]
    assert reflog.compare(a, b) == ["cycle 2: GMRES iterations [2] vs [3]"]
