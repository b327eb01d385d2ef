"""The oracle's quadrature-point physics and stabilisation parameters against the REFERENCE'S OWN CODE.

include/operator_ns.cc as a whole needs deal.II, but the part of it that is the reference's own arithmetic -- what
do_vmult_cell computes between get_value / get_gradient and submit_value / submit_gradient in both branches
(:880-1182), the same for the outflow faces in do_vmult_boundary (:1195-1301), and the body of
compute_penalty_parameters' cell loop (:348-421) -- only touches Tensor,
VectorizedArray, Table and a few FEEvaluation accessors.  oracle/build_ref_qpoint.sh cuts those lines out of
/root/reference at build time and compiles them unmodified on stand-in types (oracle/ref_shim/qpoint_shim.h,
oracle/ref_qpoint_harness.cc) into oracle/_ref/libref_qpoint.so; tests/golden/make_golden_reference_qpoint.py
recorded its output for seeded inputs in tests/golden/reference/qpoint.npz.

Checked here: (1) where the object code is available, that it reproduces the record bit for bit; (2) always, that
gls_oracle.OracleOperator._cell_newton / _cell_fixed_point / _penalty -- the functions every parity test of the
CUDA path ends in -- agree with the record to round-off (summation order) in all 20 branch / flag cases, 8
boundary-face cases and 5 parameter cases.  deal.II's own parts (sum-factorised evaluate / integrate, geometry, constraints, vector access)
stay restated from its documentation; see DESIGN.md section 1 for what is and is not pinned."""
import importlib.util
import os

import numpy as np
import pytest

from dealii_ns_gls_b200 import mesh as gm
from oracle import gls_oracle as go
from oracle import ref_qpoint as rq

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_reference_qpoint",
                                               os.path.join(HERE, "golden", "make_golden_reference_qpoint.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)


@pytest.fixture(scope="module")
def record():
    return dict(np.load(os.path.join(HERE, "golden", "reference", "qpoint.npz")))


def test_object_code_of_the_reference_reproduces_the_record(record):
    if rq.load() is None:
        pytest.skip("oracle/_ref/libref_qpoint.so not built (no reference tree on this machine)")
    assert np.array_equal(record["cases"], np.array(gen.CASES, dtype=np.float64))
    for i, case in enumerate(gen.CASES):
        _, (v, g) = gen.run_case(i, case)
        assert np.array_equal(v, record[f"value_out_{i}"]) and np.array_equal(g, record[f"grad_out_{i}"]), case
    for i, case in enumerate(gen.BOUNDARY):
        _, (v, g) = gen.run_boundary(i, case)
        assert np.array_equal(v, record[f"boundary_value_out_{i}"]) and np.array_equal(g, record[f"boundary_grad_out_{i}"])
    for i, (dim, degree, dt, nu) in enumerate(gen.PENALTY):
        r = rq.penalty(dim=dim, dt=dt, nu=nu, c1=4.0, c2=2.0, degree=degree, **gen.penalty_inputs(i, dim, degree))
        for name, arr in zip(("d1_cell", "d2_cell", "d1_q", "d2_q"), r):
            assert np.array_equal(arr, record[f"{name}_{i}"]), (name, i)


def _oracle(dim, degree, **kw):
    m = gm.hypercube(dim, 1, degree)
    return go.OracleOperator(dim=dim, degree=degree, cell_dofs=m.cell_dofs, n_dofs=m.n_dofs,
                             cell_points=m.cell_points, mapping_degree=1, constraints={}, path="sumfac", **kw)


@pytest.mark.parametrize("i", range(len(gen.CASES)))
def test_oracle_qpoint_physics_equals_the_reference(i, record):
    dim, res, inc, ctd, cw, th, old = gen.CASES[i]
    dim, res, inc, ctd, cw, old = int(dim), bool(res), bool(inc), bool(ctd), bool(cw), bool(old)
    a = gen.inputs(i, dim, 3 ** dim)
    # order > 0 whenever a time-derivative table exists; consider_time_derivative is the member the reference
    # reads (operator_ns.cc:97-98: the flag and order > 0)
    o = _oracle(dim, 2, nu=0.037, c1=4.0, c2=2.0, theta=th, order=2 if (old or ctd) else 0,
                consider_time_derivative=ctd, increment_form=inc, cell_wise_stabilization=cw)
    assert o.ctd == ctd
    t = lambda x: np.ascontiguousarray(np.moveaxis(x, 0, -1))[None]   # [q, ...] -> [1 cell, ..., q]  # noqa: E731
    o.U, o.H, o.P = t(a["u_star"]), t(a["u_star_grad"]), t(a["p_star_grad"])
    o.o = t(a["u_tdo"]) if old else None
    o.Gold, o.gold_p = (t(a["u_old_grad"]), t(a["p_old_grad"])) if th != 1.0 else (None, None)
    o.delta1_cell, o.delta2_cell = a["d1"][:1].copy(), a["d2"][:1].copy()
    o.delta1_q, o.delta2_q = a["d1"][None].copy(), a["d2"][None].copy()
    val, grad = t(a["value"]), t(a["grad"])
    if res or not inc:
        vo, go_ = o._cell_fixed_point(val, grad, 7.25, res)
    else:
        vo, go_ = o._cell_newton(val, grad, 7.25)
    ref_v, ref_g = record[f"value_out_{i}"], record[f"grad_out_{i}"]
    got_v, got_g = np.moveaxis(vo[0], -1, 0), np.moveaxis(go_[0], -1, 0)
    scale = max(np.abs(ref_v).max(), np.abs(ref_g).max())
    assert np.abs(got_v - ref_v).max() <= 4e-15 * scale and np.abs(got_g - ref_g).max() <= 4e-15 * scale, gen.CASES[i]


@pytest.mark.parametrize("i", range(len(gen.CASES)))
def test_oracle_qpoint_physics_in_float_equals_the_reference(i, record):
    """Number = float, the multigrid level operators (config.h:7): the oracle's float32 path against the float
    instantiation of the reference's kernel, to float round-off"""
    dim, res, inc, ctd, cw, th, old = gen.CASES[i]
    dim, res, inc, ctd, cw, old = int(dim), bool(res), bool(inc), bool(ctd), bool(cw), bool(old)
    if rq.load() is not None:
        _, (v, g) = gen.run_case(i, gen.CASES[i], number="float")
        assert np.array_equal(v, record[f"value_out_f32_{i}"]) and np.array_equal(g, record[f"grad_out_f32_{i}"])
    a = gen.inputs(i, dim, 3 ** dim)
    o = _oracle(dim, 2, nu=0.037, c1=4.0, c2=2.0, theta=th, order=2 if (old or ctd) else 0,
                consider_time_derivative=ctd, increment_form=inc, cell_wise_stabilization=cw, dtype=np.float32)
    t = lambda x: np.ascontiguousarray(np.moveaxis(x, 0, -1))[None].astype(np.float32)  # noqa: E731
    o.U, o.H, o.P = t(a["u_star"]), t(a["u_star_grad"]), t(a["p_star_grad"])
    o.o = t(a["u_tdo"]) if old else None
    o.Gold, o.gold_p = (t(a["u_old_grad"]), t(a["p_old_grad"])) if th != 1.0 else (None, None)
    o.delta1_cell, o.delta2_cell = a["d1"][:1].astype(np.float32), a["d2"][:1].astype(np.float32)
    o.delta1_q, o.delta2_q = a["d1"][None].astype(np.float32), a["d2"][None].astype(np.float32)
    val, grad = t(a["value"]), t(a["grad"])
    if res or not inc:
        vo, go_ = o._cell_fixed_point(val, grad, 7.25, res)
    else:
        vo, go_ = o._cell_newton(val, grad, 7.25)
    assert vo.dtype == np.float32 and go_.dtype == np.float32
    ref_v, ref_g = record[f"value_out_f32_{i}"], record[f"grad_out_f32_{i}"]
    scale = max(np.abs(ref_v).max(), np.abs(ref_g).max())
    assert np.abs(np.moveaxis(vo[0], -1, 0) - ref_v).max() <= 2e-6 * scale
    assert np.abs(np.moveaxis(go_[0], -1, 0) - ref_g).max() <= 2e-6 * scale
    # and the float result is the double result to float accuracy (nothing else changed)
    assert np.abs(ref_v - record[f"value_out_{i}"]).max() <= 1e-5 * scale


@pytest.mark.parametrize("i", range(len(gen.BOUNDARY)))
def test_oracle_outflow_face_physics_equals_the_reference(i, record):
    """do_vmult_boundary (operator_ns.cc:1195-1301) at the face quadrature points: cut faces (v, beta min(0, U.n) u)
    with U = the iterate in the residual and face_velocity otherwise, Nitsche faces with u - u_target in the
    residual; nothing on the pressure rows"""
    dim, kind, res = (int(x) for x in gen.BOUNDARY[i])
    a = gen.boundary_inputs(i, dim)
    o = _oracle(dim, 2, nu=0.037, c1=4.0, c2=2.0, theta=1.0, order=1, consider_time_derivative=True,
                increment_form=True, cell_wise_stabilization=True)
    val = np.ascontiguousarray(a["value"][:, :dim].T)[None]                      # [1 face, d, q]
    grad = np.ascontiguousarray(a["grad"][:, :dim].transpose(1, 2, 0))[None]     # [1, d, j, q]
    vr, gr = o._face_qpoint(np.array([kind]), val, grad, a["normal"][None], np.array([3.7]),
                            a["face_velocity"][None], a["target"][None, :, :dim], bool(res))
    ref_v, ref_g = record[f"boundary_value_out_{i}"], record[f"boundary_grad_out_{i}"]
    assert np.abs(ref_v[:, dim]).max() == 0 and np.abs(ref_g[:, dim]).max() == 0
    scale = max(np.abs(ref_v).max(), np.abs(ref_g).max())
    assert np.abs(vr[0].T - ref_v[:, :dim]).max() <= 4e-15 * scale
    assert np.abs(gr[0].transpose(2, 0, 1) - ref_g[:, :dim]).max() <= 4e-15 * scale


@pytest.mark.parametrize("i", range(len(gen.PENALTY)))
def test_oracle_stabilisation_parameters_equal_the_reference(i, record):
    dim, degree, dt, nu = gen.PENALTY[i]
    dim, degree = int(dim), int(degree)
    a = gen.penalty_inputs(i, dim, degree)
    o = _oracle(dim, degree, nu=nu, c1=4.0, c2=2.0, theta=1.0, order=2, consider_time_derivative=True,
                increment_form=True, cell_wise_stabilization=True)
    o.h_min, o.measure = a["h_min"].copy(), a["measure"].copy()
    o._penalty(np.ascontiguousarray(a["u"].transpose(0, 2, 1)), dt)
    for name, got in (("d1_cell", o.delta1_cell), ("d2_cell", o.delta2_cell), ("d1_q", o.delta1_q),
                      ("d2_q", o.delta2_q)):
        ref = record[f"{name}_{i}"]
        assert np.abs(got / ref - 1.0).max() <= 1e-15 if "cell" in name else np.abs(got / ref - 1.0).max() <= 2e-15, name


def test_both_regimes_of_the_cell_wise_parameters_are_in_the_record():
    """nu < h (convection-dominated formula) and nu >= h occur among the recorded cells, and so does dt = 0"""
    below = above = 0
    for i, (dim, degree, dt, nu) in enumerate(gen.PENALTY):
        h = gen.penalty_inputs(i, int(dim), int(degree))["h_min"]
        below += int((nu < h).sum())
        above += int((nu >= h).sum())
    assert below > 10 and above > 10
    assert any(dt == 0.0 for _, _, dt, _ in gen.PENALTY)


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (3, 2), (3, 3), (3, 4)])
def test_oracle_outflow_face_penalty_equals_the_reference(dim, degree, record):
    """effective_beta_face = 1 / h^(p+1), h after Lethe from the cell measure (operator_ns.cc:428-457) against
    set_outflow_faces of the oracle"""
    if rq.load() is not None:
        assert np.array_equal(rq.face_beta(dim=dim, degree=degree, measure=record["beta_measure"]),
                              record[f"beta_{dim}_{degree}"])
    n_cells = len(record["beta_measure"])
    shape = (n_cells,) + (1,) * (dim - 1)
    m = gm.structured_mesh(dim, shape, degree)
    o = go.OracleOperator(dim=dim, degree=degree, cell_dofs=m.cell_dofs, n_dofs=m.n_dofs, cell_points=m.cell_points,
                          mapping_degree=1, constraints={}, nu=0.01, c1=4.0, c2=2.0, theta=1.0, order=1,
                          consider_time_derivative=True, increment_form=True, cell_wise_stabilization=True, path="sumfac")
    o.measure = record["beta_measure"].copy()
    o.set_outflow_faces(np.arange(n_cells), np.full(n_cells, 2), np.ones(n_cells, dtype=np.int64))
    assert np.abs(o.faces["beta"] / record[f"beta_{dim}_{degree}"] - 1.0).max() < 4e-15


# ---- the PRODUCT's CUDA source against the reference record, on the host ------------------------------------------
def _product_lib():
    import ctypes as C
    import subprocess
    path = os.path.join(HERE, "cpp", "libprod_qpoint.so")
    if not os.path.exists(path):
        subprocess.check_call(["bash", os.path.join(HERE, "cpp", "build_qpoint_host.sh")])
    lib = C.CDLL(path)
    P, D, I = C.c_void_p, C.c_double, C.c_int
    lib.prod_qpoint.restype = I
    lib.prod_qpoint.argtypes = [I, I, I, D, D, D, I, I, I] + [P] * 10 + [I, P, P]
    return lib


@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("i", range(len(gen.CASES)))
def test_cuda_source_qpoint_physics_equals_the_reference(i, number, record):
    """qpoint_physics of dealii_ns_gls_b200/csrc/glsb_kernels.cuh (what the generic, column, diagonal and residual
    kernels run per quadrature point), compiled for the host by tests/cpp/build_qpoint_host.sh, on the inputs of the
    reference record: the CUDA source against the reference's own do_vmult_cell, no GPU and no oracle in between"""
    lib = _product_lib()
    dim, res, inc, ctd, cw, th, old = gen.CASES[i]
    dim, res, inc, ctd, cw, old = int(dim), int(res), int(inc), int(ctd), int(cw), int(old)
    a = gen.inputs(i, dim, 3 ** dim)
    branch = 2 if res else (0 if inc else 1)
    f = lambda x: np.ascontiguousarray(x, dtype=np.float64)  # noqa: E731
    arrs = [f(a[k]) for k in ("value", "grad", "u_star", "u_star_grad", "p_star_grad", "u_tdo", "u_old_grad",
                              "p_old_grad", "d1", "d2")]
    vo, go_ = np.zeros_like(arrs[0]), np.zeros_like(arrs[1])
    ptr = lambda x: x.ctypes.data_as(__import__("ctypes").c_void_p)  # noqa: E731
    rc = lib.prod_qpoint(int(number == "float"), dim, branch, float(th), 0.037, 7.25, ctd, old, 3 ** dim,
                         *[ptr(x) for x in arrs], cw, ptr(vo), ptr(go_))
    assert rc == 0
    sfx, tol = ("", 4e-15) if number == "double" else ("_f32", 2e-6)   # float: the reference's float instantiation
    ref_v, ref_g = record[f"value_out{sfx}_{i}"], record[f"grad_out{sfx}_{i}"]
    scale = max(np.abs(ref_v).max(), np.abs(ref_g).max())
    assert np.abs(vo - ref_v).max() <= tol * scale and np.abs(go_ - ref_g).max() <= tol * scale, gen.CASES[i]
