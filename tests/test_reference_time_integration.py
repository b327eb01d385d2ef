"""A pin against the reference ITSELF, for the one file on the path that compiles without deal.II:
include/time_integration.cc (BDF1-3 with variable step sizes, theta scheme, stationary scheme, SolutionHistory).
`make -C oracle _ref` compiles it unmodified from /root/reference against two stand-in headers
(oracle/ref_shim/) into oracle/_ref/libref_time_integration.so; tests/golden/make_golden_reference_ti.py
recorded its output in tests/golden/reference_time_integration.json.

Checked here: (1) where the object code is available, that it still produces the committed record; (2) always,
that the restatements -- oracle/gls_oracle.py OracleBDF (what every oracle time loop uses),
dealii_ns_gls_b200/time_integration.py (what the device operator reads its weights from) and the stationary
stand-in of oracle/gls_solver.py -- reproduce the record to the last bit."""
import json
import os

import pytest

from dealii_ns_gls_b200 import time_integration as ti
from oracle import gls_oracle as go
from oracle import gls_solver as gs
from oracle import ref_time_integration as rt

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def record():
    with open(os.path.join(HERE, "golden", "reference_time_integration.json")) as f:
        return json.load(f)


def test_object_code_of_the_reference_reproduces_the_record(record):
    if rt.load() is None:
        pytest.skip("oracle/_ref/libref_time_integration.so not built (no reference tree on this machine)")
    for case in record["bdf"]:
        t = rt.ReferenceTimeIntegrator(rt.ReferenceTimeIntegrator.BDF, order=case["order"])
        for dt, ref in zip(case["dts"], case["after_each_update"]):
            assert t.update_dt(dt) == ref["accepted"]
            assert t.query() == {k: v for k, v in ref.items() if k != "accepted"}
    for case in record["history"]:
        assert rt.history_after_commits(case["values"], case["commits"]) == case["after"]


@pytest.mark.parametrize("make", [lambda order: go.OracleBDF(order), lambda order: ti.TimeIntegratorDataBDF(order)],
                         ids=["oracle.OracleBDF", "product.TimeIntegratorDataBDF"])
def test_bdf_restatements_are_bit_equal_to_the_reference(make, record):
    n = 0
    for case in record["bdf"]:
        t = make(case["order"])
        for dt, ref in zip(case["dts"], case["after_each_update"]):
            assert ref["accepted"]
            t.update_dt(dt)
            weights = list(t.weights)
            assert weights == ref["weights"], (case["order"], case["dts"], weights, ref["weights"])
            if isinstance(t, ti.TimeIntegratorDataBDF):
                assert t.get_primary_weight() == ref["primary_weight"] and t.get_current_dt() == ref["current_dt"]
                assert t.get_order() == ref["order"] and t.get_theta() == ref["theta"]
            else:
                assert t.primary_weight == ref["primary_weight"] and t.current_dt == ref["current_dt"]
            n += 1
    assert n > 100
    # the weights the benchmarks use: performance.cc:44-46 (BDF2, one update_dt(0.1)) and two equal steps
    first = [c for c in record["bdf"] if c["order"] == 2 and c["dts"][:2] == [0.1, 0.1]][0]["after_each_update"]
    assert first[0]["weights"] == [10.0, -10.0, 0.0]
    assert first[1]["weights"][0] == 15.0 and abs(first[1]["weights"][1] + 20.0) < 1e-14


def test_theta_and_stationary_schemes(record):
    for case in record["theta"]:
        t = ti.TimeIntegratorDataTheta(case["theta"])
        for dt, ref in zip(case["dts"], case["after_each_update"]):
            t.update_dt(dt)
            assert list(t.get_weights()) == ref["weights"] and t.get_primary_weight() == ref["primary_weight"]
            assert t.get_current_dt() == ref["current_dt"] and t.get_theta() == ref["theta"] and t.get_order() == ref["order"]
    ref = record["none"]
    none = ti.TimeIntegratorDataNone()
    none.update_dt(0.3)
    assert (none.get_primary_weight(), none.get_current_dt(), none.get_theta(), none.get_order()) == \
        (ref["primary_weight"], ref["current_dt"], ref["theta"], ref["order"]) == (0.0, 1.0, 1.0, 0)
    assert list(none.get_weights()) == ref["weights"] == []
    assert gs._TimeNone.weights[0] == ref["primary_weight"]   # the oracle driver's stand-in: weight 0, dt 1 in step()


def test_solution_history_commit(record):
    torch = pytest.importorskip("torch")
    for case in record["history"]:
        h = ti.SolutionHistory(len(case["values"]))
        h.solutions = [torch.tensor([v]) for v in case["values"]]
        for _ in range(case["commits"]):
            h.commit_solution()
        assert [float(s[0]) for s in h.get_vectors()] == case["after"]


def test_cpp_host_mirror_is_bit_equal_to_the_reference(record):
    """glsb::TimeIntegratorDataBDF / Theta / None of dealii_ns_gls_b200/cpp/operator_b200.h (what a C++ host program
    on top of the C ABI reads its weights from), compiled for the host, on the record of the reference's own code"""
    import ctypes as C
    import subprocess
    path = os.path.join(HERE, "cpp", "libmirror_time_integration.so")
    if not os.path.exists(path):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-I/usr/local/cuda/include", "-o", path,
                               os.path.join(HERE, "cpp", "time_integration_host.cpp")])
    lib = C.CDLL(path)
    lib.mirror_create.restype = C.c_void_p
    lib.mirror_create.argtypes = [C.c_int, C.c_int, C.c_double]
    lib.mirror_destroy.argtypes = [C.c_void_p]
    lib.mirror_update_dt.restype = C.c_int
    lib.mirror_update_dt.argtypes = [C.c_void_p, C.c_double]
    lib.mirror_query.restype = C.c_int
    lib.mirror_query.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint)]

    def query(h):
        w = (C.c_double * 8)()
        pw, dt, th, order = C.c_double(), C.c_double(), C.c_double(), C.c_uint()
        n = lib.mirror_query(h, w, 8, C.byref(pw), C.byref(dt), C.byref(th), C.byref(order))
        return {"weights": [w[i] for i in range(n)], "primary_weight": pw.value, "current_dt": dt.value,
                "theta": th.value, "order": int(order.value)}

    for case in record["bdf"]:
        h = lib.mirror_create(0, case["order"], 1.0)
        for dt, ref in zip(case["dts"], case["after_each_update"]):
            assert lib.mirror_update_dt(h, dt) == 0
            assert query(h) == {k: v for k, v in ref.items() if k != "accepted"}
        lib.mirror_destroy(h)
    for case in record["theta"]:
        h = lib.mirror_create(1, 0, case["theta"])
        for dt, ref in zip(case["dts"], case["after_each_update"]):
            lib.mirror_update_dt(h, dt)
            assert query(h) == ref
        lib.mirror_destroy(h)
    h = lib.mirror_create(2, 0, 1.0)
    lib.mirror_update_dt(h, 0.3)
    assert query(h) == record["none"]
    lib.mirror_destroy(h)
