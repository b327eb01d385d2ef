"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol
include/glsb200.h declares, and fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re

import pytest

from dealii_ns_gls_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()


def test_header_symbols_are_exported_and_bound():
    _ensure_built()
    hdr = open(os.path.join(ROOT, "include", "glsb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(glsb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(L.SYMBOLS), (declared ^ set(L.SYMBOLS))
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name


def test_desc_struct_matches_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "glsb200.h")).read()
    body = hdr[hdr.index("typedef struct glsb_desc"):hdr.index("} glsb_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = decl.split("{")[-1]
        for part in decl.split(","):
            m = re.search(r"\*?\s*([A-Za-z_0-9]+)\s*$", part.strip())
            if m:
                names.append(m.group(1))
    assert names == [f[0] for f in L.GlsbDesc._fields_]


def test_create_fails_loudly_without_device_or_with_bad_desc():
    _ensure_built()
    lib = L.load()
    d = L.GlsbDesc()
    h = C.c_void_p()
    d.abi_version = 99
    assert lib.glsb_create(C.byref(d), C.byref(h)) != 0
    assert b"ABI" in lib.glsb_last_error(None)
    d.abi_version = L.GLSB_ABI_VERSION
    d.dim = 5
    assert lib.glsb_create(C.byref(d), C.byref(h)) != 0
    assert b"dim" in lib.glsb_last_error(None)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        d.dim, d.degree, d.n_cells, d.n_owned = 2, 1, 1, 12
        assert lib.glsb_create(C.byref(d), C.byref(h)) != 0
        assert b"no CUDA device" in lib.glsb_last_error(None)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            from dealii_ns_gls_b200.operator import NavierStokesOperator
            NavierStokesOperator(None, None, 0.1, 4, 2, None, False, True, True)


def test_time_integrator_mirror_matches_oracle():
    from dealii_ns_gls_b200.time_integration import TimeIntegratorDataBDF, TimeIntegratorDataNone
    from oracle.gls_oracle import OracleBDF
    for order in (1, 2, 3):
        a, b = TimeIntegratorDataBDF(order), OracleBDF(order)
        for dt in (0.1, 0.07, 0.2, 0.05):
            a.update_dt(dt)
            b.update_dt(dt)
            assert a.get_weights() == b.weights
            assert a.get_current_dt() == b.current_dt
    n = TimeIntegratorDataNone()
    assert (n.get_order(), n.get_primary_weight(), n.get_current_dt(), n.get_theta()) == (0, 0.0, 1.0, 1.0)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU restatement on the host cores) runs without a GPU and prints
    exactly one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--cpu-cells", "8",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GDoF/s" and d["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "config",
                "cpu_baseline", "e2e"):
        assert key in d
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_transfer_and_vector_entry_points_fail_loudly():
    """the multigrid-transfer and Krylov entry points reject bad descriptors / null pointers with a non-zero
    status and never fall back to the CPU"""
    lib = L.load()
    d = L.GlsbTransferDesc()
    d.abi_version = 99
    h = C.c_void_p()
    assert lib.glsb_transfer_create(C.byref(d), C.byref(h)) != 0
    assert b"abi" in lib.glsb_transfer_last_error(None).lower()
    d.abi_version = L.GLSB_ABI_VERSION
    d.dim, d.degree, d.number_type = 4, 2, L.GLSB_F32
    assert lib.glsb_transfer_create(C.byref(d), C.byref(h)) != 0
    assert b"dim" in lib.glsb_transfer_last_error(None)
    d.dim = 3
    import torch
    if not torch.cuda.is_available():
        assert lib.glsb_transfer_create(C.byref(d), C.byref(h)) != 0
        assert b"no usable CUDA device" in lib.glsb_transfer_last_error(None)
    assert lib.glsb_transfer_prolongate_and_add(None, None, None, None) != 0
    assert lib.glsb_vec_multi_dot(None, None, 0, 1, None, 0, L.GLSB_F64, None) != 0
    assert lib.glsb_vec_multi_axpy(None, None, 0, 1, None, 1.0, 0, L.GLSB_F64, None) != 0
    assert lib.glsb_vec_axpby(None, 1.0, None, 0.0, 0, L.GLSB_F64, None) != 0
    assert lib.glsb_vec_convert(None, L.GLSB_F64, None, L.GLSB_F32, 0, None) != 0
    assert lib.glsb_dense_apply(None, None, None, 1, 1, L.GLSB_F64, None) != 0
    assert lib.glsb_get_system_matrix(None, None, 0.0, None) != 0
    assert lib.glsb_vmult_host_begin(None, None, None, None, None, 0.0, None) != 0
