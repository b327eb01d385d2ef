"""The C++ host mirror of the operator interface (dealii_ns_gls_b200/cpp/operator_b200.h), driven by
tests/cpp/test_operator_b200.cpp on a golden fixture."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_operator_b200")


def test_cpp_host_mirror_is_built():
    if not os.path.exists(EXE):
        import __graft_entry__ as g
        g.build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_host_mirror_matches_golden():
    r = subprocess.run([EXE, os.path.join(ROOT, "tests", "golden", "turek_3d_q2_bdf2.bin")], capture_output=True,
                       text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr


def test_dealii_adapter_compiles_and_links():
    """cpp/dealii_adapter.h (the OperatorBase<Number> subclass a deal.II build uses) against the stub deal.II shim
    of tests/cpp/dealii_stub: every member instantiated for dim 2/3, double/float, every C-ABI / NCCL call
    resolved at link time.  Built by __graft_entry__.build()."""
    exe = os.path.join(ROOT, "tests", "cpp", "test_adapter_compiles")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "adapter instantiated" in r.stdout, r.stdout + r.stderr
    # no declared-but-undefined members left (round-1 finding): the header has no ';'-terminated private helper
    src = open(os.path.join(ROOT, "dealii_ns_gls_b200", "cpp", "dealii_adapter.h")).read()
    for name in ("update_ghost_values_start", "update_ghost_values_finish", "compress_start", "compress_finish",
                 "finish", "dev"):
        assert f" {name}(" in src
    undefined = subprocess.run(["nm", "-C", "--undefined-only", exe], capture_output=True, text=True).stdout
    assert "NavierStokesOperatorB200" not in undefined
