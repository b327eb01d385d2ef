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
