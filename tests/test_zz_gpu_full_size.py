"""Parity at BASELINE.json's full size (performance.cc recipe, 160^3 cells = 1.32e8 DoFs per GPU) and at sizes in
between: every cell of the CUDA result against the CPU oracle through periodicity (tests/full_size.py).  The file
sorts last on purpose: the full-size case builds a 20 GB operator."""
import pytest

from dealii_ns_gls_b200 import mesh as gm
from tests.full_size import PeriodicFullSizeCheck, oracle_on_small
from tests.util import TI, make_gpu

pytestmark = pytest.mark.gpu

TOL = {"double": 1e-12, "float": 2e-5}


def _run(n, degree, number, ctd, cell_wise, period=4):
    import gc

    import torch
    gc.collect()
    torch.cuda.empty_cache()      # the 160^3 case wants ~30 GB next to whatever earlier tests left in torch's cache
    tdt = torch.float64 if number == "double" else torch.float32
    big = gm.hypercube(3, n, degree)
    chk = PeriodicFullSizeCheck(big, "cuda", period_cells=period)
    weights, dt = ([15.0, -20.0, 5.0], 0.1) if ctd else ([10.0, -10.0, 0.0], 0.1)
    gpu = make_gpu(big, TI(2, weights, dt), ctd=ctd, cell_wise=cell_wise, number=number)
    variant = gpu.vmult_variant()
    del big
    (lin_s, lin_b), (src_s, src_b) = chk.field(), chk.field()
    hist = [chk.field() for _ in range(3)]
    gpu.set_previous_solution([h[1].to(tdt) for h in hist])
    gpu.set_linearization_point(lin_b.to(tdt))
    dst = gpu.initialize_dof_vector()
    gpu.vmult(dst, src_b.to(tdt))
    ref = oracle_on_small(chk, lin=lin_s, src=src_s, hist=[h[0] for h in hist], nu=0.1, c1=4.0, c2=2.0,
                          weights=weights, dt=dt, ctd=ctd, cell_wise=cell_wise)
    r = chk.compare(dst, ref)
    r["kernel_variant"] = variant
    return r


@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("degree,n,ctd,cell_wise", [(2, 48, False, True), (2, 48, True, False), (1, 64, False, True),
                                                    (3, 32, False, True)])
def test_every_cell_against_the_oracle_medium(degree, n, ctd, cell_wise, number):
    """1.1e5 ... 2.6e5 cells: some ten batches per persistent CTA, all cells compared"""
    r = _run(n, degree, number, ctd, cell_wise)
    assert r["rel_l2_all_rows"] < TOL[number] and r["max_abs_over_max_ref"] < 10 * TOL[number], r


def test_every_cell_against_the_oracle_at_the_bench_size():
    """160^3 cells, Q2, FP64, performance.cc flags: the bench workload itself (2.2e9 table elements, 17.7 GB of
    tables, 128 000 batches over 296 persistent CTAs)"""
    r = _run(160, 2, "double", False, True)
    assert r["kernel_variant"] == "q2_regtile_tma"
    assert r["n_cells"] == 160 ** 3
    assert r["rel_l2_all_rows"] < 1e-12 and r["max_abs_over_max_ref"] < 1e-11, r
