"""The periodicity argument of tests/full_size.py checked on the CPU: the oracle on a "big" block against the
oracle on the small block through the dof map (no GPU involved; the GPU test and bench.py's `parity_full_size`
key replace the big-block oracle by the CUDA path at sizes the oracle cannot run)."""
import numpy as np
import pytest
import torch

from dealii_ns_gls_b200 import mesh as gm
from tests.full_size import PeriodicFullSizeCheck, oracle_on_small
from tests.util import TI, make_oracle


@pytest.mark.parametrize("dim,degree,n,period,ctd,cell_wise", [
    (3, 2, 8, 2, False, True),     # performance.cc flags
    (3, 2, 12, 4, True, False),    # Turek-3D flags on Cartesian cells
    (3, 1, 8, 2, False, True),
    (2, 3, 12, 4, False, True),
    (2, 2, 9, 3, True, True),
])
def test_big_block_equals_mapped_small_block(dim, degree, n, period, ctd, cell_wise):
    big = gm.hypercube(dim, n, degree)
    chk = PeriodicFullSizeCheck(big, "cpu", period_cells=period)
    assert chk.small.n_cells == (3 * period) ** dim
    # the map sends every dof to one of the same component on a node with the same residues
    assert int(chk.key.max()) < chk.small.n_dofs
    weights, dt = ([15.0, -20.0, 5.0], 0.1) if ctd else ([10.0, -10.0, 0.0], 0.1)
    ti = TI(2, weights, dt)
    (lin_s, lin_b), (src_s, src_b) = chk.field(), chk.field()
    hist = [chk.field() for _ in range(3)] if ctd else None
    ref_small = oracle_on_small(chk, lin=lin_s, src=src_s, hist=None if hist is None else [h[0] for h in hist],
                                nu=0.1, c1=4.0, c2=2.0, weights=weights, dt=dt, ctd=ctd, cell_wise=cell_wise)
    ora = make_oracle(big, ti, ctd=ctd, cell_wise=cell_wise)
    if hist is not None:
        ora.set_previous_solution([h[1].numpy() for h in hist], weights)
    ora.set_linearization_point(lin_b.numpy(), dt)
    got = torch.from_numpy(ora.vmult(src_b.numpy(), weights[0]))
    r = chk.compare(got, ref_small)
    assert r["rel_l2_all_rows"] < 1e-13 and r["max_abs_over_max_ref"] < 1e-13, r
    # the check has teeth: one wrong cell contribution is seen
    got[int(big.cell_dofs[big.n_cells // 2, 5])] += 1e-6
    assert chk.compare(got, ref_small)["max_abs_over_max_ref"] > 1e-9


def test_bench_parity_full_size_object_with_the_oracle_standing_in_for_the_device():
    """bench.gpu_parity_full_size run end to end on the CPU: the operator handed to it is the oracle on the big
    block behind the device operator's method names"""
    import bench

    big = gm.hypercube(3, 12, 2)
    chk = PeriodicFullSizeCheck(big, "cpu", period_cells=4)
    ora = make_oracle(big, TI(2, [10.0, -10.0, 0.0], bench.DT), nu=bench.NU, c1=bench.C1, c2=bench.C2)

    class Op:
        def set_linearization_point(self, v):
            ora.set_linearization_point(v.numpy(), bench.DT)

        def vmult(self, dst, src):
            dst.copy_(torch.from_numpy(ora.vmult(src.numpy(), 10.0)))

        def vmult_variant(self):
            return "oracle"

    r = bench.gpu_parity_full_size(chk, Op(), torch.zeros(big.n_dofs, dtype=torch.float64), torch.float64, "double")
    assert r["ok"] and r["n_cells"] == 12 ** 3 and r["oracle_cells"] == 12 ** 3 and r["rel_l2_all_rows"] < 1e-13, r


def test_bench_time_step_cpu_baseline_runs():
    """bench.time_step_cpu (the CPU figure beside the device's wall time per time step) on the smallest hierarchy;
    the C-backed operator it uses is compared with the numpy oracle in tests/test_oracle.py"""
    import bench
    from dealii_ns_gls_b200.driver import ChannelParameters

    r = bench.time_step_cpu(0, n_steps=1)
    assert r["kind"] == "port" and r["warmup_steps"] == 2 and r["steps"] == 1 and r["levels"] == 2
    assert 0.0 < r["cell_loops_in_c_s"] < r["wall_s_per_step"]
    assert 1 <= r["newton_iterations"][0] <= 6 and all(1 <= k <= 20 for k in r["gmres_iterations"][0])
    p = ChannelParameters(dim=3, fe_degree=2, n_global_refinements=0, mg_min_level=1)
    assert r["n_dofs"] == p.level_mesh(p.n_levels()).n_dofs
