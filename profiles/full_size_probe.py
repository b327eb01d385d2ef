"""One short GPU call: the every-cell parity check (tests/full_size.py) on the bench workload itself -- 160^3 cells,
Q2, FP64, performance.cc flags -- through bench.gpu_parity_full_size, with the wall time of every setup stage and a
10-launch timing of the vmult.  Writes gpurun_out/full_size_probe.json after every stage (a cut-off call still
leaves what was reached).  Usage: python profiles/full_size_probe.py [cells_per_direction]"""
import json
import os
import sys
import time

T0 = time.perf_counter()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "full_size_probe.json")
out = {"stages_s": {}}


def mark(name):
    out["stages_s"][name] = round(time.perf_counter() - T0, 2)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(out, f)


def main():
    import torch

    import bench
    from dealii_ns_gls_b200 import mesh as gm
    from dealii_ns_gls_b200.operator import NavierStokesOperator
    from dealii_ns_gls_b200.time_integration import TimeIntegratorDataBDF
    from tests.full_size import PeriodicFullSizeCheck
    mark("imports")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    mesh = gm.hypercube(3, n, 2)
    out["native_meshgen"] = gm._native() is not None
    mark("mesh")
    ti = TimeIntegratorDataBDF(2)
    ti.update_dt(bench.DT)
    op = NavierStokesOperator(mesh, None, bench.NU, bench.C1, bench.C2, ti, False, True, True, number="double",
                              device=dev)
    torch.cuda.synchronize()
    mark("create")
    chk = PeriodicFullSizeCheck(mesh, dev, period_cells=4)
    torch.cuda.synchronize()
    mark("dof_map")
    n_dofs, n_cells = mesh.n_dofs, mesh.n_cells
    del mesh
    op.set_previous_solution([torch.zeros(n_dofs, dtype=torch.float64, device=dev) for _ in range(3)])
    dst = op.initialize_dof_vector()
    out["parity_full_size"] = bench.gpu_parity_full_size(chk, op, dst, torch.float64, "double")
    torch.cuda.synchronize()
    mark("parity")
    # the bench's own protocol on this build: random vectors, 3 warm-ups, 10 launches between events
    g = torch.Generator(device=dev).manual_seed(bench.SEED)
    rv = lambda: torch.rand(n_dofs, dtype=torch.float64, device=dev, generator=g) * 2 - 1  # noqa: E731
    op.set_linearization_point(rv())
    src = rv()
    peak = 6455.6
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = json.load(f).get("hbm_gbs", peak)
    except Exception:
        pass
    out["vmult"] = bench.side_measure(op, src, dst, 10, n_cells, n_dofs, bench.algorithmic_bytes_per_cell(3, 2, 8), peak)
    mark("timing")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
