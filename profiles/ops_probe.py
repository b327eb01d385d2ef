"""Times the non-vmult entry points (a6-a11 of SURVEY.md section 8) on config P / config C flags."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dealii_ns_gls_b200 import mesh as gm
from dealii_ns_gls_b200.operator import NavierStokesOperator
from dealii_ns_gls_b200.time_integration import TimeIntegratorDataBDF
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 96
for number in ("double", "float"):
    dt = torch.float64 if number == "double" else torch.float32
    mesh = gm.hypercube(3, cells, 2)
    ti = TimeIntegratorDataBDF(2); ti.update_dt(0.1); ti.update_dt(0.1)
    op = NavierStokesOperator(mesh, None, 0.1, 4.0, 2.0, ti, True, True, False, number=number)
    n = mesh.n_dofs
    g = torch.Generator(device="cuda").manual_seed(1)
    rv = lambda: (torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1).to(dt)
    hist = [rv() for _ in range(3)]; lin = rv(); src = rv(); dst = op.initialize_dof_vector(); diag = op.initialize_dof_vector()
    op.set_previous_solution(hist); op.set_linearization_point(lin)
    def timed(f, reps=3):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    res = {
        "vmult": timed(lambda: op.vmult(dst, src), 10),
        "set_linearization_point": timed(lambda: op.set_linearization_point(lin)),
        "set_previous_solution": timed(lambda: op.set_previous_solution(hist)),
        "evaluate_residual": timed(lambda: op.evaluate_residual(dst, src)),
        "compute_inverse_diagonal": timed(lambda: op.compute_inverse_diagonal(diag), 2),
        "get_max_u": timed(lambda: op.get_max_u(lin)),
    }
    print(f"{number} Q2 {cells}^3 cells ({n} DoFs), ctd + q-wise delta, Cartesian: " +
          "  ".join(f"{k} {v:.3f} ms ({n / v / 1e6:.2f} GDoF/s)" for k, v in res.items()), flush=True)
    del op
