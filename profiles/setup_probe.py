"""Where the time of one Newton iteration goes (setup_preconditioner pieces, one V-cycle, one fine vmult):
usage  python profiles/setup_probe.py <dim> <degree> <n_global_refinements>"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dealii_ns_gls_b200.driver import ChannelParameters, Driver
from dealii_ns_gls_b200.multigrid import MGCoarseGridDirect, DeviceVectorOps
from dealii_ns_gls_b200.smoother import PreconditionRelaxation

dim, degree, r = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d = Driver(ChannelParameters(dim=dim, fe_degree=degree, n_global_refinements=r))
for _ in range(2):
    d.step()


def timed(name, fn, n=3):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{name:50s} {best * 1e3:9.3f} ms", flush=True)


sol = d.solution.get_current_solution()
mg = {}
timed("interpolate_to_mg", lambda: d.mg_transfer_no_constraints.interpolate_to_mg(mg, sol))
for l, op in d.mg_ns_operators.items():
    timed(f"level {l} ({op.mesh.n_cells} cells) set_linearization_point", lambda: op.set_linearization_point(mg[l]))
    diag = op.initialize_dof_vector()
    timed(f"level {l} compute_inverse_diagonal", lambda: op.compute_inverse_diagonal(diag))
    sm = PreconditionRelaxation(op, diag)
    timed(f"level {l} estimate_eigenvalues (20 power iterations)", lambda: sm.estimate_eigenvalues())
    x, y = op.initialize_dof_vector(), op.initialize_dof_vector()
    x.normal_()
    timed(f"level {l} vmult", lambda: op.vmult(y, x))
    timed(f"level {l} smoother.vmult (5 sweeps)", lambda: sm.vmult(y, x))
timed("coarse matrix + inverse", lambda: MGCoarseGridDirect(d.mg_ns_operators[0], DeviceVectorOps()))
timed("preconditioner.initialize()", lambda: d.preconditioner.initialize())
b = torch.randn_like(sol)
z = torch.zeros_like(sol)
timed("preconditioner.vmult (one V-cycle)", lambda: d.preconditioner.vmult(z, b))
timed("fine vmult (double)", lambda: d.ns_operator.vmult(z, b))
for l in range(d.maxlevel, 0, -1):
    t = d.transfer.transfers[l]
    f, c = d.mg_ns_operators[l].initialize_dof_vector(), d.mg_ns_operators[l - 1].initialize_dof_vector()
    timed(f"transfer {l} prolongate_and_add", lambda: t.prolongate_and_add(f, c))
    timed(f"transfer {l} restrict_and_add", lambda: t.restrict_and_add(c, f))
    if l < d.maxlevel - 1:
        break
