"""Wall time per time step of the channel driver (main.cc:908-990 mirror) with everything on the device:
usage  python profiles/step_probe.py <dim> <degree> <n_global_refinements> [n_steps] [mg_number] [graph|eager]
       [inexact|exact]"""
import json
import sys
import time

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dealii_ns_gls_b200.driver import ChannelParameters, Driver

dim, degree, r = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n_steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
mg_number = sys.argv[5] if len(sys.argv) > 5 else "float"
use_graph = len(sys.argv) > 6 and sys.argv[6] == "graph"
inexact = len(sys.argv) > 7 and sys.argv[7] == "inexact"
t0 = time.perf_counter()
from dealii_ns_gls_b200.multigrid import PreconditionerGMGAdditionalData
d = Driver(ChannelParameters(dim=dim, fe_degree=degree, n_global_refinements=r, mg_number=mg_number,
                             newton_inexact=inexact, gmg=PreconditionerGMGAdditionalData(use_cuda_graph=use_graph)))
torch.cuda.synchronize()
print(f"setup {time.perf_counter() - t0:.1f} s; fine level: {d.meshes[d.maxlevel].n_cells} cells, "
      f"{d.meshes[d.maxlevel].n_dofs} dofs, {d.maxlevel + 1} levels", flush=True)
out = []
for i in range(n_steps):
    d.timers = {} if i == n_steps - 1 else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rec = d.step()
    torch.cuda.synchronize()
    rec["wall_s"] = time.perf_counter() - t0
    rec["timers"] = d.timers
    rec["launches"] = sum(op.launch_count() for op in d.mg_ns_operators.values()) + d.ns_operator.launch_count()
    out.append(rec)
    print(json.dumps({k: v for k, v in rec.items() if k != "newton_residuals"}), flush=True)
