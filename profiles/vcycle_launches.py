"""Kernel launches of ONE V-cycle (PreconditionerGMG::vmult) for an ncu launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
      --nvtx --nvtx-include "vcycle/" python profiles/vcycle_launches.py 3 2 3
(the V-cycle is bracketed by cudaProfilerStart/Stop: run ncu with --profile-from-start off)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dealii_ns_gls_b200.driver import ChannelParameters, Driver

dim, degree, r = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d = Driver(ChannelParameters(dim=dim, fe_degree=degree, n_global_refinements=r))
d.step()
sol = d.solution.get_current_solution()
b, z = torch.randn_like(sol), torch.zeros_like(sol)
d.preconditioner.vmult(z, b)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
d.preconditioner.vmult(z, b)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
