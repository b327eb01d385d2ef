import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dealii_ns_gls_b200 import mesh as gm
from dealii_ns_gls_b200.distributed import GhostExchange
from tests.util import TI, make_gpu

class LoopbackExchange(GhostExchange):
    def update_ghost_values(self, op, vec):
        vec[self.n_owned:] = self.ghost_values.to(vec.dtype)
    def compress_add(self, op, vec):
        vec[self.n_owned:] = 0

for (N, rank, n) in ((4, 3, 24), (4, 3, 8), (2, 1, 9)):
    mesh = gm.hypercube_box(n, 2, n_ranks=N, rank=rank, with_points=False)
    ex = LoopbackExchange(mesh.partition, torch.device("cuda", 0))
    g = torch.Generator(device="cuda").manual_seed(3 + rank)
    ex.ghost_values = torch.rand(mesh.n_dofs - mesh.n_owned, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    ti = TI(2, [10.0, -10.0, 0.0], 0.1)
    gpu = make_gpu(mesh, ti, exchange=ex)
    lin = torch.rand(mesh.n_dofs, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    src = torch.rand(mesh.n_dofs, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    lin[mesh.n_owned:] = 0; src[mesh.n_owned:] = 0
    gpu.set_linearization_point(lin)
    ref = gpu.initialize_dof_vector(); gpu.vmult(ref, src)
    h_src = torch.empty(mesh.n_dofs, dtype=torch.float64, pin_memory=True)
    h_dst = torch.full((mesh.n_dofs,), float("nan"), dtype=torch.float64).pin_memory()
    h_src.copy_(src)
    for _ in range(2):
        gpu.vmult_host(h_dst, h_src)
    torch.cuda.synchronize()
    no = mesh.n_owned
    d = (h_dst[:no] - ref[:no].cpu()).abs().numpy()
    bad = np.nonzero(d > 1e-13 * float(ref.abs().max()))[0]
    cd = mesh.cell_dofs.astype(np.int64)
    isb = mesh.cell_is_boundary
    touched_b = np.zeros(mesh.n_dofs, bool); touched_b[cd[isb].reshape(-1)] = True
    touched_i = np.zeros(mesh.n_dofs, bool); touched_i[cd[~isb].reshape(-1)] = True
    n_int = int((~isb).sum())
    # interior cells of the last (partial) interior batch, in the library's order = mesh order restricted to interior
    int_cells = np.nonzero(~isb)[0]
    last_batch = int_cells[(n_int // 32) * 32:]
    tl = np.zeros(mesh.n_dofs, bool); tl[cd[last_batch].reshape(-1)] = True
    print(f"N={N} rank={rank} n={n}: n_interior={n_int} (%32={n_int % 32}) bad={len(bad)} of {no}; bad touched by boundary cells: "
          f"{int(touched_b[bad].sum())}, by interior: {int(touched_i[bad].sum())}, by last partial interior batch: {int(tl[bad].sum())}; "
          f"dofs touched by last partial batch: {int(tl[:no].sum())}")
    if len(bad):
        # is the host value equal to ref minus the contribution of some set? check ratio
        r = (h_dst[:no].numpy()[bad] / ref[:no].cpu().numpy()[bad])
        print("   ratio host/ref quantiles", np.quantile(r, [0, .25, .5, .75, 1]))
