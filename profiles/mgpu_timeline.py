"""Timeline of the overlapped multi-GPU vmult (GhostExchange.vmult) with CUDA events; run under torchrun."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from dealii_ns_gls_b200 import mesh as gm, _lib as L
from dealii_ns_gls_b200.distributed import GhostExchange
from dealii_ns_gls_b200.operator import NavierStokesOperator
from dealii_ns_gls_b200.time_integration import TimeIntegratorDataBDF
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 160
mesh = gm.hypercube_slab(cells, 2, n_ranks=world, rank=rank)
ex = GhostExchange(mesh.partition, dev)
ti = TimeIntegratorDataBDF(2); ti.update_dt(0.1)
op = NavierStokesOperator(mesh, None, 0.1, 4.0, 2.0, ti, False, True, True, number="double", device=dev, exchange=ex)
n = mesh.n_dofs
g = torch.Generator(device=dev).manual_seed(rank)
op.set_previous_solution([torch.zeros(n, dtype=torch.float64, device=dev) for _ in range(3)])
op.set_linearization_point(torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 2 - 1)
src = torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 2 - 1; src[mesh.n_owned:] = 0
dst = op.initialize_dof_vector()
lib, h = op._lib, op._op
cur = torch.cuda.current_stream(dev); s = C.c_void_p(cur.cuda_stream)
d, x = C.c_void_p(dst.data_ptr()), C.c_void_p(src.data_ptr())
w = 10.0
def ev(): return torch.cuda.Event(enable_timing=True)
def run(reserve, comm=True):
    E = {k: ev() for k in ["t0", "A0", "A1", "B0", "B1", "I0", "I1", "c0", "c1", "d0", "d1", "end"]}
    lib.glsb_set_sm_reserve(h, reserve)
    E["t0"].record()
    lib.glsb_vmult_begin(h, d, s)
    ex.comm_stream.wait_stream(cur)
    with torch.cuda.stream(ex.comm_stream):
        E["c0"].record()
        if comm: ex.update_ghost_values(op, src)
        E["c1"].record()
    E["A0"].record()
    lib.glsb_vmult_cells_part(h, d, x, w, L.GLSB_CELLS_INTERIOR, 0, 2, s)
    E["A1"].record()
    cur.wait_stream(ex.comm_stream)
    E["B0"].record()
    lib.glsb_vmult_cells(h, d, x, w, L.GLSB_CELLS_BOUNDARY, s)
    E["B1"].record()
    ex.comm_stream.wait_stream(cur)
    buf = ex._buf("recv", dst.dtype)
    with torch.cuda.stream(ex.comm_stream):
        E["d0"].record()
        if comm:
            sends = [(r, dst[ex.n_owned + o: ex.n_owned + o + m]) for r, o, m in ex.part.recv]
            recvs = [(r, buf[o:o + m]) for r, o, m in ex.send_slices]
            ex._exchange(sends, recvs)
        E["d1"].record()
    E["I0"].record()
    lib.glsb_vmult_cells_part(h, d, x, w, L.GLSB_CELLS_INTERIOR, 1, 2, s)
    E["I1"].record()
    cur.wait_stream(ex.comm_stream)
    ex._unpack_add(op, dst, buf)
    lib.glsb_vmult_finish(h, d, x, s)
    E["end"].record()
    lib.glsb_set_sm_reserve(h, 0)
    return E
for reserve, comm in [(8, True), (8, True), (8, True), (0, True), (8, False), (0, False), (16, True), (32, True)]:
    for _ in range(2): run(reserve, comm)
    dist.barrier(); torch.cuda.synchronize()
    E = run(reserve, comm); torch.cuda.synchronize()
    t = lambda a, b: E[a].elapsed_time(E[b])
    print(f"rank {rank} reserve={reserve} comm={comm}: total {t('t0','end'):.2f} | A {t('A0','A1'):.2f} wait {t('A1','B0'):.2f} bnd {t('B0','B1'):.2f} B {t('I0','I1'):.2f} tail {t('I1','end'):.2f} | import {t('c0','c1'):.2f} (starts {t('t0','c0'):.2f}) compress {t('d0','d1'):.2f}", flush=True)
    dist.barrier()
dist.destroy_process_group()
