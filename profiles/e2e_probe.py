"""e2e probe: time glsb_vmult_host alone (bench.py's e2e leg) for the current env knobs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dealii_ns_gls_b200 import mesh as gm
from dealii_ns_gls_b200.operator import NavierStokesOperator
from dealii_ns_gls_b200.time_integration import TimeIntegratorDataBDF
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 160
order = sys.argv[2] if len(sys.argv) > 2 else "morton"
mesh = gm.hypercube(3, cells, 2, order=order)
ti = TimeIntegratorDataBDF(2); ti.update_dt(0.1)
op = NavierStokesOperator(mesh, None, 0.1, 4.0, 2.0, ti, False, True, True, number="double")
n = mesh.n_dofs
g = torch.Generator(device="cuda").manual_seed(1)
op.set_previous_solution([torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(3)])
op.set_linearization_point(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1)
hs = torch.rand(n, dtype=torch.float64).pin_memory(); hd = torch.empty(n, dtype=torch.float64).pin_memory()
for chunks in sys.argv[3:] or ["16"]:
    os.environ["GLSB_HOST_CHUNKS"] = chunks
    op._lib.glsb_destroy  # noqa
    # a fresh operator picks up the chunk count (the pipe is set up lazily on the first host call)
    op2 = NavierStokesOperator(mesh, None, 0.1, 4.0, 2.0, ti, False, True, True, number="double")
    op2.set_previous_solution([torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(3)])
    op2.set_linearization_point(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1)
    for _ in range(2):
        op2.vmult_host(hd, hs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        op2.vmult_host(hd, hs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"order={order} chunks={chunks} spec={'GLSB_HOST_NO_SPEC' not in os.environ}: {ms:.2f} ms/step = {n / ms / 1e6:.2f} GDoF/s", flush=True)
    del op2
# raw PCIe: one-way and both-way copies of the same size
d = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (("h2d", lambda: d.copy_(hs, non_blocking=True)), ("d2h", lambda: hd.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize()
    print(name, f"{n * 8 / (time.perf_counter() - t) / 1e9:.1f} GB/s")
d2 = torch.empty(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
with torch.cuda.stream(s1):
    d.copy_(hs, non_blocking=True)
with torch.cuda.stream(s2):
    hd.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
print(f"both directions at once: {n * 8 / (time.perf_counter() - t) / 1e9:.1f} GB/s each way")
