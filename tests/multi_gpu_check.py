"""Multi-GPU parity check, launched by hand on the GPU box (not collected by pytest):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tests/multi_gpu_check.py

Every rank builds its part (Morton box by default, GLSB_CHECK_PARTITION=slab for z-slabs) of a small hypercube, applies the CUDA operator with the NCCL ghost
exchange (interior cells overlapped with the import, compress(add) afterwards) and the result is
compared with the CPU oracle evaluated on the union of the slabs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dealii_ns_gls_b200 import mesh as gm  # noqa: E402
from dealii_ns_gls_b200.distributed import GhostExchange  # noqa: E402
from dealii_ns_gls_b200.operator import NavierStokesOperator  # noqa: E402
from tests.util import TI, make_oracle  # noqa: E402


def make_part(*a, **kw):
    """GLSB_CHECK_PARTITION = box (default: Morton halves / quarters / octants, the reference's p4est owner
    ranks, up to 7 neighbours) or slab (z-slabs, the round-1 partition)"""
    fn = gm.hypercube_slab if os.environ.get("GLSB_CHECK_PARTITION", "box") == "slab" else gm.hypercube_box
    return fn(*a, **kw)


def field(ids, seed):
    x = (ids.astype(np.float64) * 0.6180339887498949 + seed * 0.137) % 1.0
    return 2.0 * x - 1.0


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    n, degree, w = 7, 2, 10.0
    m = make_part(n, degree, n_ranks=world, rank=rank)
    ex = GhostExchange(m.partition, dev)
    ti = TI(2, [w, -w, 0.0], 0.1)
    op = NavierStokesOperator(m, None, 0.1, 4.0, 2.0, ti, False, True, True, number="double", device=dev, exchange=ex)
    lin = torch.tensor(field(m.canonical_ids, 1), device=dev)
    src = torch.tensor(field(m.canonical_ids, 2), device=dev)
    lin[m.n_owned:] = 0
    src[m.n_owned:] = 0
    op.set_linearization_point(lin)
    dst = op.initialize_dof_vector()
    op.vmult(dst, src)
    diag = op.initialize_dof_vector()
    op.compute_inverse_diagonal(diag)
    umax = op.get_max_u(src)
    torch.cuda.synchronize()
    # reference on the union of the slabs (every rank computes it; sizes are tiny)
    meshes = [make_part(n, degree, n_ranks=world, rank=r) for r in range(world)]
    ng = meshes[0].n_global_dofs
    acc, dacc, um = np.zeros(ng), np.zeros(ng), 0.0
    for mm in meshes:
        o = make_oracle(mm, ti)
        o.set_linearization_point(field(mm.canonical_ids, 1), 0.1)
        np.add.at(acc, mm.canonical_ids, o._scatter(o._apply_cells(o._gather(field(mm.canonical_ids, 2)), w, False)))
        A = o.cell_matrices(w)
        dl = np.zeros(mm.n_dofs)
        np.add.at(dl, mm.cell_dofs.reshape(-1).astype(np.int64), np.einsum("kii->ki", A).reshape(-1))
        np.add.at(dacc, mm.canonical_ids, dl)
        um = max(um, o.get_max_u(field(mm.canonical_ids, 2)))
    ids = m.canonical_ids[: m.n_owned]
    e1 = np.linalg.norm(dst[: m.n_owned].cpu().numpy() - acc[ids]) / np.linalg.norm(acc)
    dref = np.where(np.abs(dacc) > 1e-10, 1.0 / dacc, 1.0)
    e2 = np.linalg.norm(diag[: m.n_owned].cpu().numpy() - dref[ids]) / np.linalg.norm(dref)
    e3 = abs(umax - um)
    ok = e1 < 1e-12 and e2 < 1e-11 and e3 < 1e-12 and op.vmult_variant() == "q2_regtile_tma"
    print(f"rank {rank}/{world}: vmult rel_l2 {e1:.2e}  inv_diag rel_l2 {e2:.2e}  max_u err {e3:.1e}  "
          f"interior/boundary cells {m.n_cells - int(m.cell_is_boundary.sum())}/{int(m.cell_is_boundary.sum())} "
          f"neighbours recv {[r for r, _, _ in m.partition.recv]} send {[r for r, _ in m.partition.send]} "
          f"variant {op.vmult_variant()}  {'OK' if ok else 'FAIL'}", flush=True)
    # relaxation smoother on the partitioned operator: omega from the distributed power iteration and 5 sweeps,
    # against the CPU restatement on the union of the slabs in the SAME global numbering (owner offset + local)
    from dealii_ns_gls_b200.smoother import PreconditionRelaxation
    from oracle import gls_oracle as go
    from oracle.gls_smoother import OracleRelaxation
    sm = PreconditionRelaxation(op, diag)
    sm.estimate_eigenvalues()
    bvec = torch.tensor(field(m.canonical_ids, 3), device=dev)
    bvec[m.n_owned:] = 0
    xs = op.initialize_dof_vector()
    sm.vmult(xs, bvec)
    canon_to_global = np.zeros(ng, dtype=np.int64)
    for mm in meshes:
        canon_to_global[mm.canonical_ids[: mm.n_owned]] = mm.partition.owned_offset + np.arange(mm.n_owned)
    gl = [canon_to_global[mm.canonical_ids] for mm in meshes]
    union_dofs = np.concatenate([g[mm.cell_dofs.astype(np.int64)] for g, mm in zip(gl, meshes)])
    union_pts = np.concatenate([mm.cell_points for mm in meshes])
    ou = go.OracleOperator(dim=3, degree=degree, cell_dofs=union_dofs, n_dofs=ng, cell_points=union_pts,
                           mapping_degree=1, constraints={}, nu=0.1, c1=4.0, c2=2.0, theta=1.0, order=2,
                           consider_time_derivative=False, increment_form=True, cell_wise_stabilization=True,
                           path="sumfac")
    to_global = np.zeros(ng)
    for g, mm in zip(gl, meshes):
        to_global[g] = field(mm.canonical_ids, 1)
    ou.set_linearization_point(to_global, 0.1)
    osm = OracleRelaxation(ou, w, ou.compute_inverse_diagonal(w))
    osm.estimate_eigenvalues()
    bg = np.zeros(ng)
    for g, mm in zip(gl, meshes):
        bg[g] = field(mm.canonical_ids, 3)
    xref = osm.vmult(bg)
    e5 = abs(sm.get_relaxation() / osm.get_relaxation() - 1.0)
    e6 = np.linalg.norm(xs[: m.n_owned].cpu().numpy() - xref[gl[rank][: m.n_owned]]) / np.linalg.norm(xref)
    ok56 = e5 < 1e-10 and e6 < 1e-10
    print(f"rank {rank}/{world}: relaxation omega rel err {e5:.1e}  5 sweeps rel_l2 {e6:.1e}  {'OK' if ok56 else 'FAIL'}",
          flush=True)
    ok = ok and ok56
    # host-vector entry of a partitioned operator: chunked pipeline + exchange in between (glsb_vmult_host_begin /
    # _finish) against the device-vector vmult, on a mesh large enough for several chunks and with Dirichlet rows
    n2 = int(os.environ.get("GLSB_CHECK_CELLS", "48"))
    eps = 1e-12
    m2 = make_part(n2, degree, n_ranks=world, rank=rank, with_points=False)
    ex2 = GhostExchange(m2.partition, dev)
    op2 = NavierStokesOperator(m2, None, 0.1, 4.0, 2.0, ti, False, True, True, number="double", device=dev, exchange=ex2)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    lin2 = torch.rand(m2.n_dofs, dtype=torch.float64, device=dev, generator=g) * 2 - 1
    src2 = torch.rand(m2.n_dofs, dtype=torch.float64, device=dev, generator=g) * 2 - 1
    lin2[m2.n_owned:] = 0
    src2[m2.n_owned:] = 0
    op2.set_linearization_point(lin2)
    ref2 = op2.initialize_dof_vector()
    op2.vmult(ref2, src2)
    h_src = torch.empty(m2.n_dofs, dtype=torch.float64, pin_memory=True)
    h_dst = torch.full((m2.n_dofs,), 7.0, dtype=torch.float64).pin_memory()
    h_src.copy_(src2)
    for _ in range(2):
        op2.vmult_host(h_dst, h_src)
    torch.cuda.synchronize()
    e4 = float((h_dst[: m2.n_owned] - ref2[: m2.n_owned].cpu()).abs().max() / ref2.abs().max())
    ok4 = e4 < 1e-13
    print(f"rank {rank}/{world}: vmult_host (pipelined, {m2.n_cells} cells) max err {e4:.2e}  {'OK' if ok4 else 'FAIL'}",
          flush=True)
    ok = ok and ok4
    # config C partitioned (mesh.cylinder_shell_box: curved Q2 cells, periodic direction, no-slip rows, Turek-3D
    # flags): vmult on the N parts against the oracle's cell loops summed over the union by canonical ids
    ti3 = TI(2, [15.0, -20.0, 5.0], 0.1)
    kw3 = dict(nu=0.001, ctd=True, cell_wise=False)
    shape3 = (2, 6, 2)
    m3 = gm.cylinder_shell_box(shape3, degree, n_ranks=world, rank=rank)
    ex3 = GhostExchange(m3.partition, dev)
    op3 = NavierStokesOperator(m3, None, 0.001, 4.0, 2.0, ti3, True, True, False, number="double", device=dev, exchange=ex3)
    vec3 = lambda seed: torch.tensor(field(m3.canonical_ids, seed), device=dev)  # noqa: E731
    hist3 = [vec3(5), vec3(6), vec3(7)]
    lin3, src3 = vec3(1), vec3(2)
    for v in hist3 + [lin3, src3]:
        v[m3.n_owned:] = 0
    op3.set_previous_solution(hist3)
    op3.set_linearization_point(lin3)
    dst3 = op3.initialize_dof_vector()
    op3.vmult(dst3, src3)
    torch.cuda.synchronize()
    parts = [gm.cylinder_shell_box(shape3, degree, n_ranks=world, rank=r) for r in range(world)]
    ng3 = parts[0].n_global_dofs
    acc3, cons3 = np.zeros(ng3), np.zeros(ng3, dtype=bool)
    for mm in parts:
        o = make_oracle(mm, ti3, **kw3)
        o.set_previous_solution([field(mm.canonical_ids, s_) for s_ in (5, 6, 7)], ti3.get_weights())
        o.set_linearization_point(field(mm.canonical_ids, 1), 0.1)
        x = field(mm.canonical_ids, 2).copy()
        c = np.array(sorted(mm.constraints), dtype=np.int64)
        x[c] = 0.0
        cons3[mm.canonical_ids[c]] = True
        np.add.at(acc3, mm.canonical_ids, o._scatter(o._apply_cells(o._gather(x), 15.0, False)))
    ids3 = m3.canonical_ids[: m3.n_owned]
    got3 = dst3[: m3.n_owned].cpu().numpy()
    free3 = ~cons3[ids3]
    e7 = np.linalg.norm((got3 - acc3[ids3])[free3]) / np.linalg.norm(acc3[~cons3])
    ident3 = bool(np.array_equal(got3[~free3], field(ids3, 2)[~free3]))
    ok7 = e7 < 1e-12 and ident3
    print(f"rank {rank}/{world}: config C (O-grid, {m3.n_cells} curved cells per rank, {int((~free3).sum())} no-slip rows) "
          f"vmult rel_l2 on free rows {e7:.2e}  identity rows bit-equal {ident3}  {'OK' if ok7 else 'FAIL'}", flush=True)
    ok = ok and ok7
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(int(t.item()) != 0)


if __name__ == "__main__":
    main()
