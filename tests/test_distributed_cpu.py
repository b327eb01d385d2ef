"""world_size-2 gloo tests of the N > 1 host path (partition lists + ghost exchange protocol),
run on CPU with the oracle as each rank's local cell loop."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _slab_reference(n, degree, R, lin_fn, src_fn, weight, kind="slab"):
    from dealii_ns_gls_b200 import mesh as gm
    from tests.util import TI, make_oracle
    full = gm.hypercube_slab(n, degree, n_ranks=1, rank=0)
    # one rank holding the whole R-slab domain: stack the slabs by canonical ids
    meshes = [_make(gm, kind, n, degree, R, r) for r in range(R)]
    ng = meshes[0].n_global_dofs
    ti = TI(2, [weight, -weight, 0.0], 0.1)
    acc = np.zeros(ng)
    for m in meshes:
        o = make_oracle(m, ti)
        o.set_linearization_point(lin_fn(m.canonical_ids), 0.1)
        loc = o._scatter(o._apply_cells(o._gather(src_fn(m.canonical_ids)), weight, False))
        np.add.at(acc, m.canonical_ids, loc)
    del full
    return acc


def _field(ids, seed):
    # deterministic pseudo-random value per canonical dof id, identical on every rank
    x = (ids.astype(np.float64) * 0.6180339887498949 + seed * 0.137) % 1.0
    return 2.0 * x - 1.0


def _make(gm, kind, n, degree, world, rank):
    return (gm.hypercube_box if kind == "box" else gm.hypercube_slab)(n, degree, n_ranks=world, rank=rank)


def _worker(rank, world, port, n, degree, out, kind="slab"):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dealii_ns_gls_b200 import mesh as gm
    from dealii_ns_gls_b200.distributed import GhostExchange
    from tests.util import TI, make_oracle
    m = _make(gm, kind, n, degree, world, rank)
    ex = GhostExchange(m.partition, "cpu")
    ti = TI(2, [10.0, -10.0, 0.0], 0.1)
    o = make_oracle(m, ti)
    n_owned = m.n_owned
    # vectors arrive without ghost values
    lin = torch.tensor(_field(m.canonical_ids, 1))
    src = torch.tensor(_field(m.canonical_ids, 2))
    lin[n_owned:] = 0
    src[n_owned:] = 0
    ex.update_ghost_values(None, lin)
    o.set_linearization_point(lin.numpy(), 0.1)
    ex.update_ghost_values(None, src)
    dst = torch.tensor(o._scatter(o._apply_cells(o._gather(src.numpy()), 10.0, False)))
    ex.compress_add(None, dst)
    mx = ex.allreduce_max(float(rank + 1))
    torch.save({"ids": m.canonical_ids[:n_owned], "dst": dst[:n_owned].numpy(), "mx": mx,
                "ghost_zero": bool((dst[n_owned:] == 0).all())}, out + f".{rank}")
    dist.destroy_process_group()


@pytest.mark.parametrize("degree", [1, 2])
def test_two_rank_ghost_exchange_matches_single_domain(tmp_path, degree):
    n, world = 3, 2
    port = 29500 + (os.getpid() % 2000) + degree
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, port, n, degree, out), nprocs=world, join=True)
    ref = _slab_reference(n, degree, world, lambda ids: _field(ids, 1), lambda ids: _field(ids, 2), 10.0)
    seen = np.zeros(len(ref), dtype=bool)
    for r in range(world):
        d = torch.load(out + f".{r}", weights_only=False)
        assert d["ghost_zero"] and d["mx"] == float(world)
        assert np.linalg.norm(d["dst"] - ref[d["ids"]]) <= 1e-13 * np.linalg.norm(ref)
        assert not seen[d["ids"]].any()          # every dof owned exactly once
        seen[d["ids"]] = True
    assert seen.all()


def test_slab_partition_lists_are_consistent():
    from dealii_ns_gls_b200 import mesh as gm
    R, n, p = 3, 2, 2
    ms = [gm.hypercube_slab(n, p, n_ranks=R, rank=r) for r in range(R)]
    assert sum(m.n_owned for m in ms) == ms[0].n_global_dofs
    for r in range(R - 1):
        (to, exp), = ms[r].partition.send
        (frm, off, cnt), = ms[r + 1].partition.recv
        assert to == r + 1 and frm == r and cnt == len(exp) and off == 0
        # the sender's export order equals the receiver's ghost order (same canonical ids)
        assert np.array_equal(ms[r].canonical_ids[exp], ms[r + 1].canonical_ids[ms[r + 1].n_owned:])
    assert not ms[0].partition.recv and not ms[-1].partition.send


@pytest.mark.parametrize("R", [2, 4, 8])
def test_morton_box_partition_lists_are_consistent(R):
    """p4est-style partition (halves / quarters / octants of the Morton curve, performance.cc:29-31): every
    dof owned once by the lowest touching rank, contiguous owned ranges, and the sender's export order equal
    to the receiver's ghost order for every pair, including edge- and corner-neighbours."""
    from dealii_ns_gls_b200 import mesh as gm
    n, p = 2, 2
    ms = [gm.hypercube_box(n, p, n_ranks=R, rank=r) for r in range(R)]
    ng = ms[0].n_global_dofs
    assert sum(m.n_owned for m in ms) == ng
    owner = np.full(ng, -1)
    off = 0
    for r, m in enumerate(ms):
        assert m.partition.owned_offset == off
        off += m.n_owned
        assert (owner[m.canonical_ids[:m.n_owned]] == -1).all()
        owner[m.canonical_ids[:m.n_owned]] = r
    assert (owner >= 0).all()
    for r, m in enumerate(ms):
        # lowest touching rank owns
        touch = m.canonical_ids
        assert (owner[touch] <= r).all()
        assert (m.partition.ghost_owner == owner[m.canonical_ids[m.n_owned:]]).all()
        for frm, o, cnt in m.partition.recv:
            (exp,) = [idx for to, idx in ms[frm].partition.send if to == r]
            assert cnt == len(exp)
            assert np.array_equal(ms[frm].canonical_ids[exp], m.canonical_ids[m.n_owned + o:m.n_owned + o + cnt])
        assert sum(c for _, _, c in m.partition.recv) == m.n_dofs - m.n_owned
    if R == 8:
        assert len(ms[7].partition.recv) == 7 and len(ms[0].partition.send) == 7  # faces, edges and the corner


def test_four_rank_quarters_ghost_exchange_matches_single_domain(tmp_path):
    """(z, y)-quarters: ranks with 1 and 3 neighbours, edge-shared dofs contributed to by all four ranks."""
    n, world, degree = 2, 4, 2
    port = 31500 + (os.getpid() % 2000)
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, port, n, degree, out, "box"), nprocs=world, join=True)
    ref = _slab_reference(n, degree, world, lambda ids: _field(ids, 1), lambda ids: _field(ids, 2), 10.0, kind="box")
    seen = np.zeros(len(ref), dtype=bool)
    for r in range(world):
        d = torch.load(out + f".{r}", weights_only=False)
        assert d["ghost_zero"] and d["mx"] == float(world)
        assert np.linalg.norm(d["dst"] - ref[d["ids"]]) <= 1e-13 * np.linalg.norm(ref)
        seen[d["ids"]] = True
    assert seen.all()


def test_partitioned_cylinder_shell_equals_single_rank_shell():
    """config C on N > 1 ranks (mesh.cylinder_shell_box: curved cells, periodic direction, no-slip rows): the cell
    loops of the 4 parts, summed over shared dofs, equal the single-rank O-grid's vmult on the unconstrained rows."""
    from dealii_ns_gls_b200 import mesh as gm
    from tests.util import TI, make_oracle
    ti = TI(2, [15.0, -20.0, 5.0], 0.1)

    def keys(m):
        x = gm.dof_coordinates(m).copy()
        x[:, 1] = np.mod(np.round(x[:, 1], 9), 1.0)
        c = gm.dof_components(m)
        return [tuple(np.round(x[i], 8)) + (int(c[i]),) for i in range(m.n_dofs)]

    def fld(ks, seed):
        return np.array([np.sin(37.0 * k[0] + 69.0 * k[1] + 5.0 * k[2] + seed + k[3]) for k in ks])

    N, sh = 4, (2, 5, 2)
    full = gm.cylinder_shell((2 * sh[0], sh[1], 2 * sh[2]), 2)
    kf = keys(full)
    pos = {k: i for i, k in enumerate(kf)}
    of = make_oracle(full, ti, ctd=True, cell_wise=False, nu=0.001)
    of.set_previous_solution([fld(kf, s) for s in (5, 6, 7)], ti.get_weights())
    of.set_linearization_point(fld(kf, 1), 0.1)
    ref = of.vmult(fld(kf, 2), 15.0)
    acc, owned = np.zeros(full.n_dofs), np.zeros(full.n_dofs, dtype=int)
    for r in range(N):
        m = gm.cylinder_shell_box(sh, 2, n_ranks=N, rank=r)
        km = keys(m)
        idx = np.array([pos[k] for k in km])
        owned[idx[:m.n_owned]] += 1
        o = make_oracle(m, ti, ctd=True, cell_wise=False, nu=0.001)
        o.set_previous_solution([fld(km, s) for s in (5, 6, 7)], ti.get_weights())
        o.set_linearization_point(fld(km, 1), 0.1)
        x = fld(km, 2).copy()
        cons = np.array(sorted(m.constraints), dtype=np.int64)
        x[cons] = 0
        np.add.at(acc, idx, o._scatter(o._apply_cells(o._gather(x), 15.0, False)))
        assert set(idx[cons].tolist()) <= set(full.constraints.keys())
    assert (owned == 1).all()
    free = np.ones(full.n_dofs, dtype=bool)
    free[np.array(sorted(full.constraints), dtype=np.int64)] = False
    assert np.linalg.norm((acc - ref)[free]) <= 1e-13 * np.linalg.norm(ref[free])
