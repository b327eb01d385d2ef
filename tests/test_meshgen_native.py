"""The C helpers of the synthetic generator (csrc/glsb_meshgen.c) against the numpy code paths of mesh.py: the
arrays have to be identical, whatever the shape, degree, periodicity, traversal order or partition."""
import numpy as np
import pytest

from dealii_ns_gls_b200 import mesh as gm


def _both(fn, monkeypatch):
    monkeypatch.delenv("GLSB_MESHGEN_NUMPY", raising=False)
    assert gm._native() is not None, "libglsb_meshgen.so missing: run build()"
    a = fn()
    monkeypatch.setenv("GLSB_MESHGEN_NUMPY", "1")
    assert gm._native() is None
    b = fn()
    return a, b


def _same_mesh(a, b):
    assert a.cell_dofs.dtype == b.cell_dofs.dtype and np.array_equal(a.cell_dofs, b.cell_dofs)
    assert np.array_equal(a.cell_coords, b.cell_coords)
    assert np.array_equal(a.cell_is_boundary, b.cell_is_boundary)
    assert (a.n_dofs, a.n_owned, a.n_cells, a.n_global_dofs) == (b.n_dofs, b.n_owned, b.n_cells, b.n_global_dofs)
    assert sorted(a.constraints) == sorted(b.constraints)
    if a.cell_points is not None:
        assert np.array_equal(a.cell_points, b.cell_points)
    if a.canonical_ids is not None:
        assert np.array_equal(a.canonical_ids, b.canonical_ids)


@pytest.mark.parametrize("dim,shape,p,periodic,order", [
    (3, (5, 4, 3), 2, (False, False, False), "morton"),
    (3, (4, 4, 4), 1, (False, False, False), "morton"),
    (3, (2, 6, 3), 3, (False, True, False), "morton"),
    (3, (3, 3, 2), 4, (False, False, False), "lex"),
    (2, (7, 5), 2, (False, True), "morton"),
    (2, (1, 9), 3, (False, False), "lex"),
])
def test_number_nodes(dim, shape, p, periodic, order, monkeypatch):
    a, b = _both(lambda: gm._number_nodes(dim, shape, p, periodic, order), monkeypatch)
    for x, y in zip(a[:5], b[:5]):
        assert np.array_equal(x, y)
    assert tuple(a[5]) == tuple(b[5])


def test_generators(monkeypatch):
    _same_mesh(*_both(lambda: gm.hypercube(3, 6, 2), monkeypatch))
    _same_mesh(*_both(lambda: gm.hypercube(2, 5, 3, index_dtype=np.int64), monkeypatch))
    _same_mesh(*_both(lambda: gm.cylinder_shell((2, 8, 3), 2), monkeypatch))
    _same_mesh(*_both(lambda: gm.hypercube(3, 4, 2, numbering="component"), monkeypatch))
    for n_ranks in (2, 4, 8):
        for rank in range(n_ranks):
            _same_mesh(*_both(lambda: gm.hypercube_box(4, 2, n_ranks=n_ranks, rank=rank), monkeypatch))
    _same_mesh(*_both(lambda: gm.hypercube_slab(4, 2, n_ranks=3, rank=1), monkeypatch))
    _same_mesh(*_both(lambda: gm.cylinder_shell_box((2, 8, 2), 2, n_ranks=4, rank=3), monkeypatch))
    _same_mesh(*_both(lambda: gm.structured_mesh(3, (4, 4, 4), 2, n_ranks=3, rank=1), monkeypatch))


@pytest.mark.parametrize("shape,p", [((3, 8), 2), ((2, 6, 2), 2), ((2, 5, 2), 3), ((3, 8), 1)])
def test_general_geometry(shape, p, monkeypatch):
    m = gm.cylinder_shell(shape, p)
    (ia, ja), (ib, jb) = _both(lambda: gm.general_geometry(m), monkeypatch)
    assert ia.shape == ib.shape and ja.shape == jb.shape
    assert np.abs(ia - ib).max() <= 1e-13 * np.abs(ib).max()
    assert np.abs(ja - jb).max() <= 1e-14 * np.abs(jb).max()
    (ia, ja), (ib, jb) = _both(lambda: gm.general_geometry(m, n_q_1d=p + 2), monkeypatch)
    assert np.abs(ia - ib).max() <= 1e-13 * np.abs(ib).max() and np.abs(ja - jb).max() <= 1e-14 * np.abs(jb).max()


def test_vertex_geometry(monkeypatch):
    for m in (gm.cylinder_shell((3, 8), 2), gm.cylinder_shell((2, 6, 3), 2)):
        k, dim = m.mapping_degree, m.dim
        vid = [sum((k * ((v >> e) & 1)) * (k + 1) ** e for e in range(dim)) for v in range(2 ** dim)]
        verts = m.cell_points[:, vid, :]
        (ha, ma), (hb, mb) = _both(lambda: gm._vertex_geometry(verts, dim), monkeypatch)
        assert np.abs(ha / hb - 1).max() < 1e-15 and np.abs(ma / mb - 1).max() < 1e-14
