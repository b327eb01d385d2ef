"""Synthetic mesh / DoF generators for the B200 GLS Navier-Stokes operator.

The reference obtains cells, DoF numbering, constraints, geometry and the
partition from deal.II (DoFHandler, AffineConstraints, MappingQ,
parallel::distributed::Triangulation; performance.cc:29-42, main.cc:230-310).
deal.II is not available to this build, so this module produces the same kind
of *description* (the arrays the deal.II adapter would extract, SURVEY.md
Appendix B) for structured, optionally deformed and optionally periodic blocks:

  * hypercube()      -- performance.cc's unit hypercube, Cartesian cells
  * cylinder_shell() -- an O-grid around a cylinder, extruded (curved cells)

DoF numbering imitates DoFHandler::distribute_dofs on a Morton-ordered forest:
cells are walked in (Morton) order and every node is numbered by the first
cell that touches it; the dim+1 components of a node are consecutive
(FESystem(FE_Q(p), dim+1) numbers all components of a vertex/line/quad/hex
dof together).  A rank owns the nodes first touched by its contiguous cell
range (= lowest touching rank, like deal.II).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
import os

import numpy as np


def gauss_lobatto_points(p: int) -> np.ndarray:
    """Support points of FE_Q(p) on [0,1]."""
    if p == 1:
        return np.array([0.0, 1.0])
    c = np.zeros(p + 1)
    c[p] = 1.0
    dc = np.polynomial.legendre.legder(c)
    x = np.sort(np.real(np.polynomial.legendre.legroots(dc)))
    ddc = np.polynomial.legendre.legder(dc)
    for _ in range(3):
        x = x - np.polynomial.legendre.legval(x, dc) / np.polynomial.legendre.legval(x, ddc)
    x = np.concatenate([[-1.0], x, [1.0]])
    x = 0.5 * (x - x[::-1])
    return 0.5 * (x + 1.0)


def _morton_key(coords: np.ndarray) -> np.ndarray:
    """Interleave the bits of integer coordinates [n, dim] (x lowest)."""
    dim = coords.shape[1]
    key = np.zeros(coords.shape[0], dtype=np.uint64)
    nbits = max(1, int(coords.max()).bit_length()) if coords.size else 1
    for b in range(nbits):
        for e in range(dim):
            key |= ((coords[:, e].astype(np.uint64) >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * dim + e)
    return key


# ---- native helpers (csrc/glsb_meshgen.c -> libglsb_meshgen.so) ---------------------------------------------
# The numbering passes below are O(cells) array sweeps that numpy needs ~20 s for at the bench size; the C
# versions give identical arrays in well under a second.  GLSB_MESHGEN_NUMPY=1 forces the numpy code (the tests
# compare the two); without the library the numpy code is used as well -- this is the synthetic generator, not
# the operator, whose library has no fallback.
_meshgen = None


def _native():
    global _meshgen
    if os.environ.get("GLSB_MESHGEN_NUMPY") == "1":
        return None
    if _meshgen is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libglsb_meshgen.so")
        if not os.path.exists(path):
            _meshgen = False
        else:
            lib = ctypes.CDLL(path)
            P, I, L = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
            lib.glsm_number_nodes.restype = I
            lib.glsm_number_nodes.argtypes = [I, P, I, P, I, P, P, P, P]
            lib.glsm_assemble_cell_dofs.restype = I
            lib.glsm_assemble_cell_dofs.argtypes = [L, I, I, P, P, L, I, P, P]
            lib.glsm_vertex_geometry.restype = I
            lib.glsm_vertex_geometry.argtypes = [I, L, P, P, P]
            lib.glsm_general_geometry.restype = L
            lib.glsm_general_geometry.argtypes = [I, L, I, I, P, P, P, P, P]
            _meshgen = lib
    return _meshgen or None


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _assemble_cell_dofs(cell_nodes, local_of_node, C, n_owned, index_dtype):
    """cell_dofs[cell, c * n_loc + l] = local_of_node[cell_nodes[cell, l]] * C + c (components of a node
    consecutive) in `index_dtype`, and the mask of cells touching an index >= n_owned (ghosts)."""
    lib = _native()
    ncell, n_loc = cell_nodes.shape
    idt = np.dtype(index_dtype)
    if lib is not None and idt in (np.dtype(np.uint32), np.dtype(np.int64)):
        cn = np.ascontiguousarray(cell_nodes, dtype=np.int64)
        lo = np.ascontiguousarray(local_of_node, dtype=np.int64)
        out = np.empty((ncell, C * n_loc), dtype=idt)
        bnd = np.empty(ncell, dtype=np.uint8)
        rc = lib.glsm_assemble_cell_dofs(ncell, n_loc, C, _ptr(cn), _ptr(lo), int(n_owned), idt.itemsize, _ptr(out),
                                         _ptr(bnd))
        if rc != 0:
            raise OverflowError("dof indices do not fit the index type")
        return out, bnd.astype(bool)
    ln = np.asarray(local_of_node)[cell_nodes]
    cell_dofs = np.concatenate([ln * C + c for c in range(C)], axis=1)
    return cell_dofs.astype(index_dtype), (cell_dofs >= n_owned).any(axis=1)


@dataclass
class RankPartition:
    """What Utilities::MPI::Partitioner holds for one rank."""
    rank: int
    n_ranks: int
    n_owned: int
    n_ghost: int
    owned_offset: int                 # first global index owned by this rank
    ghost_global: np.ndarray          # global indices of the ghosts (sorted => grouped by owner)
    ghost_owner: np.ndarray           # owner rank of each ghost
    # per neighbour: (rank, recv_offset_in_ghost_block, recv_count)
    recv: list = field(default_factory=list)
    # per neighbour: (rank, owned-local indices this rank must send)
    send: list = field(default_factory=list)


@dataclass
class Mesh:
    dim: int
    degree: int
    n_cells: int
    n_dofs: int                       # local vector length (owned + ghost)
    n_owned: int
    cell_dofs: np.ndarray             # uint32/int64 [n_cells, C*n^dim], comp-blocked lexicographic
    geometry_type: int                # 0 Cartesian, 2 general
    cell_points: np.ndarray           # [n_cells, (k+1)^dim, dim] mapping support points
    mapping_degree: int
    constraints: dict                 # {dof: [(master, weight), ...]} local indices
    cell_h_min: np.ndarray
    cell_measure: np.ndarray
    cart_inv_jac: np.ndarray | None = None   # [n_cells, dim] diag of J^-1 (Cartesian)
    cart_det: np.ndarray | None = None       # [n_cells]
    partition: RankPartition | None = None
    node_compact: bool = True         # comp c of a node at index0 + c
    n_global_dofs: int = 0
    cell_is_boundary: np.ndarray | None = None  # touches ghost dofs
    canonical_ids: np.ndarray | None = None     # slab meshes: partition-independent id of each local dof
    cell_coords: np.ndarray | None = None       # structured blocks: integer cell coordinates [n_cells, dim]
    shape: tuple | None = None                  # structured blocks: cells per direction
    outflow_faces: dict | None = None           # boundary_faces(...) result: faces with outflow terms
    extent: np.ndarray | None = None            # structured blocks: size of the block
    origin: np.ndarray | None = None

    @property
    def C(self):
        return self.dim + 1

    @property
    def n_loc(self):
        return (self.degree + 1) ** self.dim


def _det_small(J):
    """determinants of a stack of 2 x 2 or 3 x 3 matrices [..., d, d] in closed form"""
    if J.shape[-1] == 2:
        return J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]
    return (J[..., 0, 0] * (J[..., 1, 1] * J[..., 2, 2] - J[..., 1, 2] * J[..., 2, 1])
            - J[..., 0, 1] * (J[..., 1, 0] * J[..., 2, 2] - J[..., 1, 2] * J[..., 2, 0])
            + J[..., 0, 2] * (J[..., 1, 0] * J[..., 2, 1] - J[..., 1, 1] * J[..., 2, 0]))


def _vertex_geometry(verts: np.ndarray, dim: int):
    """minimum_vertex_distance() and measure() from the 2^dim vertices."""
    lib = _native()
    if lib is not None and dim in (2, 3):
        v = np.ascontiguousarray(verts, dtype=np.float64)
        h, meas = np.empty(v.shape[0]), np.empty(v.shape[0])
        if lib.glsm_vertex_geometry(dim, v.shape[0], _ptr(v), _ptr(h), _ptr(meas)) != 0:
            raise RuntimeError("glsm_vertex_geometry failed")
        return h, meas
    nv = verts.shape[1]
    h2 = np.full(verts.shape[0], np.inf)
    for a in range(nv):
        for b in range(a + 1, nv):
            d = verts[:, a] - verts[:, b]
            h2 = np.minimum(h2, np.einsum("ki,ki->k", d, d))
    h = np.sqrt(h2)
    # measure of the multilinear cell: 2-point Gauss per direction is exact
    g = np.array([0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0)])
    meas = np.zeros(verts.shape[0])
    vt = np.ascontiguousarray(verts.transpose(0, 2, 1))            # [cell, i, v]
    for qi in range(2 ** dim):
        xi = [g[(qi >> e) & 1] for e in range(dim)]
        dphi = np.ones((nv, dim))                                  # d phi_v / d xi_e at this point
        for v in range(2 ** dim):
            for e in range(dim):
                for f in range(dim):
                    bit = (v >> f) & 1
                    if f == e:
                        dphi[v, e] *= 1.0 if bit else -1.0
                    else:
                        dphi[v, e] *= xi[f] if bit else 1.0 - xi[f]
        meas += _det_small(vt @ dphi) / 2 ** dim                   # J[cell, i, e]
    return h, meas


def _grid_node_coordinates(npts, p, hcell, origin):
    """[nnode, dim] reference coordinates of the grid nodes of a structured block (node ids lexicographic, x
    fastest): node i of a direction lies in cell i // p at the Gauss-Lobatto point i % p."""
    dim = len(npts)
    gp = gauss_lobatto_points(p)
    axes = []
    for e in range(dim):
        i = np.arange(npts[e])
        axes.append(origin[e] + (i // p + gp[i % p]) * hcell[e])
    grids = np.meshgrid(*axes[::-1], indexing="ij")               # slowest direction first
    return np.stack([g.reshape(-1) for g in grids[::-1]], axis=1)


def _zero_constraints_by_node(constraints, dirichlet, coords, local_of_node, C):
    """constraints[dof] = [] for every (node, component) the mask callable selects; same rows, same insertion
    order (by component, ascending dof) as the cell-wise evaluation, one evaluation per node instead of one per
    cell-local node"""
    for c in range(C):
        mask = np.asarray(dirichlet(coords, c), dtype=bool)
        dofs = np.sort(local_of_node[mask] * C + c)
        constraints.update({d: [] for d in dofs.tolist()})


def _number_nodes(dim, shape, p, periodic, order):
    """Cells in traversal order and the first-touch node numbering on a structured block.

    Returns cc[ncell, dim] (cell coordinates in traversal order), loc[n_loc, dim] (local node
    offsets, x fastest), cell_nodes[ncell, n_loc] (lexicographic grid node ids), node_rank[nnode]
    (position of every grid node in the first-touch numbering), first_cell[nnode], npts."""
    n = p + 1
    n_loc = n ** dim
    ncell = int(np.prod(shape))
    lib = _native()
    if lib is not None:
        npts = tuple(p * shape[e] + (0 if periodic[e] else 1) for e in range(dim))
        nnode = int(np.prod(npts))
        cc = np.empty((ncell, dim), dtype=np.int64)
        cell_nodes = np.empty((ncell, n_loc), dtype=np.int64)
        node_rank = np.empty(nnode, dtype=np.int64)
        first_cell = np.empty(nnode, dtype=np.int64)
        shp = np.asarray(shape, dtype=np.int64)
        per = np.asarray(periodic, dtype=np.uint8)
        rc = lib.glsm_number_nodes(dim, _ptr(shp), p, _ptr(per), 1 if order == "morton" else 0, _ptr(cc),
                                   _ptr(cell_nodes), _ptr(node_rank), _ptr(first_cell))
        if rc != 0:
            raise RuntimeError("glsm_number_nodes failed")
        loc = np.stack(np.meshgrid(*[np.arange(n)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)[:, ::-1].copy()
        return cc, loc, cell_nodes, node_rank, first_cell, npts
    # ---- cell traversal order -------------------------------------------------
    cc = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), axis=-1).reshape(-1, dim)
    if order == "morton":
        perm = np.argsort(_morton_key(cc), kind="stable")
    else:  # lexicographic, x fastest
        key = np.zeros(ncell, dtype=np.int64)
        mul = 1
        for e in range(dim):
            key += cc[:, e] * mul
            mul *= shape[e]
        perm = np.argsort(key, kind="stable")
    cc = cc[perm]

    # ---- node grid ------------------------------------------------------------
    npts = tuple(p * shape[e] + (0 if periodic[e] else 1) for e in range(dim))
    nnode = int(np.prod(npts))
    # local lexicographic offsets
    loc = np.stack(np.meshgrid(*[np.arange(n)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)
    # make x fastest: meshgrid 'ij' makes the last axis fastest -> reverse columns
    loc = loc[:, ::-1].copy()  # loc[l] = (i_x, i_y, i_z) with x fastest
    cell_nodes = np.zeros((ncell, n_loc), dtype=np.int64)
    mul = 1
    for e in range(dim):
        g = (p * cc[:, e][:, None] + loc[None, :, e]) % npts[e]
        cell_nodes += g * mul
        mul *= npts[e]

    # ---- numbering: first-touching cell wins ----------------------------------
    # local ordering inside a cell: vertices, lines, quads, hexes, then lexicographic
    ent_dim = ((loc > 0) & (loc < p)).sum(axis=1)
    sortpos = np.empty(n_loc, dtype=np.int64)
    sortpos[np.lexsort((np.arange(n_loc), ent_dim))] = np.arange(n_loc)
    colorder = np.argsort(sortpos)
    flat_nodes = cell_nodes[:, colorder].reshape(-1)   # ascending key = position in this array
    first_key = np.empty(nnode, dtype=np.int64)
    # walk the keys in descending order: the smallest key of a node is written last and wins
    first_key[flat_nodes[::-1]] = np.arange(flat_nodes.size - 1, -1, -1, dtype=np.int64)
    node_rank = np.empty(nnode, dtype=np.int64)
    node_rank[np.argsort(first_key, kind="stable")] = np.arange(nnode)
    first_cell = first_key // n_loc

    return cc, loc, cell_nodes, node_rank, first_cell, npts


def structured_mesh(dim, shape, degree, *, deform=None, mapping_degree=1, periodic=None,
                    order="morton", numbering="node", dirichlet=None,
                    n_ranks=1, rank=0, extent=None, origin=None, index_dtype=np.uint32):
    """Structured block of prod(shape) cells.

    deform:    callable(points[..., dim]) -> points[..., dim]; None => Cartesian cells
    periodic:  tuple of bools per direction (node identification, O-grid)
    dirichlet: callable(ref_coords[n_nodes, dim], comp) -> bool mask of zero-constrained nodes
    numbering: "node" (components of a node consecutive, deal.II-like) or
               "component" (component-major blocks; exercises the general index path)
    """
    shape = tuple(int(s) for s in shape)
    assert len(shape) == dim
    periodic = tuple(periodic) if periodic is not None else (False,) * dim
    extent = np.ones(dim) if extent is None else np.asarray(extent, dtype=np.float64)
    origin = np.zeros(dim) if origin is None else np.asarray(origin, dtype=np.float64)
    p = degree
    n = p + 1
    C = dim + 1
    n_loc = n ** dim
    ncell = int(np.prod(shape))

    cc, loc, cell_nodes, node_rank, first_cell, npts = _number_nodes(dim, shape, p, periodic, order)
    nnode = int(np.prod(npts))

    # ---- partition ------------------------------------------------------------
    bounds = [(ncell * r) // n_ranks for r in range(n_ranks + 1)]
    if n_ranks == 1:
        owned_counts = np.array([nnode])
    else:
        node_owner_of_rank = np.searchsorted(np.asarray(bounds[1:]), first_cell, side="right")
        # owned node ranges in the global numbering are contiguous
        owned_counts = np.bincount(node_owner_of_rank, minlength=n_ranks)
    owned_off_nodes = np.concatenate([[0], np.cumsum(owned_counts)])

    c0, c1 = bounds[rank], bounds[rank + 1]
    my_cells = slice(c0, c1)
    ncell_loc = c1 - c0
    fast = n_ranks == 1 and numbering == "node"   # one rank, components of a node consecutive: assembled in C
    my_nodes_g = None if fast else node_rank[cell_nodes[my_cells]]  # global node ranks [ncell_loc, n_loc]

    if numbering == "node":
        def gdof(noderank, c):
            return noderank * C + c
    else:
        def gdof(noderank, c):
            # component-major inside each owner's range keeps ownership contiguous
            own = np.searchsorted(owned_off_nodes[1:], noderank, side="right")
            base = owned_off_nodes[own]
            cnt = owned_counts[own]
            return base * C + c * cnt + (noderank - base)

    n_global_dofs = nnode * C
    owned_lo = owned_off_nodes[rank] * C
    owned_hi = owned_off_nodes[rank + 1] * C
    n_owned = int(owned_hi - owned_lo)
    is_boundary = None
    cell_gdofs = None if fast else np.concatenate([gdof(my_nodes_g, c) for c in range(C)], axis=1)  # comp-blocked
    if fast:
        n_ghost = 0
        local, is_boundary = _assemble_cell_dofs(cell_nodes, node_rank, C, n_owned, index_dtype)
    elif n_ranks == 1:
        ghosts = np.zeros(0, dtype=np.int64)
        n_ghost = 0
        local = cell_gdofs
    else:
        uniq = np.unique(cell_gdofs)
        ghosts = uniq[(uniq < owned_lo) | (uniq >= owned_hi)]
        n_ghost = len(ghosts)
        # local index map
        is_owned = (cell_gdofs >= owned_lo) & (cell_gdofs < owned_hi)
        local = np.where(is_owned, cell_gdofs - owned_lo, 0)
        if n_ghost:
            gpos = np.searchsorted(ghosts, cell_gdofs)
            gpos = np.clip(gpos, 0, n_ghost - 1)
            local = np.where(is_owned, local, n_owned + gpos)
    n_local = n_owned + n_ghost

    part = None
    if n_ranks > 1:
        ghost_owner = np.searchsorted(owned_off_nodes[1:] * C, ghosts, side="right")
        part = RankPartition(rank=rank, n_ranks=n_ranks, n_owned=n_owned, n_ghost=n_ghost,
                             owned_offset=int(owned_lo), ghost_global=ghosts, ghost_owner=ghost_owner)
        for r in np.unique(ghost_owner):
            sel = np.nonzero(ghost_owner == r)[0]
            part.recv.append((int(r), int(sel[0]), int(len(sel))))
        # what others need from me: recompute their ghost sets (cheap for test sizes; the
        # deal.II Partitioner gets this from a consensus exchange)
        for r in range(n_ranks):
            if r == rank:
                continue
            oc = slice(bounds[r], bounds[r + 1])
            on = node_rank[cell_nodes[oc]]
            og = np.unique(np.concatenate([gdof(on, c) for c in range(C)], axis=1))
            mine = og[(og >= owned_lo) & (og < owned_hi)]
            if len(mine):
                part.send.append((r, (mine - owned_lo).astype(np.int64)))

    # ---- geometry -------------------------------------------------------------
    hcell = extent / np.asarray(shape, dtype=np.float64)
    k = mapping_degree if deform is not None else 1
    mp = gauss_lobatto_points(k)
    mloc = np.stack(np.meshgrid(*[np.arange(k + 1)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)[:, ::-1]
    ref = origin[None, None, :] + (cc[my_cells][:, None, :] + mp[mloc][None, :, :]) * hcell[None, None, :]
    pts = deform(ref) if deform is not None else ref
    verts_idx = [sum((k * ((v >> e) & 1)) * (k + 1) ** e for e in range(dim)) for v in range(2 ** dim)]
    if deform is None:
        h_min = np.full(ncell_loc, float(hcell.min()))
        meas = np.full(ncell_loc, float(np.prod(hcell)))
    else:
        h_min, meas = _vertex_geometry(pts[:, verts_idx, :], dim)

    mesh = Mesh(dim=dim, degree=p, n_cells=ncell_loc, n_dofs=n_local, n_owned=n_owned,
                cell_dofs=local.astype(index_dtype, copy=False), geometry_type=0 if deform is None else 2,
                cell_points=pts, mapping_degree=k, constraints={}, cell_h_min=h_min,
                cell_measure=meas, partition=part, node_compact=(numbering == "node"),
                n_global_dofs=n_global_dofs, cell_coords=np.ascontiguousarray(cc[my_cells]), shape=shape,
                extent=extent, origin=origin)
    if deform is None:
        mesh.cart_inv_jac = np.broadcast_to(1.0 / hcell, (ncell_loc, dim)).copy()
        mesh.cart_det = np.full(ncell_loc, float(np.prod(hcell)))
    mesh.cell_is_boundary = (local >= n_owned).any(axis=1) if is_boundary is None else is_boundary

    # ---- zero (Dirichlet-type) constraints -------------------------------------
    if dirichlet is not None and fast:
        _zero_constraints_by_node(mesh.constraints, dirichlet, _grid_node_coordinates(npts, p, hcell, origin),
                                  node_rank, C)
    elif dirichlet is not None:
        gp = gauss_lobatto_points(p)
        nref = origin[None, None, :] + (cc[my_cells][:, None, :] + gp[loc][None, :, :]) * hcell[None, None, :]
        for c in range(C):
            mask = dirichlet(nref.reshape(-1, dim), c).reshape(ncell_loc, n_loc)
            dofs = local[:, c * n_loc:(c + 1) * n_loc][mask]
            for dof in np.unique(dofs):
                mesh.constraints[int(dof)] = []
    return mesh


def _cells_touching(mesh: Mesh, only):
    """(cells, local node, component) triples of the entries of cell_dofs that hold one of the dofs `only`, and
    the position of each in `only`"""
    pos = np.full(mesh.n_dofs, -1, dtype=np.int64)
    pos[np.asarray(only, dtype=np.int64)] = np.arange(len(only))
    hit = pos[mesh.cell_dofs]                       # [n_cells, C * n_loc]
    cell, col = np.nonzero(hit >= 0)
    return cell, col % mesh.n_loc, col // mesh.n_loc, hit[cell, col]


def dof_coordinates(mesh: Mesh, only=None) -> np.ndarray:
    """[n_dofs, dim] coordinates of the support point of every local dof of an undeformed structured block
    (what DoFTools::map_dofs_to_support_points gives; used for boundary values).  only = array of dofs: the
    coordinates of just those, [len(only), dim]."""
    dim, p = mesh.dim, mesh.degree
    n = p + 1
    gp = gauss_lobatto_points(p)
    loc = np.stack(np.meshgrid(*[np.arange(n)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)[:, ::-1]
    hcell = mesh.extent / np.asarray(mesh.shape, dtype=np.float64)
    if only is not None:
        cell, l, _, where = _cells_touching(mesh, only)
        out = np.zeros((len(only), dim))
        out[where] = mesh.origin[None, :] + (mesh.cell_coords[cell] + gp[loc][l]) * hcell[None, :]
        return out
    x = mesh.origin[None, None, :] + (mesh.cell_coords[:, None, :] + gp[loc][None, :, :]) * hcell[None, None, :]
    out = np.zeros((mesh.n_dofs, dim))
    for c in range(dim + 1):
        out[mesh.cell_dofs[:, c * mesh.n_loc:(c + 1) * mesh.n_loc].astype(np.int64).reshape(-1)] = x.reshape(-1, dim)
    return out


def cell_diameters(mesh: Mesh) -> np.ndarray:
    """cell->diameter(): the largest distance between two vertices (GridTools::minimal_cell_diameter takes the
    minimum over the cells; the CFL time step of main.cc:916-918 uses it)."""
    k, dim = mesh.mapping_degree, mesh.dim
    vid = [sum((k * ((v >> e) & 1)) * (k + 1) ** e for e in range(dim)) for v in range(2 ** dim)]
    verts = mesh.cell_points[:, vid, :]
    d = np.zeros(mesh.n_cells)
    for a in range(len(vid)):
        for b in range(a + 1, len(vid)):
            d = np.maximum(d, np.sqrt(((verts[:, a] - verts[:, b]) ** 2).sum(axis=1)))
    return d


def dof_components(mesh: Mesh, only=None) -> np.ndarray:
    """[n_dofs] component (0..dim) of every local dof (only = array of dofs: of just those)."""
    if only is not None:
        _, _, comp, where = _cells_touching(mesh, only)
        out = np.zeros(len(only), dtype=np.int64)
        out[where] = comp
        return out
    out = np.zeros(mesh.n_dofs, dtype=np.int64)
    for c in range(mesh.dim + 1):
        out[mesh.cell_dofs[:, c * mesh.n_loc:(c + 1) * mesh.n_loc].astype(np.int64).reshape(-1)] = c
    return out


def child_cells(coarse: Mesh, fine: Mesh) -> np.ndarray:
    """[n_coarse_cells, 2^dim] indices of the children of every coarse cell in the once-refined structured
    block `fine` (child number cx + 2 cy + 4 cz, like GeometryInfo); what MGTwoLevelTransfer::reinit finds by
    walking the two triangulations (main.cc:540-556)."""
    dim = coarse.dim
    assert fine.shape == tuple(2 * s for s in coarse.shape), "fine must be the global refinement of coarse"
    lookup = np.full(fine.shape[::-1], -1, dtype=np.int64)  # indexed [z][y][x]
    lookup[tuple(fine.cell_coords[:, e] for e in reversed(range(dim)))] = np.arange(fine.n_cells)
    out = np.empty((coarse.n_cells, 2 ** dim), dtype=np.int64)
    for ch in range(2 ** dim):
        fc = 2 * coarse.cell_coords + np.array([(ch >> e) & 1 for e in range(dim)])
        out[:, ch] = lookup[tuple(fc[:, e] for e in reversed(range(dim)))]
    assert (out >= 0).all()
    return out


def hypercube(dim, n_per_dir, degree, **kw):
    """performance.cc:29-31: GridGenerator::hyper_cube + uniform cells."""
    shape = (n_per_dir,) * dim if np.isscalar(n_per_dir) else tuple(n_per_dir)
    return structured_mesh(dim, shape, degree, **kw)


def hypercube_slab(n_per_dir, degree, *, n_ranks=1, rank=0, dim=3, order="morton", with_points=True,
                   index_dtype=np.uint32):
    """Rank `rank`'s part of a hypercube of n x .. x (n * n_ranks) cells cut into n_ranks slabs
    along the last direction (the weak-scaling workload of bench.py).

    Only this rank's box is generated (O(local cells)).  Ownership follows deal.II: a node shared
    by two ranks belongs to the lower rank, so rank r > 0 sees its bottom node plane as ghosts
    owned by r - 1 and rank r < R - 1 exports its top plane.  Both sides order the exchanged
    plane canonically (node lexicographic in the plane, then component), which is all the
    Partitioner-style send / receive lists need."""
    p, C, n = degree, dim + 1, degree + 1
    n_loc = n ** dim
    shape = (n_per_dir,) * dim
    cc, loc, cell_nodes, node_rank, first_cell, npts = _number_nodes(dim, shape, p, (False,) * dim, order)
    nnode = int(np.prod(npts))
    ncell = cc.shape[0]
    plane = int(np.prod(npts[:-1]))               # nodes per plane normal to the slab direction
    node_ids = np.arange(nnode)
    k_of = node_ids // plane                      # index along the slab direction
    is_ghost_node = (k_of == 0) if rank > 0 else np.zeros(nnode, dtype=bool)
    # owned nodes keep their first-touch order
    owned_nodes = np.nonzero(~is_ghost_node)[0]
    owned_sorted = owned_nodes[np.argsort(node_rank[owned_nodes], kind="stable")]
    n_owned_nodes = len(owned_sorted)
    local_of_node = np.empty(nnode, dtype=np.int64)
    local_of_node[owned_sorted] = np.arange(n_owned_nodes)
    ghost_nodes = np.nonzero(is_ghost_node)[0]    # lexicographic in the plane = canonical order
    local_of_node[ghost_nodes] = n_owned_nodes + np.arange(len(ghost_nodes))
    n_owned, n_ghost = n_owned_nodes * C, len(ghost_nodes) * C
    cell_dofs, is_boundary = _assemble_cell_dofs(cell_nodes, local_of_node, C, n_owned, index_dtype)

    # rank 0 owns all its node planes, every other rank all but its bottom plane
    first_owned = 0 if rank == 0 else C * plane * (p * n_per_dir * rank + 1)
    part = RankPartition(rank=rank, n_ranks=n_ranks, n_owned=n_owned, n_ghost=n_ghost, owned_offset=first_owned,
                         ghost_global=np.zeros(0, dtype=np.int64), ghost_owner=np.full(n_ghost, rank - 1))
    if rank > 0:
        part.recv.append((rank - 1, 0, n_ghost))
    if rank < n_ranks - 1:
        top = node_ids[k_of == npts[-1] - 1]
        exp = (local_of_node[top][:, None] * C + np.arange(C)[None, :]).reshape(-1)
        part.send.append((rank + 1, exp.astype(np.int64)))

    ext = np.ones(dim)
    ext[-1] = float(n_ranks)
    hcell = 1.0 / n_per_dir
    nodes_last = p * n_per_dir * n_ranks + 1
    mesh = Mesh(dim=dim, degree=p, n_cells=ncell, n_dofs=n_owned + n_ghost, n_owned=n_owned,
                cell_dofs=cell_dofs, geometry_type=0, cell_points=None, mapping_degree=1,
                constraints={}, cell_h_min=np.full(ncell, hcell), cell_measure=np.full(ncell, hcell ** dim),
                partition=part, node_compact=True, n_global_dofs=plane * nodes_last * C)
    mesh.cart_inv_jac = np.full((ncell, dim), 1.0 / hcell)
    mesh.cart_det = np.full(ncell, hcell ** dim)
    mesh.cell_is_boundary = is_boundary
    # canonical global id of every local dof (tests use it to compare against a 1-rank run)
    gz = k_of + rank * p * n_per_dir
    gnode = (node_ids % plane) + plane * gz
    canon = np.empty(n_owned + n_ghost, dtype=np.int64)
    for c in range(C):
        canon[local_of_node * C + c] = gnode * C + c
    mesh.canonical_ids = canon
    if with_points:
        origin = np.zeros(dim)
        origin[-1] = rank * 1.0
        mloc = np.stack(np.meshgrid(*[np.arange(2)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)[:, ::-1]
        mesh.cell_points = origin[None, None, :] + (cc[:, None, :] + mloc[None, :, :].astype(np.float64)) * hcell
    return mesh


def morton_rank_grid(n_ranks, dim=3):
    """Boxes per direction when the Morton curve of a uniformly refined hypercube is cut into n_ranks equal
    contiguous pieces (p4est's partition, performance.cc:29-31): the leading bits of the key belong to
    z, y, x in turn, so 2 ranks are z-halves, 4 ranks (z, y)-quarters, 8 ranks octants, 16 ranks octants
    halved in z again, ...  Returns (grid, coords_of_rank) with rank = position along the curve."""
    k = int(n_ranks).bit_length() - 1
    assert 1 << k == n_ranks, "power-of-two rank counts only"
    bits = [0] * dim                              # key bits per direction, assigned from the top: z, y, x, z, ...
    for b in range(k):
        bits[dim - 1 - (b % dim)] += 1
    grid = tuple(1 << b for b in bits)
    coords = np.zeros((n_ranks, dim), dtype=np.int64)
    for r in range(n_ranks):
        used = [0] * dim
        for b in range(k):                        # b-th bit from the top of the k-bit rank
            e = dim - 1 - (b % dim)
            bit = (r >> (k - 1 - b)) & 1
            used[e] += 1
            coords[r, e] |= bit << (bits[e] - used[e])
    return grid, coords


def structured_box(shape, degree, *, grid, box_of, rank, periodic=None, deform=None, mapping_degree=1,
                   dirichlet=None, box_extent=None, order="morton", with_points=True, index_dtype=np.uint32):
    """Rank `rank`'s box of prod(shape) cells out of a grid of equal boxes (box_of[r] = box coordinates of rank r,
    monotone: a box with smaller coordinates has the smaller rank), generated in O(local cells).

    Ownership follows deal.II (lowest touching rank): a box owns its nodes except those on a low face behind
    which another box lies; those are ghosts owned by the box diagonally below through every such face the node
    lies on, so a box imports from up to 2^dim - 1 neighbours (faces, edges, corner) and exports its high
    faces to as many.  Ghosts are grouped by owner; both sides order a neighbour's nodes lexicographically,
    which is all the Partitioner-style send / receive lists need.  Directions with periodic node identification
    (O-grid) must not be split.  deform / dirichlet act on the reference coordinates of the WHOLE block."""
    dim = len(shape)
    p, C, n = degree, dim + 1, degree + 1
    shape = tuple(int(x) for x in shape)
    grid = tuple(int(g) for g in grid)
    box_of = np.asarray(box_of, dtype=np.int64)
    n_ranks = len(box_of)
    periodic = tuple(periodic) if periodic is not None else (False,) * dim
    assert all(not (periodic[e] and grid[e] > 1) for e in range(dim)), "periodic directions are not split"
    b = box_of[rank]
    rank_of_box = {tuple(box_of[r]): r for r in range(n_ranks)}
    cc, loc, cell_nodes, node_rank, first_cell, npts = _number_nodes(dim, shape, p, periodic, order)
    nnode, ncell = int(np.prod(npts)), cc.shape[0]
    node_ids = np.arange(nnode)
    ijk = np.stack([(node_ids // int(np.prod(npts[:e]))) % npts[e] for e in range(dim)], axis=1)  # x fastest
    low = (ijk == 0) & (b[None, :] > 0)           # on a low face with a neighbour behind it
    is_ghost_node = low.any(axis=1)
    owner_box = b[None, :] - low.astype(np.int64)
    owner_rank = np.array([rank_of_box[tuple(o)] for o in owner_box[is_ghost_node]], dtype=np.int64) \
        if is_ghost_node.any() else np.zeros(0, dtype=np.int64)
    owned_nodes = np.nonzero(~is_ghost_node)[0]
    owned_sorted = owned_nodes[np.argsort(node_rank[owned_nodes], kind="stable")]   # first-touch order
    n_owned_nodes = len(owned_sorted)
    local_of_node = np.empty(nnode, dtype=np.int64)
    local_of_node[owned_sorted] = np.arange(n_owned_nodes)
    ghost_nodes = np.nonzero(is_ghost_node)[0]
    gorder = np.lexsort((ghost_nodes, owner_rank))                                  # by owner, then lexicographic
    ghost_nodes, owner_rank = ghost_nodes[gorder], owner_rank[gorder]
    local_of_node[ghost_nodes] = n_owned_nodes + np.arange(len(ghost_nodes))
    n_owned, n_ghost = n_owned_nodes * C, len(ghost_nodes) * C
    cell_dofs, is_boundary = _assemble_cell_dofs(cell_nodes, local_of_node, C, n_owned, index_dtype)

    def owned_count(bb):
        return int(np.prod([p * shape[e] + (1 if (bb[e] == 0 and not periodic[e]) else 0) for e in range(dim)])) * C

    first_owned = sum(owned_count(box_of[r]) for r in range(rank))
    part = RankPartition(rank=rank, n_ranks=n_ranks, n_owned=n_owned, n_ghost=n_ghost, owned_offset=first_owned,
                         ghost_global=np.zeros(0, dtype=np.int64), ghost_owner=np.repeat(owner_rank, C))
    for r in np.unique(owner_rank):
        sel = np.nonzero(owner_rank == r)[0]
        part.recv.append((int(r), int(sel[0]) * C, int(len(sel)) * C))
    # exports: neighbour b + delta imports from me the nodes on my high faces in supp(delta) that I own
    top = np.asarray(npts) - 1
    for r in range(n_ranks):
        delta = box_of[r] - b
        if r == rank or (delta < 0).any() or (delta > 1).any():
            continue
        sel = ~is_ghost_node
        for e in range(dim):
            if delta[e] == 1:
                sel &= ijk[:, e] == top[e]
        exp_nodes = node_ids[sel]                                                    # lexicographic
        exp = (local_of_node[exp_nodes][:, None] * C + np.arange(C)[None, :]).reshape(-1)
        part.send.append((int(r), exp.astype(np.int64)))

    box_extent = np.ones(dim) if box_extent is None else np.asarray(box_extent, dtype=np.float64)
    hcell = box_extent / np.asarray(shape, dtype=np.float64)
    origin = b.astype(np.float64) * box_extent
    gn = [p * shape[e] * grid[e] + (0 if periodic[e] else 1) for e in range(dim)]
    k = mapping_degree if deform is not None else 1
    pts = None
    if with_points or deform is not None:
        mp = gauss_lobatto_points(k)
        mloc = np.stack(np.meshgrid(*[np.arange(k + 1)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)[:, ::-1]
        ref = origin[None, None, :] + (cc[:, None, :] + mp[mloc][None, :, :]) * hcell[None, None, :]
        pts = deform(ref) if deform is not None else ref
    if deform is None:
        h_min, meas = np.full(ncell, float(hcell.min())), np.full(ncell, float(np.prod(hcell)))
    else:
        verts_idx = [sum((k * ((v >> e) & 1)) * (k + 1) ** e for e in range(dim)) for v in range(2 ** dim)]
        h_min, meas = _vertex_geometry(pts[:, verts_idx, :], dim)
    mesh = Mesh(dim=dim, degree=p, n_cells=ncell, n_dofs=n_owned + n_ghost, n_owned=n_owned,
                cell_dofs=cell_dofs, geometry_type=0 if deform is None else 2, cell_points=pts,
                mapping_degree=k, constraints={}, cell_h_min=h_min, cell_measure=meas,
                partition=part, node_compact=True, n_global_dofs=int(np.prod(gn)) * C)
    if deform is None:
        mesh.cart_inv_jac = np.broadcast_to(1.0 / hcell, (ncell, dim)).copy()
        mesh.cart_det = np.full(ncell, float(np.prod(hcell)))
    mesh.cell_is_boundary = is_boundary
    gijk = ijk + (b * p * np.asarray(shape))[None, :]
    gnode = np.zeros(nnode, dtype=np.int64)
    mul = 1
    for e in range(dim):
        gnode += (gijk[:, e] % gn[e]) * mul
        mul *= gn[e]
    canon = np.empty(n_owned + n_ghost, dtype=np.int64)
    for c in range(C):
        canon[local_of_node * C + c] = gnode * C + c
    mesh.canonical_ids = canon
    mesh.shape, mesh.cell_coords = shape, np.ascontiguousarray(cc)
    mesh.extent, mesh.origin = box_extent, origin
    if dirichlet is not None:
        _zero_constraints_by_node(mesh.constraints, dirichlet, _grid_node_coordinates(npts, p, hcell, origin),
                                  local_of_node, C)
    return mesh


def hypercube_box(n_per_dir, degree, *, n_ranks=1, rank=0, dim=3, order="morton", with_points=True,
                  index_dtype=np.uint32):
    """Rank `rank`'s box of n^dim cells of a hypercube partitioned the way deal.II / p4est partition it: equal
    pieces of the Morton curve (morton_rank_grid: halves, quarters, octants; performance.cc:29-31)."""
    grid, box_of = morton_rank_grid(n_ranks, dim)
    return structured_box((n_per_dir,) * dim, degree, grid=grid, box_of=box_of, rank=rank, order=order,
                          with_points=with_points, index_dtype=index_dtype)


def cylinder_shell_box(shape, degree, *, n_ranks=1, rank=0, r_inner=0.05, r_outer=0.5, length=0.41,
                       mapping_degree=None, no_slip=True, **kw):
    """Rank `rank`'s part of the O-grid of cylinder_shell() with `shape` = (radial, circumferential, axial) cells
    PER RANK (weak scaling): the block is cut in the axial direction first, then radially (2 ranks: 1 x 1 x 2
    boxes, 4: 2 x 1 x 2, 8: 2 x 1 x 4); the periodic circumferential direction is never split.  Same map, same
    no-slip rows (cylinder surface and outer wall) as the single-rank generator."""
    mapping_degree = degree if mapping_degree is None else mapping_degree
    grid = {1: (1, 1, 1), 2: (1, 1, 2), 4: (2, 1, 2), 8: (2, 1, 4), 16: (2, 1, 8)}[n_ranks]
    box_of = np.array([(r % grid[0], 0, r // grid[0]) for r in range(n_ranks)], dtype=np.int64)

    def deform(x):
        r = r_inner + (r_outer - r_inner) * x[..., 0] ** 1.5
        th = 2.0 * np.pi * x[..., 1]
        out = np.empty_like(x)
        out[..., 0] = r * np.cos(th)
        out[..., 1] = r * np.sin(th)
        out[..., 2] = length * x[..., 2]
        return out

    def dirichlet(ref, c):
        if c == 3:
            return np.zeros(len(ref), dtype=bool)
        return (np.abs(ref[:, 0]) < 1e-12) | (np.abs(ref[:, 0] - 1.0) < 1e-12)

    return structured_box(shape, degree, grid=grid, box_of=box_of, rank=rank, periodic=(False, True, False),
                          deform=deform, mapping_degree=mapping_degree, dirichlet=dirichlet if no_slip else None,
                          box_extent=1.0 / np.asarray(grid, dtype=np.float64), **kw)


def hypercube_hanging(dim, n_coarse, degree, *, refine=None, dirichlet=None, index_dtype=np.uint32):
    """Unit hypercube of n_coarse^dim cells in which the cells selected by `refine(cell_coords) -> bool mask`
    (default: the block with all coordinates >= n_coarse / 2) are refined once: a true 2:1 mesh with hanging
    nodes on faces (and, in 3-D, on edges where the refined block has a re-entrant edge), as
    DoFTools::make_hanging_node_constraints sees it in the reference (main.cc:293, simulation.cc:803-809,
    input/rotation.json).

    Active cells are walked along the Morton curve of the fine level (p4est order); every node is numbered by the
    first active cell touching it, components of a node consecutive.  A node of a refined cell that lies on an
    unrefined cell K without being one of K's nodes is a hanging node: it is constrained to the value of K's
    finite element function there, x_h = sum_j phi_j^K(x_h) x_j -- deal.II's FE_Q face/line interpolation
    constraints (Q1: 1/2, 1/2 and 1/4 x 4; Q2 on a line at 1/4: 3/8, 3/4, -1/8); weights below 1e-14 dropped.
    Masters are never constrained themselves (one level of difference)."""
    p, C, n = degree, dim + 1, degree + 1
    n_loc = n ** dim
    gp = gauss_lobatto_points(p)
    cc0 = np.stack(np.meshgrid(*[np.arange(n_coarse)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)
    ref_mask = (cc0 >= n_coarse // 2).all(axis=1) if refine is None else np.asarray(refine(cc0), dtype=bool)
    # active cells: (fine-level integer origin, size in fine units)
    cells = []
    for k in range(len(cc0)):
        if ref_mask[k]:
            for ch in range(2 ** dim):
                cells.append((2 * cc0[k] + np.array([(ch >> e) & 1 for e in range(dim)]), 1))
        else:
            cells.append((2 * cc0[k], 2))
    org = np.array([c[0] for c in cells], dtype=np.int64)
    size = np.array([c[1] for c in cells], dtype=np.int64)
    perm = np.argsort(_morton_key(org), kind="stable")
    org, size = org[perm], size[perm]
    ncell = len(org)
    hf = 1.0 / (2 * n_coarse)
    loc = np.stack(np.meshgrid(*[np.arange(n)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)[:, ::-1]
    # local numbering order inside a cell: vertices, lines, quads, hexes (like DoFHandler::distribute_dofs)
    ent_dim = ((loc > 0) & (loc < p)).sum(axis=1)
    colorder = np.lexsort((np.arange(n_loc), ent_dim))
    pts = (org[:, None, :] + size[:, None, None] * gp[loc][None, :, :]) * hf       # [ncell, n_loc, dim]
    keys = np.round(pts * 2 ** 40).astype(np.int64)
    node_of = {}
    cell_nodes = np.zeros((ncell, n_loc), dtype=np.int64)
    for k in range(ncell):
        for l in colorder:
            cell_nodes[k, l] = node_of.setdefault(tuple(keys[k, l]), len(node_of))
    nnode = len(node_of)
    node_xyz = np.zeros((nnode, dim))
    node_xyz[cell_nodes.reshape(-1)] = pts.reshape(-1, dim)
    cell_dofs = np.concatenate([cell_nodes * C + c for c in range(C)], axis=1)

    # ---- hanging-node constraints ------------------------------------------------------------
    def lagrange(x):
        v = np.ones(n)
        for i in range(n):
            for j in range(n):
                if j != i:
                    v[i] *= (x - gp[j]) / (gp[i] - gp[j])
        return v

    fine_nodes = np.unique(cell_nodes[size == 1]) if (size == 1).any() else np.zeros(0, dtype=np.int64)
    node_rows = {}
    tol = 1e-12
    for k in np.nonzero(size == 2)[0]:
        lo, hi = org[k] * hf, (org[k] + 2) * hf
        inside = fine_nodes[((node_xyz[fine_nodes] >= lo - tol) & (node_xyz[fine_nodes] <= hi + tol)).all(axis=1)]
        mine = set(cell_nodes[k].tolist())
        for nd in inside:
            if int(nd) in mine or int(nd) in node_rows:
                continue
            xi = (node_xyz[nd] - lo) / (hi - lo)
            w1 = [lagrange(xi[e]) for e in range(dim)]
            row = []
            for l in range(n_loc):
                w = float(np.prod([w1[e][loc[l, e]] for e in range(dim)]))
                if abs(w) > 1e-14:
                    row.append((int(cell_nodes[k, l]), w))
            node_rows[int(nd)] = row
    constraints = {}
    for nd, row in node_rows.items():
        for c in range(C):
            constraints[nd * C + c] = [(m * C + c, w) for m, w in row]

    hcell = size.astype(np.float64) * hf
    mloc = np.stack(np.meshgrid(*[np.arange(2)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)[:, ::-1]
    cpts = (org[:, None, :] + size[:, None, None] * mloc[None, :, :]) * hf
    mesh = Mesh(dim=dim, degree=p, n_cells=ncell, n_dofs=nnode * C, n_owned=nnode * C,
                cell_dofs=cell_dofs.astype(index_dtype), geometry_type=0, cell_points=cpts, mapping_degree=1,
                constraints=constraints, cell_h_min=hcell.copy(), cell_measure=hcell ** dim, node_compact=True,
                n_global_dofs=nnode * C)
    mesh.cart_inv_jac = np.repeat((1.0 / hcell)[:, None], dim, axis=1)
    mesh.cart_det = hcell ** dim
    mesh.cell_is_boundary = np.zeros(ncell, dtype=bool)
    mesh.node_xyz = node_xyz
    mesh.hanging_nodes = np.array(sorted(node_rows.keys()), dtype=np.int64)
    if dirichlet is not None:
        for c in range(C):
            for nd in np.nonzero(dirichlet(node_xyz, c))[0]:
                if int(nd) not in node_rows:
                    constraints[int(nd) * C + c] = []
        # AffineConstraints::close() resolves chains: a master that is itself constrained to zero drops out of
        # the hanging-node row (a hanging node on the wall ends up as a zero row)
        for dof in list(constraints.keys()):
            row = constraints[dof]
            if row:
                constraints[dof] = [(m, w) for m, w in row if not (m in constraints and not constraints[m])]
    return mesh


def cylinder_shell(shape, degree, *, r_inner=0.05, r_outer=0.5, length=0.41,
                   mapping_degree=None, no_slip=True, **kw):
    """O-grid around a cylinder (r, theta periodic[, z]) -- curved cells, like the
    blocks next to the cylinder in include/grid_cylinder.h:22-90,154-191."""
    dim = len(shape)
    mapping_degree = degree if mapping_degree is None else mapping_degree

    def deform(x):
        r = r_inner + (r_outer - r_inner) * x[..., 0] ** 1.5  # graded towards the wall
        th = 2.0 * np.pi * x[..., 1]
        out = np.empty_like(x)
        out[..., 0] = r * np.cos(th)
        out[..., 1] = r * np.sin(th)
        if dim == 3:
            out[..., 2] = length * x[..., 2]
        return out

    periodic = (False, True) + ((False,) if dim == 3 else ())

    def dirichlet(ref, c):
        if c == dim:
            return np.zeros(len(ref), dtype=bool)
        m = np.abs(ref[:, 0]) < 1e-12          # cylinder surface
        m |= np.abs(ref[:, 0] - 1.0) < 1e-12   # outer wall
        return m

    return structured_mesh(dim, shape, degree, deform=deform, mapping_degree=mapping_degree,
                           periodic=periodic, dirichlet=dirichlet if no_slip else None, **kw)


def add_random_constraints(mesh: Mesh, n_weighted: int, n_zero: int, seed: int = 0,
                           owned_only: bool = True):
    """Attach synthetic affine constraints x_i = sum_j w_ij x_j (hanging-node-like rows
    with 2-4 masters) and zero constraints, masters always unconstrained."""
    rng = np.random.default_rng(seed)
    hi = mesh.n_owned if owned_only else mesh.n_dofs
    cand = rng.permutation(hi)
    already = set(mesh.constraints.keys())
    cand = [int(c) for c in cand if int(c) not in already]
    chosen = cand[: n_weighted + n_zero]
    cset = set(chosen) | already
    free = np.array([i for i in range(hi) if i not in cset], dtype=np.int64)
    for t, dof in enumerate(chosen):
        if t < n_weighted:
            m = rng.choice(free, size=int(rng.integers(2, 5)), replace=False)
            w = rng.uniform(-0.5, 1.0, size=len(m))
            mesh.constraints[dof] = [(int(a), float(b)) for a, b in zip(m, w)]
        else:
            mesh.constraints[dof] = []
    return mesh


def general_geometry(mesh: Mesh, n_q_1d: int | None = None):
    """J^{-T} and JxW at the quadrature points from the mapping support points
    (what MatrixFree's MappingInfo stores for 'general' cells).

    Returns inv_jac[k, q, e, j] = (J^-1)_{e j} and jxw[k, q]."""
    dim, kdeg = mesh.dim, mesh.mapping_degree
    nq = (mesh.degree + 1) if n_q_1d is None else n_q_1d
    xg, wg = np.polynomial.legendre.leggauss(nq)
    xg = 0.5 * (xg + 1.0)
    wg = 0.5 * wg
    nodes = gauss_lobatto_points(kdeg)

    def lag(x):
        m = len(nodes)
        V = np.ones((len(x), m))
        D = np.zeros((len(x), m))
        for i in range(m):
            for a in range(m):
                if a != i:
                    V[:, i] *= (x - nodes[a]) / (nodes[i] - nodes[a])
            for l in range(m):
                if l == i:
                    continue
                t = np.ones(len(x)) / (nodes[i] - nodes[l])
                for a in range(m):
                    if a != i and a != l:
                        t *= (x - nodes[a]) / (nodes[i] - nodes[a])
                D[:, i] += t
        return V, D

    V, D = lag(xg)
    Ts = []
    for e in range(dim):
        mats = [V] * dim
        mats[e] = D
        T = mats[0]
        for mtx in mats[1:]:
            T = np.kron(mtx, T)
        Ts.append(T)
    w = wg
    for _ in range(dim - 1):
        w = np.kron(wg, w)
    lib = _native()
    if lib is not None:
        T = np.ascontiguousarray(np.stack(Ts), dtype=np.float64)
        pts = np.ascontiguousarray(mesh.cell_points, dtype=np.float64)
        inv_jac = np.empty((mesh.n_cells, nq ** dim, dim, dim))
        jxw = np.empty((mesh.n_cells, nq ** dim))
        wq = np.ascontiguousarray(w, dtype=np.float64)
        rc = lib.glsm_general_geometry(dim, mesh.n_cells, nq ** dim, pts.shape[1], _ptr(T), _ptr(wq), _ptr(pts),
                                       _ptr(inv_jac), _ptr(jxw))
        if rc < 0:
            raise RuntimeError("glsm_general_geometry failed")
        return inv_jac, jxw
    J = np.zeros((mesh.n_cells, nq ** dim, dim, dim))
    for e in range(dim):
        J[:, :, :, e] = np.einsum("qm,kmi->kqi", Ts[e], mesh.cell_points)
    return np.linalg.inv(J), np.linalg.det(J) * w[None, :]


def _lagrange_1d(nodes, x):
    """values and derivatives [len(x), len(nodes)] of the Lagrange basis on `nodes`"""
    m = len(nodes)
    V = np.ones((len(x), m))
    D = np.zeros((len(x), m))
    for i in range(m):
        for a in range(m):
            if a != i:
                V[:, i] *= (x - nodes[a]) / (nodes[i] - nodes[a])
        for l in range(m):
            if l == i:
                continue
            t = np.ones(len(x)) / (nodes[i] - nodes[l])
            for a in range(m):
                if a != i and a != l:
                    t *= (x - nodes[a]) / (nodes[i] - nodes[a])
            D[:, i] += t
    return V, D


def boundary_faces(mesh: Mesh, kind_of_boundary_id, target_velocity=None):
    """Boundary faces of a structured block that carry outflow terms (operator_ns.cc:1195-1301) and what
    MatrixFree's face MappingInfo holds for them (mapping_update_flags_boundary_faces, operator_ns.cc:113-117).

    kind_of_boundary_id: {boundary id: 1 (all_outflow_bcs_cut) | 2 (all_outflow_bcs_nitsche)}, boundary ids as
    GridGenerator colorises a block: 2 * direction + side.  target_velocity: callable(points[..., dim]) ->
    [..., dim] for the Nitsche residual.  Face quadrature points: QGauss(degree + 1) in the tangential
    directions, ascending direction fastest.  Returns a dict of arrays:
      face_cell[f], face_no[f], face_kind[f], normal[f, q, dim], jxw[f, q], inv_jac[f, q, e, j],
      points[f, q, dim], target[f, q, dim] (zeros without target_velocity)"""
    dim, kdeg, nq = mesh.dim, mesh.mapping_degree, mesh.degree + 1
    xg, wg = np.polynomial.legendre.leggauss(nq)
    xg, wg = 0.5 * (xg + 1.0), 0.5 * wg
    nodes = gauss_lobatto_points(kdeg)
    Vt, Dt = _lagrange_1d(nodes, xg)
    cells, nos, kinds, normals, jxws, ijs, pts = [], [], [], [], [], [], []
    for bid, kind in sorted(kind_of_boundary_id.items()):
        e, side = bid // 2, bid % 2
        sel = np.nonzero(mesh.cell_coords[:, e] == (mesh.shape[e] - 1 if side else 0))[0]
        if len(sel) == 0:
            continue
        Vn, Dn = _lagrange_1d(nodes, np.array([float(side)]))
        vals = [Vn if a == e else Vt for a in range(dim)]

        def kron(mats):
            T = mats[0]
            for mtx in mats[1:]:
                T = np.kron(mtx, T)
            return T

        Nm = kron(vals)
        J = np.zeros((len(sel), Nm.shape[0], dim, dim))
        for a in range(dim):
            mats = list(vals)
            mats[a] = Dn if a == e else Dt
            J[:, :, :, a] = np.einsum("qm,kmi->kqi", kron(mats), mesh.cell_points[sel])
        w = kron([np.ones((1, 1)) if a == e else wg.reshape(-1, 1) for a in range(dim)]).reshape(-1)
        Jinv = np.linalg.inv(J)
        nref = np.zeros(dim)
        nref[e] = 1.0 if side else -1.0
        nn = np.einsum("kqej,e->kqj", Jinv, nref)
        ln = np.linalg.norm(nn, axis=2)
        cells.append(sel), nos.append(np.full(len(sel), bid)), kinds.append(np.full(len(sel), kind))
        normals.append(nn / ln[:, :, None]), jxws.append(np.abs(np.linalg.det(J)) * ln * w[None, :]), ijs.append(Jinv)
        pts.append(np.einsum("qm,kmi->kqi", Nm, mesh.cell_points[sel]))
    if not cells:
        return None
    out = dict(face_cell=np.concatenate(cells).astype(np.uint32), face_no=np.concatenate(nos).astype(np.uint32),
               face_kind=np.concatenate(kinds).astype(np.uint32), normal=np.concatenate(normals),
               jxw=np.concatenate(jxws), inv_jac=np.concatenate(ijs), points=np.concatenate(pts))
    out["target"] = target_velocity(out["points"]) if target_velocity is not None else np.zeros_like(out["normal"])
    return out
