"""Parity at sizes the CPU oracle cannot run: EVERY cell of a large structured hypercube checked against the
oracle on a small one, through periodicity (TEST INFRASTRUCTURE: imports oracle/; used by tests/ and by the
`parity_full_size` key of bench.py).

On a uniform Cartesian block without constraints the operator is translation-invariant cell by cell: if the
linearization point, the history and src repeat with a period of P cells, every cell of the block sees local data
that some cell of a small block with the same mesh size h and the same periodic fields sees, and the result at a
node depends only on the node's position inside the period and on whether it lies on the domain boundary.  So

    big block  (N cells per direction, N % P == 0)   <->   small block (M = 3 P cells per direction, same h)
    node i (per direction, 0 .. p N)                 ->    0            if i == 0
                                                           p M          if i == p N
                                                           p P + i % (p P)   otherwise (an interior period)

maps every dof of the big block to a dof of the small one with identical touching cells.  The fields are a random
table over one period (so nothing about them is smooth or symmetric), the small block is evaluated by the CPU
oracle, and the big block's result has to equal the mapped small result in all its entries: index arithmetic beyond
2^31 table elements, ring refills over hundreds of batches per CTA, the persistent grid's wraparound and the tail
batches are all inside the comparison.  Reference recipe being checked: performance.cc:16-87 (config P).
"""
from __future__ import annotations

import numpy as np

from dealii_ns_gls_b200 import mesh as gm


def _local_offsets(dim, n):
    """[n^dim, dim] local node offsets, x fastest (FEEvaluation's lexicographic order, mesh.py)"""
    loc = np.stack(np.meshgrid(*[np.arange(n)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)
    return loc[:, ::-1].copy()


class PeriodicFullSizeCheck:
    """Builds the periodic fields of a big hypercube block and compares a result vector of the big block with the
    oracle's result on the small block.  All big-block arrays live on `device` (torch); the small block is numpy."""

    def __init__(self, big: gm.Mesh, device, period_cells: int = 4, seed: int = 4321):
        import torch
        self.torch = torch
        dim, p = big.dim, big.degree
        assert big.geometry_type == 0 and not big.constraints and big.partition is None, \
            "uniform Cartesian single-rank block without constraints expected"
        N = tuple(int(s) for s in big.shape)
        P = int(period_cells)
        assert all(n % P == 0 and n >= 3 * P for n in N), "cells per direction must be a multiple of the period, >= 3 P"
        M = 3 * P
        self.dim, self.p, self.P, self.M, self.N = dim, p, P, M, N
        self.device = torch.device(device)
        h = np.asarray(big.extent, dtype=np.float64) / np.asarray(N, dtype=np.float64)
        self.small = gm.structured_mesh(dim, (M,) * dim, p, extent=h * M, origin=np.asarray(big.origin))
        C, n = dim + 1, p + 1
        n_loc = n ** dim
        loc = _local_offsets(dim, n)
        # ---- small block: (component, node) -> dof ----
        nn = p * M + 1
        s_dofs = np.asarray(self.small.cell_dofs, dtype=np.int64)
        s_cc = np.asarray(self.small.cell_coords, dtype=np.int64)
        lut = np.full((C,) + (nn,) * dim, -1, dtype=np.int64)
        for c in range(C):
            for l in range(n_loc):
                node = p * s_cc + loc[l][None, :]
                lut[(c,) + tuple(node[:, e] for e in range(dim))] = s_dofs[:, c * n_loc + l]
        assert (lut >= 0).all()
        self._lut = lut
        # ---- big block: dof -> dof of the small block ----
        pP = p * P
        lut_d = torch.from_numpy(lut).to(self.device)
        cc = torch.from_numpy(np.ascontiguousarray(big.cell_coords).astype(np.int32)).to(self.device).long()
        key = torch.full((big.n_dofs,), -1, dtype=torch.int32, device=self.device)
        assert big.n_dofs < 2 ** 31
        cd = np.ascontiguousarray(big.cell_dofs)
        cd = torch.from_numpy(cd.view(np.int32) if cd.dtype == np.uint32 else cd).to(self.device)  # one upload
        for c in range(C):
            for l in range(n_loc):
                idx = cd[:, c * n_loc + l].long()
                mapped = []
                for e in range(dim):
                    i = p * cc[:, e] + int(loc[l][e])
                    m = pP + i % pP
                    m = torch.where(i == 0, torch.zeros_like(m), m)
                    m = torch.where(i == p * N[e], torch.full_like(m, p * M), m)
                    mapped.append(m)
                key[idx] = lut_d[(c,) + tuple(mapped)].to(torch.int32)
        del cd
        assert int(key.min()) >= 0
        self.key = key.long()
        self.rng = np.random.default_rng(seed)
        self.n_big = int(big.n_dofs)
        self.n_cells_big = int(big.n_cells)

    # ---- periodic fields ----
    def field(self):
        """One random periodic field: (numpy vector of the small block, torch float64 vector of the big block)"""
        dim, p, C = self.dim, self.p, self.dim + 1
        pP = p * self.P
        T = self.rng.uniform(-1.0, 1.0, (C,) + (pP,) * dim)
        node = np.stack(np.meshgrid(*[np.arange(p * self.M + 1) % pP] * dim, indexing="ij"), axis=0)
        small = np.empty(self.small.n_dofs)
        for c in range(C):
            small[self._lut[c].reshape(-1)] = T[(c,) + tuple(node[e].reshape(-1) for e in range(dim))]
        big = self.torch.from_numpy(small).to(self.device)[self.key]
        return small, big

    def compare(self, got_big, ref_small):
        """got_big: torch vector of the big block (any float dtype); ref_small: numpy result on the small block"""
        torch = self.torch
        ref = torch.from_numpy(np.asarray(ref_small, dtype=np.float64)).to(self.device)[self.key]
        got = got_big.double()
        diff = got - ref
        rel = float(torch.linalg.vector_norm(diff) / torch.linalg.vector_norm(ref))
        worst = float(diff.abs().max() / ref.abs().max())
        return {"rel_l2_all_rows": rel, "max_abs_over_max_ref": worst, "n_dofs": self.n_big,
                "n_cells": self.n_cells_big, "period_cells": self.P, "oracle_cells": int(self.small.n_cells)}


def oracle_on_small(chk: PeriodicFullSizeCheck, *, lin, src, hist=None, nu, c1, c2, weights, dt, ctd, cell_wise,
                    order=2):
    """vmult of the CPU oracle (numpy restatement, Newton branch) on the small block"""
    from oracle import gls_oracle as go
    m = chk.small
    ora = go.OracleOperator(dim=m.dim, degree=m.degree, cell_dofs=m.cell_dofs, n_dofs=m.n_dofs,
                            cell_points=m.cell_points, mapping_degree=m.mapping_degree, constraints=m.constraints,
                            nu=nu, c1=c1, c2=c2, theta=1.0, order=order, consider_time_derivative=ctd,
                            increment_form=True, cell_wise_stabilization=cell_wise, path="sumfac")
    if hist is not None:
        ora.set_previous_solution(hist, weights)
    ora.set_linearization_point(lin, dt)
    return ora.vmult(src, weights[0])
