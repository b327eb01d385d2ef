"""Ghost exchange of the distributed vectors: the B200 counterpart of what deal.II's
MatrixFree::cell_loop does implicitly around the cell work (operator_ns.cc:703-708):

    src.update_ghost_values()   owners -> ghosts (before the cells that touch ghosts)
    dst.compress(add)           ghost contributions -> owners, added

One process per GPU.  torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU
tests) only moves contiguous buffers; packing / unpack-add run as kernels of libglsb200.so
(glsb_pack_export / glsb_unpack_add).  The ghost block of a vector is contiguous and grouped
by owner, so ghost values are received in place and ghost contributions are sent from place.
vmult overlaps the exchange with the interior cells on a second stream."""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib as L


class GhostExchange:
    def __init__(self, partition, device, group=None):
        self.part = partition
        self.device = torch.device(device)
        self.group = group
        self.n_owned = partition.n_owned
        self.n_export = int(sum(len(s[1]) for s in partition.send))
        self._bufs = {}
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        # offsets of each neighbour's slice in the packed export buffer
        self.send_slices = []
        o = 0
        for r, idx in partition.send:
            self.send_slices.append((r, o, len(idx)))
            o += len(idx)
        self._export_idx_cpu = None

    # ---- buffers ----
    def _buf(self, name, dtype):
        key = (name, dtype)
        if key not in self._bufs:
            self._bufs[key] = torch.empty(max(self.n_export, 1), dtype=dtype, device=self.device)
        return self._bufs[key]

    # ---- pack / unpack: library kernels on CUDA, index ops on CPU tensors (tests) ----
    def _pack(self, op, buf, vec):
        if vec.is_cuda:
            s = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            L.check(op._lib, op._op, op._lib.glsb_pack_export(op._op, C.c_void_p(buf.data_ptr()),
                                                              C.c_void_p(vec.data_ptr()), s), "pack_export")
        else:
            buf[: self.n_export] = vec[self._export_idx(vec)]

    def _unpack_add(self, op, vec, buf):
        if vec.is_cuda:
            s = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            L.check(op._lib, op._op, op._lib.glsb_unpack_add(op._op, C.c_void_p(vec.data_ptr()),
                                                             C.c_void_p(buf.data_ptr()), s), "unpack_add")
        else:
            vec.index_add_(0, self._export_idx(vec), buf[: self.n_export])

    def _export_idx(self, vec):
        if self._export_idx_cpu is None:
            import numpy as np
            idx = np.concatenate([s[1] for s in self.part.send]) if self.part.send else np.zeros(0, dtype=np.int64)
            self._export_idx_cpu = torch.as_tensor(idx, dtype=torch.long)
        return self._export_idx_cpu.to(vec.device)

    # ---- point-to-point rounds ----
    def _exchange(self, sends, recvs):
        """sends / recvs: lists of (peer_rank, tensor_view)."""
        ops = [dist.P2POp(dist.irecv, t, r, group=self.group) for r, t in recvs]
        ops += [dist.P2POp(dist.isend, t, r, group=self.group) for r, t in sends]
        if not ops:
            return
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def update_ghost_values(self, op, vec):
        """owners -> ghosts; vec is [owned | ghost]."""
        buf = self._buf("send", vec.dtype)
        self._pack(op, buf, vec)
        sends = [(r, buf[o:o + n]) for r, o, n in self.send_slices]
        recvs = [(r, vec[self.n_owned + o: self.n_owned + o + n]) for r, o, n in self.part.recv]
        self._exchange(sends, recvs)

    def compress_add(self, op, vec):
        """ghost contributions -> owners (added); ghosts are zeroed afterwards like compress() does."""
        buf = self._buf("recv", vec.dtype)
        sends = [(r, vec[self.n_owned + o: self.n_owned + o + n]) for r, o, n in self.part.recv]
        recvs = [(r, buf[o:o + n]) for r, o, n in self.send_slices]
        self._exchange(sends, recvs)
        self._unpack_add(op, vec, buf)
        if vec.numel() > self.n_owned:
            vec[self.n_owned:] = 0

    def allreduce_max(self, val: float) -> float:
        t = torch.tensor([val], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t[0])

    def allreduce_sum(self, vals) -> list:
        """MPI::sum of a few scalars (norms / inner products of distributed vectors)"""
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return [float(x) for x in t]

    # ---- vmult ----
    SM_RESERVE = 8  # multiprocessors left to the NCCL send/recv kernels next to the persistent cell kernel
    # exchanged bytes per vmult above which the exchanges are hidden behind the interior cells.  Below it
    # (one node plane of a 160^3-cell slab is 3.3 MB: ~50 us on NVLink against a 4 ms cell kernel) the
    # plain sequence import -> all cells in one launch -> compress is faster: the overlapped schedule
    # pays for three launches of the persistent kernel, the reserved SMs, and NCCL kernels that only get
    # scheduled when cell CTAs retire (measured on 2 B200: 5.5 ms overlapped, 4.3 ms in sequence).
    OVERLAP_MIN_BYTES = int(os.environ.get("GLSB_OVERLAP_MIN_BYTES", str(64 << 20)))

    def vmult(self, op, dst, src, weight, kernel_events=None):
        """cell_loop(..., zero_dst = true) of operator_ns.cc:703-708 on one rank."""
        if self.n_export * src.element_size() >= self.OVERLAP_MIN_BYTES or \
                sum(n for _, _, n in self.part.recv) * src.element_size() >= self.OVERLAP_MIN_BYTES:
            return self._vmult_overlapped(op, dst, src, weight, kernel_events)
        lib, h = op._lib, op._op
        s = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        d, x = C.c_void_p(dst.data_ptr()), C.c_void_p(src.data_ptr())
        L.check(lib, h, lib.glsb_vmult_begin(h, d, s), "vmult")
        self.update_ghost_values(op, src)
        if kernel_events is not None:
            kernel_events[0].record()
        L.check(lib, h, lib.glsb_vmult_cells(h, d, x, weight, L.GLSB_CELLS_ALL, s), "vmult")
        if kernel_events is not None:
            kernel_events[1].record()
        self.compress_add(op, dst)  # also zeroes the ghost block of dst
        L.check(lib, h, lib.glsb_vmult_finish(h, d, x, s), "vmult")
        if src.numel() > self.n_owned:
            src[self.n_owned:] = 0  # like cell_loop, leave src without ghost values

    def _vmult_overlapped(self, op, dst, src, weight, kernel_events=None):
        """Both exchanges hidden behind the interior cells:

            compute stream : zero dst | interior half A | boundary cells | interior half B | unpack-add, finish
            comm stream    :   pack, ghost import ------^        ghost contributions -> owners ---^
        """
        lib, h = op._lib, op._op
        cur = torch.cuda.current_stream(self.device)
        s = C.c_void_p(cur.cuda_stream)
        d, x = C.c_void_p(dst.data_ptr()), C.c_void_p(src.data_ptr())
        lib.glsb_set_sm_reserve(h, self.SM_RESERVE)
        L.check(lib, h, lib.glsb_vmult_begin(h, d, s), "vmult")
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            self.update_ghost_values(op, src)
        if kernel_events is not None:
            kernel_events[0].record()
        L.check(lib, h, lib.glsb_vmult_cells_part(h, d, x, weight, L.GLSB_CELLS_INTERIOR, 0, 2, s), "vmult")
        cur.wait_stream(self.comm_stream)
        L.check(lib, h, lib.glsb_vmult_cells(h, d, x, weight, L.GLSB_CELLS_BOUNDARY, s), "vmult")
        self.comm_stream.wait_stream(cur)
        buf = self._buf("recv", dst.dtype)
        with torch.cuda.stream(self.comm_stream):
            sends = [(r, dst[self.n_owned + o: self.n_owned + o + n]) for r, o, n in self.part.recv]
            recvs = [(r, buf[o:o + n]) for r, o, n in self.send_slices]
            self._exchange(sends, recvs)
        L.check(lib, h, lib.glsb_vmult_cells_part(h, d, x, weight, L.GLSB_CELLS_INTERIOR, 1, 2, s), "vmult")
        if kernel_events is not None:
            kernel_events[1].record()
        cur.wait_stream(self.comm_stream)
        self._unpack_add(op, dst, buf)
        L.check(lib, h, lib.glsb_vmult_finish(h, d, x, s), "vmult")
        # like cell_loop, leave both vectors without ghost values
        if src.numel() > self.n_owned:
            src[self.n_owned:] = 0
            dst[self.n_owned:] = 0
        lib.glsb_set_sm_reserve(h, 0)
