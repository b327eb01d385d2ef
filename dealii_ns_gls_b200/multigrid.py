"""Host-side mirror of the reference's geometric multigrid preconditioner around the device level operators.

    PreconditionerGMG<dim>                (include/multigrid.h:60-141, include/multigrid.cc:178-590)
      = deal.II PreconditionMG( Multigrid( mg::Matrix(level operators), coarse solver,
                                           MGTransferGlobalCoarsening, MGSmootherPrecondition x 2 ) )

The reference keeps all level vectors on the host and calls the level operators' vmult from deal.II's
Multigrid; here the level vectors live on the device and every step of the V-cycle is a kernel of
libglsb200.so enqueued on the current stream: relaxation sweeps
(glsb_relaxation_*), residual (glsb_vmult + glsb_vec_axpby), restriction / prolongation (glsb_transfer_*),
the coarse solve as a dense device matrix-vector product with the inverse of the level-0 matrix
(glsb_dense_apply), and the double <-> float copies around the cycle (glsb_vec_convert).  Nothing in a
V-cycle synchronises with the host.  The class and method names are deal.II's / the reference's.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .mesh import Mesh, child_cells
from .operator import AffineConstraints, _torch_dtype
from .smoother import PreconditionRelaxation


def _type_of(t: torch.Tensor):
    return L.GLSB_F64 if t.dtype == torch.float64 else L.GLSB_F32


def _ptr(t: torch.Tensor):
    return C.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class DeviceVectorOps:
    """The vector operations deal.II's Multigrid / SolverGMRES do on LinearAlgebra::distributed::Vector,
    forwarded to the glsb_vec_* kernels."""

    # distributed vectors: set by the driver of a multi-rank run to a function that sums a small float64 device
    # tensor over the ranks in place (NCCL all-reduce: Utilities::MPI::sum of the local inner products, what
    # LinearAlgebra::distributed::Vector::operator* does).  Vectors carry no ghost values between operator
    # calls, so the local kernels may run over the whole local array.
    allreduce_sum = None

    def __init__(self):
        self._lib = L.load()

    def _chk(self, rc, what):
        if rc != 0:
            raise L.GlsbError(f"{what} failed (rc={rc})")

    def axpby(self, y, a, x, b):
        """y = a x + b y"""
        self._chk(self._lib.glsb_vec_axpby(_ptr(y), float(a), _ptr(x), float(b), y.numel(), _type_of(y), _stream(y)),
                  "vec_axpby")

    def convert(self, dst, src):
        self._chk(self._lib.glsb_vec_convert(_ptr(dst), _type_of(dst), _ptr(src), _type_of(src), dst.numel(),
                                             _stream(dst)), "vec_convert")

    def multi_dot(self, out, V, k, w):
        """out[j] = V[j] . w for j < k; V is a contiguous [m, n] block, out a float64 device vector."""
        self._chk(self._lib.glsb_vec_multi_dot(_ptr(out), _ptr(V), V.stride(0), int(k), _ptr(w), w.numel(),
                                               _type_of(w), _stream(w)), "vec_multi_dot")
        if DeviceVectorOps.allreduce_sum is not None:
            DeviceVectorOps.allreduce_sum(out[:int(k)])

    def multi_axpy(self, w, V, k, coef, scale):
        """w += scale * sum_j coef[j] V[j]"""
        self._chk(self._lib.glsb_vec_multi_axpy(_ptr(w), _ptr(V), V.stride(0), int(k), _ptr(coef), float(scale),
                                                w.numel(), _type_of(w), _stream(w)), "vec_multi_axpy")

    def set_zero_indexed(self, v, idx):
        self._chk(self._lib.glsb_vec_set_zero_indexed(_ptr(v), _ptr(idx), idx.numel(), _type_of(v), _stream(v)),
                  "vec_set_zero_indexed")

    def dense_apply(self, y, A, x):
        self._chk(self._lib.glsb_dense_apply(_ptr(y), _ptr(A), _ptr(x), A.shape[0], A.shape[1], _type_of(y),
                                             _stream(y)), "dense_apply")


class MGTwoLevelTransfer:
    """deal.II's MGTwoLevelTransfer<dim, VectorType> between a level and the next coarser one, as the
    reference sets it up with ``reinit(dof_handler_fine, dof_handler_coarse[, constraints_fine,
    constraints_coarse])`` (main.cc:540-556).  Weights follow deal.II: 1 / (number of fine cells touching a
    dof), zero on constrained fine dofs."""

    def __init__(self):
        self._lib = L.load()
        self._t = None

    def reinit(self, mesh_fine: Mesh, mesh_coarse: Mesh, constraints_fine: AffineConstraints | None = None,
               constraints_coarse: AffineConstraints | None = None, number="float", device=None,
               op_fine=None, op_coarse=None):
        """op_fine / op_coarse: the level operators of a PARTITIONED hierarchy (their ghost exchange and pack /
        unpack kernels are used around the cell-wise transfer kernels).  Children live on the rank of their
        parent (the partitions of consecutive levels nest), so the exchange is the levels' own vector exchange:
        ghost import of the source, compress(add) of the destination."""
        if not torch.cuda.is_available():
            raise RuntimeError("MGTwoLevelTransfer needs a CUDA device; there is no CPU fallback")
        self.dtype = _torch_dtype(number)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_fine, self.n_coarse = mesh_fine.n_dofs, mesh_coarse.n_dofs
        self.op_fine = op_fine if (op_fine is not None and op_fine.exchange is not None) else None
        self.op_coarse = op_coarse if (op_coarse is not None and op_coarse.exchange is not None) else None
        self.n_owned_fine, self.n_owned_coarse = mesh_fine.n_owned, mesh_coarse.n_owned
        ndof = (mesh_coarse.dim + 1) * mesh_coarse.n_loc
        ch = child_cells(mesh_coarse, mesh_fine)
        fidx = np.ascontiguousarray(mesh_fine.cell_dofs[ch.reshape(-1)].reshape(mesh_coarse.n_cells, -1),
                                    dtype=np.uint32)
        cidx = np.ascontiguousarray(mesh_coarse.cell_dofs, dtype=np.uint32)  # replaced, never written in place
        row_ptr, ecol, ev = [0], [], []
        rows = constraints_coarse.rows if constraints_coarse is not None else {}
        if rows:
            cdofs = np.array(sorted(rows.keys()), dtype=np.int64)
            for d in cdofs:
                for m, w in rows[int(d)]:
                    ecol.append(m), ev.append(w)
                row_ptr.append(len(ecol))
            look = np.arange(mesh_coarse.n_dofs, dtype=np.uint32)
            look[cdofs] = np.arange(len(cdofs), dtype=np.uint32) | np.uint32(L.GLSB_CONSTRAINED_BIT)
            cidx = np.take(look, cidx)
        cidx = np.ascontiguousarray(cidx)
        touch = np.bincount(mesh_fine.cell_dofs.reshape(-1).astype(np.int64), minlength=mesh_fine.n_dofs)
        if self.op_fine is not None:
            # number of fine cells touching a dof over ALL ranks: compress(add) the local counts, send them back
            t = torch.tensor(touch, dtype=self.op_fine.dtype, device=self.device)
            self.op_fine.exchange.compress_add(self.op_fine, t)
            self.op_fine.exchange.update_ghost_values(self.op_fine, t)
            touch = np.rint(t.double().cpu().numpy()).astype(np.int64)
        weights = np.where(touch > 0, 1.0 / np.maximum(touch, 1), 0.0)
        if constraints_fine is not None and constraints_fine.rows:
            weights[np.fromiter(constraints_fine.rows.keys(), dtype=np.int64)] = 0.0
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        row_ptr = np.array(row_ptr, dtype=np.uint32)
        ecol = np.array(ecol, dtype=np.uint32)
        ev = np.array(ev, dtype=np.float64)

        def ptr(a):
            return a.ctypes.data_as(C.c_void_p) if a.size else None

        d = L.GlsbTransferDesc()
        d.abi_version, d.device = L.GLSB_ABI_VERSION, self.device.index or 0
        d.dim, d.degree = mesh_coarse.dim, mesh_coarse.degree
        d.number_type = L.GLSB_F64 if self.dtype == torch.float64 else L.GLSB_F32
        d.n_coarse_cells, d.n_fine_dofs, d.n_coarse_dofs = mesh_coarse.n_cells, self.n_fine, self.n_coarse
        d.coarse_dof_indices, d.fine_dof_indices = ptr(cidx), ptr(fidx)
        d.n_constraint_rows = len(row_ptr) - 1
        d.row_ptr, d.entry_col, d.entry_val = ptr(row_ptr), ptr(ecol), ptr(ev)
        d.weights = ptr(weights)
        h = C.c_void_p()
        self._release()
        with torch.cuda.device(self.device):
            rc = self._lib.glsb_transfer_create(C.byref(d), C.byref(h))
        if rc != 0:
            raise L.GlsbError("glsb_transfer_create failed: " + self._lib.glsb_transfer_last_error(None).decode())
        self._t = h
        return self

    def _release(self):
        if getattr(self, "_t", None):
            self._lib.glsb_transfer_destroy(self._t)
            self._t = None

    def __del__(self):
        self._release()

    def _call(self, fn, dst, src, n_dst, n_src, what):
        for t, n, name in ((dst, n_dst, "dst"), (src, n_src, "src")):
            if not (t.is_cuda and t.dtype == self.dtype and t.is_contiguous() and t.numel() == n):
                raise ValueError(f"{what}: {name} must be a contiguous CUDA {self.dtype} vector of length {n}")
        rc = fn(self._t, _ptr(dst), _ptr(src), _stream(dst))
        if rc != 0:
            raise L.GlsbError(f"{what} failed: " + self._lib.glsb_transfer_last_error(self._t).decode())

    @staticmethod
    def _import(op, vec):
        if op is not None:
            op.exchange.update_ghost_values(op, vec)

    @staticmethod
    def _drop_ghosts(op, vec):
        if op is not None and vec.numel() > op.n_owned:
            vec[op.n_owned:] = 0

    @staticmethod
    def _compress(op, vec):
        if op is not None:
            op.exchange.compress_add(op, vec)  # leaves the ghost block zeroed

    def prolongate_and_add(self, dst_fine, src_coarse):
        self._import(self.op_coarse, src_coarse)
        self._call(self._lib.glsb_transfer_prolongate_and_add, dst_fine, src_coarse, self.n_fine, self.n_coarse,
                   "prolongate_and_add")
        self._drop_ghosts(self.op_coarse, src_coarse)
        self._compress(self.op_fine, dst_fine)

    def restrict_and_add(self, dst_coarse, src_fine):
        self._import(self.op_fine, src_fine)
        self._call(self._lib.glsb_transfer_restrict_and_add, dst_coarse, src_fine, self.n_coarse, self.n_fine,
                   "restrict_and_add")
        self._drop_ghosts(self.op_fine, src_fine)
        self._compress(self.op_coarse, dst_coarse)

    def interpolate(self, dst_coarse, src_fine):
        self._import(self.op_fine, src_fine)
        self._call(self._lib.glsb_transfer_interpolate, dst_coarse, src_fine, self.n_coarse, self.n_fine,
                   "interpolate")
        self._drop_ghosts(self.op_fine, src_fine)
        self._drop_ghosts(self.op_coarse, dst_coarse)  # every owned coarse dof was written by a local cell


class MGTransferGlobalCoarsening:
    """deal.II's MGTransferGlobalCoarsening over MGLevelObject<MGTwoLevelTransfer> (main.cc:546-563):
    ``transfers[l]`` moves between level l and l - 1; ``initialize_dof_vector(l)`` makes a level vector."""

    def __init__(self, transfers: dict, initialize_dof_vector):
        self.transfers = dict(transfers)
        self._init_vec = initialize_dof_vector
        self._max = max(self.transfers) if self.transfers else 0
        self._min = min(self.transfers) - 1 if self.transfers else 0
        self._ops = DeviceVectorOps()

    def min_level(self):
        return self._min

    def max_level(self):
        return self._max

    def prolongate_and_add(self, to_level, dst, src):
        self.transfers[to_level].prolongate_and_add(dst, src)

    def restrict_and_add(self, from_level, dst, src):
        self.transfers[from_level].restrict_and_add(dst, src)

    def copy_to_mg(self, dst: dict, src: torch.Tensor):
        """finest level <- src (converted to the level number type), coarser levels zeroed"""
        for l in range(self._min, self._max + 1):
            if l not in dst or dst[l] is None:
                dst[l] = self._init_vec(l)
            elif l != self._max:
                dst[l].zero_()
        self._ops.convert(dst[self._max], src)

    def copy_from_mg(self, dst: torch.Tensor, src: dict):
        self._ops.convert(dst, src[self._max])

    def interpolate_to_mg(self, dst: dict, src: torch.Tensor):
        """main.cc:789-790, :825-827: the fine solution interpolated to every level"""
        for l in range(self._min, self._max + 1):
            if l not in dst or dst[l] is None:
                dst[l] = self._init_vec(l)
        self._ops.convert(dst[self._max], src)
        for l in range(self._max, self._min, -1):
            self.transfers[l].interpolate(dst[l - 1], dst[l])


class MGCoarseGridDirect:
    """The coarse-grid solver of the V-cycle for ``"gmg coarse grid solver": "direct"`` without iteration
    (multigrid.cc:419-425, :472-476: Trilinos SolverDirect on op[min_level]->get_system_matrix() behind the
    float <-> double shim MGCoarseGridApplyPreconditioner, :6-149).  The level-0 system matrix is obtained
    column by column from the level operator's own vmult on the device (what MatrixFreeTools::compute_matrix
    does cell-wise, operator_ns.cc:1407-1430); (glsb_get_system_matrix); it is inverted once on the host in double (LAPACK, the
    direct solver's factorisation) and applied on the device as a dense matrix-vector product."""

    def __init__(self, op, ops: DeviceVectorOps):
        self.op, self._ops = op, ops
        self.matrix = op.get_system_matrix()
        self._gid = None
        if op.exchange is not None:
            # partitioned coarse level: the rank-local matrices (local cells only) are summed into the global one
            # over the partition-independent ids of the dofs, and every rank factorises / applies the same small
            # dense matrix (the reference runs its direct solver on the distributed Trilinos matrix)
            import torch.distributed as dist
            mesh, dev = op.mesh, op.device
            ng, nl = mesh.n_global_dofs, mesh.n_dofs
            gid = torch.as_tensor(mesh.canonical_ids, dtype=torch.long, device=dev)
            cons = torch.zeros(nl, dtype=torch.bool, device=dev)
            if mesh.constraints:
                cons[torch.as_tensor(np.fromiter(mesh.constraints.keys(), dtype=np.int64), device=dev)] = True
            A = self.matrix.clone()
            A[cons, :] = 0   # identity rows are set once, globally, below
            A[:, cons] = 0   # constrained columns read 0 (zero-type rows)
            G = torch.zeros((ng, ng), dtype=torch.float64, device=dev)
            G.index_put_((gid[:, None].expand(nl, nl), gid[None, :].expand(nl, nl)), A, accumulate=True)
            gc = torch.zeros(ng, dtype=torch.float64, device=dev)
            gc[gid[cons]] = 1.0
            dist.all_reduce(G, group=op.exchange.group)
            dist.all_reduce(gc, group=op.exchange.group)
            ci = torch.nonzero(gc > 0).flatten()
            G[ci, ci] = 1.0
            self.matrix = G
            self._gid = gid[:op.n_owned]
            self._gsrc = torch.zeros(ng, dtype=op.dtype, device=dev)
            self._gdst = torch.zeros(ng, dtype=op.dtype, device=dev)
        # the factorisation of the direct solver, in double, once per initialize(): LAPACK on the host for the
        # few hundred dofs of a level-0 mesh, cuSOLVER on the device (torch.linalg.inv) above that -- a coarsest
        # level that every rank of an 8-GPU run holds cells on has 1 700 dofs (3-D Q2 channel, level 1)
        if self.matrix.shape[0] <= 512:
            self.inverse = torch.from_numpy(np.ascontiguousarray(np.linalg.inv(self.matrix.cpu().numpy()))).to(op.device)
        else:
            self.inverse = torch.linalg.inv(self.matrix).contiguous()

    def __call__(self, level, dst, src):
        if self._gid is None:
            self._ops.dense_apply(dst, self.inverse, src)
            return
        import torch.distributed as dist
        self._gsrc.zero_()
        self._gsrc[self._gid] = src[:self.op.n_owned]
        dist.all_reduce(self._gsrc, group=self.op.exchange.group)
        self._ops.dense_apply(self._gdst, self.inverse, self._gsrc)
        dst.zero_()
        dst[:self.op.n_owned] = self._gdst[self._gid]


class MGCoarseGridIdentity:
    """``"gmg coarse grid solver": "identity"`` (multigrid.cc:426-429)."""

    def __init__(self, ops):
        self._ops = ops

    def __call__(self, level, dst, src):
        self._ops.axpby(dst, 1.0, src, 0.0)


class Multigrid:
    """deal.II's Multigrid<VectorType> restricted to what the reference uses: V-cycle, the same smoother
    object before and after (multigrid.cc:534-540).  MGSmootherPrecondition semantics: pre-smoothing is
    ``apply`` (smoother.vmult from a zero guess), post-smoothing is ``smooth`` (u += P (rhs - A u))."""

    def __init__(self, matrices: dict, coarse, transfer: MGTransferGlobalCoarsening, smoothers: dict, min_level,
                 max_level):
        self.matrix, self.coarse, self.transfer, self.smoothers = matrices, coarse, transfer, smoothers
        self.minlevel, self.maxlevel = min_level, max_level
        self.defect, self.solution, self.t, self.d = {}, {}, {}, {}
        self._ops = DeviceVectorOps()

    def _reinit(self):
        for l in range(self.minlevel, self.maxlevel + 1):
            if l not in self.solution:
                self.solution[l] = self.matrix[l].initialize_dof_vector()
                self.t[l] = self.matrix[l].initialize_dof_vector()
                self.d[l] = self.matrix[l].initialize_dof_vector()
            else:
                self.solution[l].zero_()

    def cycle(self):
        self._reinit()
        self.level_v_step(self.maxlevel)

    def level_v_step(self, level):
        sol, rhs, t = self.solution[level], self.defect[level], self.t[level]
        if level == self.minlevel:
            self.coarse(level, sol, rhs)
            return
        # pre-smoothing: MGSmootherPrecondition::apply
        self.smoothers[level].vmult(sol, rhs)
        # residual t = rhs - A sol
        self.matrix[level].vmult(t, sol)
        self._ops.axpby(t, 1.0, rhs, -1.0)
        # restriction (the coarser defect was zeroed by copy_to_mg / the previous level)
        self.transfer.restrict_and_add(level, self.defect[level - 1], t)
        self.level_v_step(level - 1)
        # coarse-grid correction
        self.transfer.prolongate_and_add(level, sol, self.solution[level - 1])
        # post-smoothing: MGSmootherPrecondition::smooth, one step: u += P (rhs - A u)
        self.matrix[level].vmult(t, sol)
        self._ops.axpby(t, 1.0, rhs, -1.0)
        self.smoothers[level].vmult(self.d[level], t)
        self._ops.axpby(sol, 1.0, self.d[level], 1.0)


class PreconditionerGMGAdditionalData:
    """include/multigrid.h:24-57 (the parameters the device path honours)."""

    def __init__(self, **kw):
        self.smoothing_range = 20.0
        self.smoothing_n_iterations = 5
        self.smoothing_eig_cg_n_iterations = 20
        self.coarse_grid_solver = "direct"
        self.coarse_grid_iterate = False
        self.use_cuda_graph = False  # not a parameter of the reference: replay the V-cycle as one CUDA graph
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError(f"unknown parameter {k}")
            setattr(self, k, v)


class PreconditionerGMG:
    """include/multigrid.h:60-141.  ``op`` maps level -> level operator (MGNumber), ``transfer`` is the
    MGTransferGlobalCoarsening built WITH constraints (main.cc:552-563).  ``vmult`` takes vectors of the
    outer (double) number type like PreconditionerBase::vmult (preconditioner.h:19-21)."""

    def __init__(self, op: dict, transfer: MGTransferGlobalCoarsening, additional_data=None):
        self.op = dict(op)
        self.transfer = transfer
        self.additional_data = additional_data or PreconditionerGMGAdditionalData()
        self.mg = None
        self.n_vmult = 0
        self.use_cuda_graph = bool(self.additional_data.use_cuda_graph)
        self._graph = None

    def initialize(self):
        """multigrid.cc:248-590"""
        ad = self.additional_data
        lo, hi = self.transfer.min_level(), self.transfer.max_level()
        ops = DeviceVectorOps()
        self.smoothers = {}
        for level in range(lo, hi + 1):
            self.smoothers[level] = PreconditionRelaxation(
                self.op[level], None, relaxation=0.0, n_iterations=ad.smoothing_n_iterations,
                smoothing_range=ad.smoothing_range, eig_cg_n_iterations=ad.smoothing_eig_cg_n_iterations)
        for level in range(lo + 1, hi + 1):  # multigrid.cc:353-370 with compute_evs_n_levels == 0
            self.smoothers[level].estimate_eigenvalues()
        if ad.coarse_grid_iterate:
            raise NotImplementedError("coarse-grid GMRES around Trilinos AMG / ILU stays with the host path")
        if ad.coarse_grid_solver == "direct":
            coarse = MGCoarseGridDirect(self.op[lo], ops)
        elif ad.coarse_grid_solver == "identity":
            coarse = MGCoarseGridIdentity(ops)
        else:
            raise NotImplementedError(f"coarse grid solver {ad.coarse_grid_solver!r} (Trilinos) is host-side")
        self.coarse = coarse
        self.mg = Multigrid(self.op, coarse, self.transfer, self.smoothers, lo, hi)
        self._graph = None  # the smoothers' diagonals and relaxation parameters changed: capture again

    def vmult(self, dst: torch.Tensor, src: torch.Tensor):
        """PreconditionMG::vmult: copy_to_mg, one V-cycle, copy_from_mg (multigrid.cc:202-220).  With
        ``use_cuda_graph`` the whole sequence (~45 launches per level, none of which returns to the host) is
        captured once per initialize() into a CUDA graph and replayed: one launch per V-cycle.  Off by default:
        the smoother data changes with every initialize(), so the capture (a warm-up cycle, the capture pass and
        the instantiation, about three cycles' worth) has to be paid per preconditioner setup, and the channel
        runs apply only ~7 V-cycles per setup -- measured 52.6 against 28.3 ms of linear-solve time per step at
        4.3e6 DoFs, 14.9 against 12.4 ms at 5.6e5.  It pays for solves with dozens of V-cycles per setup."""
        if not self.use_cuda_graph:
            self._vcycle(dst, src)
        else:
            if self._graph is None:
                self._capture(src)
            self._g_src.copy_(src)
            self._graph.replay()
            dst.copy_(self._g_dst)
        self.n_vmult += 1

    def _vcycle(self, dst, src):
        self.transfer.copy_to_mg(self.mg.defect, src)
        self.mg.cycle()
        self.transfer.copy_from_mg(dst, self.mg.solution)

    def _capture(self, src):
        self._g_src, self._g_dst = torch.zeros_like(src), torch.zeros_like(src)
        self._g_src.copy_(src)
        # eager passes on a side stream first: every buffer of the cycle exists before the capture starts
        side = torch.cuda.Stream(device=src.device)
        side.wait_stream(torch.cuda.current_stream(src.device))
        with torch.cuda.stream(side):
            self._vcycle(self._g_dst, self._g_src)
            # capture_begin / capture_end directly: the torch.cuda.graph context runs gc.collect(), which costs tens
            # of milliseconds next to the mesh dictionaries of the host layer
            side.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            self._graph.capture_begin()
            self._vcycle(self._g_dst, self._g_src)
            self._graph.capture_end()
        torch.cuda.current_stream(src.device).wait_stream(side)

    def print_stats(self):
        pass
