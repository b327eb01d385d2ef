"""Host-side mirror of the reference's operator interface for the hot path.

``NavierStokesOperator`` has the methods of ``OperatorBase<Number>``
(include/operator_base.h:13-73) as implemented by ``NavierStokesOperator<dim,Number>``
(include/operator_ns.h:17-189): m, vmult, Tvmult, compute_inverse_diagonal,
set_previous_solution, set_linearization_point, evaluate_rhs, evaluate_residual,
invalidate_system, initialize_dof_vector, get_max_u, get_constraints.  Every
method forwards to libglsb200.so through the C ABI of include/glsb200.h; torch
is used only for device buffers and streams.  Vectors are 1-D CUDA tensors of
length n_owned + n_ghost (the layout of LinearAlgebra::distributed::Vector).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .mesh import Mesh, general_geometry


def _torch_dtype(number):
    if number in ("double", "float64", torch.float64, np.float64):
        return torch.float64
    if number in ("float", "float32", torch.float32, np.float32):
        return torch.float32
    raise ValueError(f"unsupported Number {number!r}")


class AffineConstraints:
    """The slice of dealii::AffineConstraints the hot path touches: homogeneous rows
    x_i = sum_j w_ij x_j plus optional inhomogeneities (for constraints_inhomogeneous,
    operator_ns.cc:655-656)."""

    def __init__(self, rows=None, inhomogeneities=None):
        self.rows = dict(rows or {})
        self.inhomogeneities = dict(inhomogeneities or {})
        self._dev = None

    def is_constrained(self, i):
        return i in self.rows

    def n_constraints(self):
        return len(self.rows)

    def _device_tables(self, device, dtype):
        key = (str(device), dtype)
        if self._dev is None or self._dev[0] != key:
            dofs = sorted(self.rows.keys())
            r, c, w = [], [], []
            for k, d in enumerate(dofs):
                for m, ww in self.rows[d]:
                    r.append(k), c.append(m), w.append(ww)
            t = dict(
                dofs=torch.tensor(dofs, dtype=torch.long, device=device),
                inhom=torch.tensor([self.inhomogeneities.get(d, 0.0) for d in dofs], dtype=dtype, device=device),
                r=torch.tensor(r, dtype=torch.long, device=device),
                c=torch.tensor(c, dtype=torch.long, device=device),
                w=torch.tensor(w, dtype=dtype, device=device))
            self._dev = (key, t)
        return self._dev[1]

    def distribute(self, vec: torch.Tensor):
        """AffineConstraints::distribute: constrained entries <- inhomogeneity + sum w * master."""
        if not self.rows:
            return
        t = self._device_tables(vec.device, vec.dtype)
        vals = t["inhom"].clone()
        if t["r"].numel():
            vals.index_add_(0, t["r"], t["w"] * vec[t["c"]])
        vec[t["dofs"]] = vals

    def set_zero(self, vec: torch.Tensor):
        if not self.rows:
            return
        t = self._device_tables(vec.device, vec.dtype)
        vec[t["dofs"]] = 0


def build_desc(mesh: Mesh, *, number, nu, c_1, c_2, theta, time_order, consider_time_derivative,
               increment_form, cell_wise_stabilization, device_index=0):
    """Flatten a Mesh into the glsb_desc the C ABI takes. Returns (desc, keepalive)."""
    keep = {}
    C_ = mesh.dim + 1
    ndof = C_ * mesh.n_loc
    idx = np.ascontiguousarray(mesh.cell_dofs, dtype=np.uint32)  # never written in place below
    assert idx.shape == (mesh.n_cells, ndof)

    # constraint rows
    cdofs = np.array(sorted(mesh.constraints.keys()), dtype=np.int64)
    row_ptr = [0]
    ecol, ev = [], []
    for d in cdofs:
        for m, w in mesh.constraints[int(d)]:
            ecol.append(m)
            ev.append(w)
        row_ptr.append(len(ecol))
    if len(cdofs):
        # one gather through a lookup that is the identity on unconstrained dofs and GLSB_CONSTRAINED_BIT | row on
        # constrained ones
        look = np.arange(mesh.n_dofs, dtype=np.uint32)
        look[cdofs] = np.arange(len(cdofs), dtype=np.uint32) | np.uint32(L.GLSB_CONSTRAINED_BIT)
        idx = np.take(look, idx)
    keep["idx"] = np.ascontiguousarray(idx)
    keep["row_dof"] = cdofs.astype(np.uint32)
    keep["row_ptr"] = np.array(row_ptr, dtype=np.uint32)
    keep["ecol"] = np.array(ecol, dtype=np.uint32)
    keep["eval"] = np.array(ev, dtype=np.float64)
    keep["cidx"] = cdofs[cdofs < mesh.n_owned].astype(np.uint32)  # get_constrained_dofs(): owned rows

    if mesh.geometry_type == L.GLSB_GEOM_CARTESIAN:
        keep["inv_jac"] = np.ascontiguousarray(mesh.cart_inv_jac, dtype=np.float64)
        keep["jxw"] = np.ascontiguousarray(mesh.cart_det, dtype=np.float64)
    else:
        ij, jxw = general_geometry(mesh)
        keep["inv_jac"] = np.ascontiguousarray(ij, dtype=np.float64)
        keep["jxw"] = np.ascontiguousarray(jxw, dtype=np.float64)
    keep["h_min"] = np.ascontiguousarray(mesh.cell_h_min, dtype=np.float64)
    keep["measure"] = np.ascontiguousarray(mesh.cell_measure, dtype=np.float64)

    export = np.zeros(0, dtype=np.uint32)
    if mesh.partition is not None and mesh.partition.send:
        export = np.concatenate([s[1] for s in mesh.partition.send]).astype(np.uint32)
    keep["export"] = export

    def ptr(a):
        return a.ctypes.data_as(C.c_void_p) if a.size else None

    d = L.GlsbDesc()
    d.abi_version = L.GLSB_ABI_VERSION
    d.device = device_index
    d.dim, d.degree = mesh.dim, mesh.degree
    d.number_type = L.GLSB_F64 if _torch_dtype(number) == torch.float64 else L.GLSB_F32
    d.increment_form = int(increment_form)
    d.consider_time_derivative = int(consider_time_derivative)
    d.cell_wise_stabilization = int(cell_wise_stabilization)
    d.time_order = int(time_order)
    d.nu, d.c1, d.c2, d.theta = float(nu), float(c_1), float(c_2), float(theta)
    d.n_cells, d.n_owned, d.n_ghost = mesh.n_cells, mesh.n_owned, mesh.n_dofs - mesh.n_owned
    d.dof_indices = ptr(keep["idx"])
    d.n_constraint_rows = len(cdofs)
    d.row_dof, d.row_ptr = ptr(keep["row_dof"]), ptr(keep["row_ptr"])
    d.entry_col, d.entry_val = ptr(keep["ecol"]), ptr(keep["eval"])
    d.n_constrained_indices = len(keep["cidx"])
    d.constrained_indices = ptr(keep["cidx"])
    d.geometry_type = mesh.geometry_type
    d.inv_jac, d.jxw = ptr(keep["inv_jac"]), ptr(keep["jxw"])
    d.cell_h_min, d.cell_measure = ptr(keep["h_min"]), ptr(keep["measure"])
    d.n_export = len(export)
    d.export_indices = ptr(keep["export"])
    edge = np.ascontiguousarray(getattr(mesh, "edge_constrained_indices", np.zeros(0)), dtype=np.uint32)
    keep["edge"] = edge
    d.n_edge_constrained_indices = len(edge)
    d.edge_constrained_indices = ptr(edge)
    d.has_edge_constrained_indices = int(getattr(mesh, "has_edge_constrained_indices", len(edge) > 0))
    # boundary faces with outflow terms: mesh.outflow_faces = mesh.boundary_faces(...) (operator_ns.h:33-35)
    faces = getattr(mesh, "outflow_faces", None)
    if faces is not None and len(faces["face_cell"]):
        for k_, dt_ in (("face_cell", np.uint32), ("face_no", np.uint32), ("face_kind", np.uint32),
                        ("normal", np.float64), ("jxw", np.float64), ("inv_jac", np.float64), ("target", np.float64)):
            keep["f_" + k_] = np.ascontiguousarray(faces[k_], dtype=dt_)
        d.n_outflow_faces = len(keep["f_face_cell"])
        d.face_cell, d.face_no, d.face_kind = ptr(keep["f_face_cell"]), ptr(keep["f_face_no"]), ptr(keep["f_face_kind"])
        d.face_normal, d.face_jxw, d.face_inv_jac = ptr(keep["f_normal"]), ptr(keep["f_jxw"]), ptr(keep["f_inv_jac"])
        d.face_target_velocity = ptr(keep["f_target"])
    return d, keep


class NavierStokesOperator:
    """Drop-in for NavierStokesOperator<dim, Number> on the hot path (operator_ns.h:24-92).

    Constructor arguments follow the reference's (operator_ns.h:24-41): the
    mapping / dof_handler / constraints_homogeneous / quadrature quadruple is the
    ``mesh`` description, ``constraints_inhomogeneous`` is kept by reference.
    """

    def __init__(self, mesh: Mesh, constraints_inhomogeneous: AffineConstraints | None,
                 nu, c_1, c_2, time_integrator_data, consider_time_derivative, increment_form,
                 cell_wise_stabilization, number="double", device=None, exchange=None):
        if not torch.cuda.is_available():
            raise RuntimeError("NavierStokesOperator needs a CUDA device; there is no CPU fallback")
        self._lib = L.load()
        self.mesh = mesh
        self.dtype = _torch_dtype(number)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.constraints_inhomogeneous = constraints_inhomogeneous or AffineConstraints()
        self.constraints_homogeneous = AffineConstraints(mesh.constraints)
        self.time_integrator_data = time_integrator_data
        self.exchange = exchange
        self.increment_form = bool(increment_form)
        desc, keep = build_desc(mesh, number=number, nu=nu, c_1=c_1, c_2=c_2,
                                theta=time_integrator_data.get_theta(),
                                time_order=time_integrator_data.get_order(),
                                consider_time_derivative=consider_time_derivative,
                                increment_form=increment_form,
                                cell_wise_stabilization=cell_wise_stabilization,
                                device_index=self.device.index or 0)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self._lib.glsb_create(C.byref(desc), C.byref(h))
        if rc != 0:
            raise L.GlsbError("glsb_create failed: " + self._lib.glsb_last_error(None).decode())
        self._op = h
        del keep
        self.n_local = mesh.n_dofs
        self.n_owned = mesh.n_owned
        self._pinned = {}

    def __del__(self):
        op = getattr(self, "_op", None)
        if op:
            self._lib.glsb_destroy(op)
            self._op = None

    # ---- helpers ----
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, rc, what):
        L.check(self._lib, self._op, rc, what)

    def _vec(self, t: torch.Tensor, name):
        if not (t.is_cuda and t.dtype == self.dtype and t.is_contiguous() and t.numel() == self.n_local):
            raise ValueError(f"{name}: expected a contiguous CUDA {self.dtype} vector of length {self.n_local}")
        return C.c_void_p(t.data_ptr())

    def _update_ghost_values(self, vec):
        if self.exchange is not None:
            self.exchange.update_ghost_values(self, vec)

    def _compress_add(self, vec):
        if self.exchange is not None:
            self.exchange.compress_add(self, vec)

    def _zero_ghosts(self, vec):
        """vectors leave every call without ghost values, like the reference's (operator_ns.cc:540, :591): the
        vector operations of the solvers (dots, norms, axpys over the whole local array) rely on it"""
        if self.exchange is not None and vec.numel() > self.n_owned:
            vec[self.n_owned:] = 0

    # ---- OperatorBase interface ----
    def m(self):
        return self.mesh.n_global_dofs

    def get_constraints(self):
        return self.constraints_homogeneous

    def initialize_dof_vector(self):
        return torch.zeros(self.n_local, dtype=self.dtype, device=self.device)

    def invalidate_system(self):
        self._chk(self._lib.glsb_invalidate_system(self._op), "invalidate_system")

    def vmult(self, dst: torch.Tensor, src: torch.Tensor, kernel_events=None):
        """operator_ns.cc:684-732.  kernel_events = (start, end) torch.cuda.Event pair recorded
        around the cell kernel on the launching stream (bench.py's roofline timing)."""
        w = self.time_integrator_data.get_primary_weight()
        s = self._stream()
        if self.exchange is not None:
            d, x = self._vec(dst, "dst"), self._vec(src, "src")
            self._chk(self._lib.glsb_edge_begin(self._op, x, s), "vmult")
            self.exchange.vmult(self, dst, src, w, kernel_events)
            self._chk(self._lib.glsb_edge_finish(self._op, d, x, s), "vmult")
        elif kernel_events is None:
            self._chk(self._lib.glsb_vmult(self._op, self._vec(dst, "dst"), self._vec(src, "src"), w, s), "vmult")
        else:
            d, x = self._vec(dst, "dst"), self._vec(src, "src")
            self._chk(self._lib.glsb_vmult_begin(self._op, d, s), "vmult")
            kernel_events[0].record()
            self._chk(self._lib.glsb_vmult_cells(self._op, d, x, w, L.GLSB_CELLS_ALL, s), "vmult")
            kernel_events[1].record()
            self._chk(self._lib.glsb_vmult_finish(self._op, d, x, s), "vmult")

    def vmult_interface_down(self, dst: torch.Tensor, src: torch.Tensor):
        """operator_ns.cc:734-752 (GMG-LS edge matrix, down)."""
        w = self.time_integrator_data.get_primary_weight()
        if self.exchange is not None:
            self.exchange.vmult(self, dst, src, w)
            return
        self._chk(self._lib.glsb_vmult_interface_down(self._op, self._vec(dst, "dst"), self._vec(src, "src"), w,
                                                      self._stream()), "vmult_interface_down")

    def vmult_interface_up(self, dst: torch.Tensor, src: torch.Tensor):
        """operator_ns.cc:754-787 (GMG-LS edge matrix, up)."""
        w = self.time_integrator_data.get_primary_weight()
        s = self._stream()
        if self.exchange is None:
            self._chk(self._lib.glsb_vmult_interface_up(self._op, self._vec(dst, "dst"), self._vec(src, "src"), w, s),
                      "vmult_interface_up")
            return
        d = self._vec(dst, "dst")
        self._chk(self._lib.glsb_vmult_begin(self._op, d, s), "vmult_interface_up")
        if not getattr(self.mesh, "has_edge_constrained_indices", False):
            return
        cpy = torch.empty_like(src)
        self._chk(self._lib.glsb_edge_extract(self._op, self._vec(cpy, "cpy"), self._vec(src, "src"), s),
                  "vmult_interface_up")
        self._update_ghost_values(cpy)
        self._chk(self._lib.glsb_vmult_cells(self._op, d, self._vec(cpy, "cpy"), w, L.GLSB_CELLS_ALL, s),
                  "vmult_interface_up")
        self._compress_add(dst)

    def Tvmult(self, dst, src):
        """operator_base.cc:12-18: Tvmult forwards to vmult."""
        self.vmult(dst, src)

    def set_linearization_point(self, vec: torch.Tensor):
        """operator_ns.cc:570-620 (+ compute_penalty_parameters)."""
        self._update_ghost_values(vec)
        dt = self.time_integrator_data.get_current_dt()
        self._chk(self._lib.glsb_set_linearization_point(self._op, self._vec(vec, "vec"), dt, self._stream()),
                  "set_linearization_point")
        self._zero_ghosts(vec)

    def set_previous_solution(self, history):
        """operator_ns.cc:234-320; history = SolutionHistory or a list of vectors."""
        vecs = history.get_vectors() if hasattr(history, "get_vectors") else list(history)
        order = self.time_integrator_data.get_order()
        if order == 0:
            return
        for v in vecs[1:order + 1]:
            self._update_ghost_values(v)
        ptrs = (C.c_void_p * (order + 1))(*[self._vec(v, "history") for v in vecs[:order + 1]])
        w = self.time_integrator_data.get_weights()
        ws = (C.c_double * (order + 1))(*[float(x) for x in w[:order + 1]])
        self._chk(self._lib.glsb_set_previous_solution(self._op, ptrs, ws, order, self._stream()),
                  "set_previous_solution")
        for v in vecs[1:order + 1]:
            self._zero_ghosts(v)

    def evaluate_residual(self, dst: torch.Tensor, src: torch.Tensor):
        """operator_ns.cc:648-682."""
        tmp = src.clone()
        self.constraints_inhomogeneous.distribute(tmp)
        self._update_ghost_values(tmp)
        w = self.time_integrator_data.get_primary_weight()
        self._chk(self._lib.glsb_evaluate_residual(self._op, self._vec(dst, "dst"), self._vec(tmp, "src"), w,
                                                   self._stream()), "evaluate_residual")
        self._compress_add(dst)

    def evaluate_rhs(self, dst: torch.Tensor):
        """operator_ns.cc:622-646."""
        self.evaluate_residual(dst, torch.zeros_like(dst))

    def compute_inverse_diagonal(self, diagonal: torch.Tensor):
        """operator_ns.cc:195-225."""
        w = self.time_integrator_data.get_primary_weight()
        s = self._stream()
        if self.exchange is None:
            self._chk(self._lib.glsb_compute_inverse_diagonal(self._op, self._vec(diagonal, "diagonal"), w, s),
                      "compute_inverse_diagonal")
        else:
            self._chk(self._lib.glsb_diagonal_cells(self._op, self._vec(diagonal, "diagonal"), w, s),
                      "compute_inverse_diagonal")
            self._compress_add(diagonal)
            self._chk(self._lib.glsb_diagonal_finish(self._op, self._vec(diagonal, "diagonal"), s),
                      "compute_inverse_diagonal")

    def get_system_matrix(self) -> torch.Tensor:
        """operator_ns.cc:1303-1434 for the coarse-grid solver: the dense matrix of vmult ([n, n] float64 on the
        device), computed by the library from the operator's own cell loop."""
        n = self.n_local
        A = torch.empty((n, n), dtype=torch.float64, device=self.device)
        w = self.time_integrator_data.get_primary_weight()
        self._chk(self._lib.glsb_get_system_matrix(self._op, C.c_void_p(A.data_ptr()), w, self._stream()),
                  "get_system_matrix")
        return A

    def get_max_u(self, vec: torch.Tensor) -> float:
        """operator_ns.cc:530-568."""
        self._update_ghost_values(vec)
        out = C.c_double(0.0)
        self._chk(self._lib.glsb_get_max_u(self._op, self._vec(vec, "vec"), C.byref(out), self._stream()),
                  "get_max_u")
        val = out.value
        self._zero_ghosts(vec)
        if self.exchange is not None:
            val = self.exchange.allreduce_max(val)
        return val

    # ---- introspection ----
    def get_table(self, name):
        """Copy a q-point table back as [field, cell, q]."""
        d, nq, nc = self.mesh.dim, self.mesh.n_loc, self.mesh.n_cells
        nf = {"u_star_value": d, "u_star_gradient": d * d, "p_star_gradient": d, "u_time_derivative_old": d,
              "u_old_gradient": d * d, "p_old_gradient": d, "delta_1": 1, "delta_2": 1,
              "delta_1_q": 1, "delta_2_q": 1}[name]
        q = 1 if name in ("delta_1", "delta_2") else nq
        out = torch.empty(nf * nc * q, dtype=self.dtype, device=self.device)
        self._chk(self._lib.glsb_get_table(self._op, name.encode(), C.c_void_p(out.data_ptr()), out.numel(),
                                           self._stream()), "get_table")
        return out.view(nf, nc, q)

    def launch_count(self):
        return int(self._lib.glsb_launch_count(self._op))

    def vmult_variant(self):
        return self._lib.glsb_vmult_variant(self._op).decode()

    def set_variant(self, v):
        self._chk(self._lib.glsb_set_variant(self._op, int(v)), "set_variant")

    # ---- end-to-end entry with HOST vectors (what a host-vector deal.II caller pays) ----
    def vmult_host(self, dst_host: torch.Tensor, src_host: torch.Tensor):
        """vmult on host tensors (the reference's VectorType lives in host memory, config.h:9-10): forwards
        to glsb_vmult_host, which pipelines upload, cell kernels and download in chunks.  Pinned tensors
        make the copies asynchronous.  With a ghost exchange attached the vectors go to the device whole."""
        for t, name in ((dst_host, "dst_host"), (src_host, "src_host")):
            if t.is_cuda or t.dtype != self.dtype or not t.is_contiguous() or t.numel() != self.n_local:
                raise ValueError(f"{name}: expected a contiguous host {self.dtype} vector of length {self.n_local}")
        if self.exchange is None:
            w = self.time_integrator_data.get_primary_weight()
            self._chk(self._lib.glsb_vmult_host(self._op, C.c_void_p(dst_host.data_ptr()),
                                                C.c_void_p(src_host.data_ptr()), w, self._stream()), "vmult_host")
            return
        key = "vh"
        if key not in self._pinned:
            self._pinned[key] = (self.initialize_dof_vector(), self.initialize_dof_vector())
        d_src, d_dst = self._pinned[key]
        w = self.time_integrator_data.get_primary_weight()
        s = self._stream()
        edge = getattr(self.mesh, "has_edge_constrained_indices", False) or getattr(self.mesh, "outflow_faces", None)
        if dst_host.is_pinned() and src_host.is_pinned() and not edge:
            # chunked upload / interior cells / download pipeline; the cells at the partition surface and the
            # two halves of the ghost exchange run in between on the device vectors
            self._chk(self._lib.glsb_vmult_host_begin(self._op, self._vec(d_dst, "dst"), self._vec(d_src, "src"),
                                                      C.c_void_p(dst_host.data_ptr()), C.c_void_p(src_host.data_ptr()),
                                                      w, s), "vmult_host_begin")
            self.exchange.update_ghost_values(self, d_src)
            self._chk(self._lib.glsb_vmult_cells(self._op, self._vec(d_dst, "dst"), self._vec(d_src, "src"), w,
                                                 L.GLSB_CELLS_BOUNDARY, s), "vmult_host")
            self.exchange.compress_add(self, d_dst)
            self._chk(self._lib.glsb_vmult_host_finish(self._op, self._vec(d_dst, "dst"),
                                                       C.c_void_p(dst_host.data_ptr()), s), "vmult_host_finish")
            return
        d_src.copy_(src_host, non_blocking=True)
        self.vmult(d_dst, d_src)
        dst_host.copy_(d_dst, non_blocking=True)
