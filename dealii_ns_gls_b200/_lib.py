"""ctypes binding of include/glsb200.h.  Fails loudly when libglsb200.so is missing:
there is no CPU or PyTorch fallback for the operator."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GLSB_LIB: an alternative build of the same library (kernel experiments under profiles/), default the in-tree one
LIB_PATH = os.environ.get("GLSB_LIB") or os.path.join(_HERE, "libglsb200.so")

GLSB_ABI_VERSION = 3
GLSB_F64, GLSB_F32 = 0, 1
GLSB_GEOM_CARTESIAN, GLSB_GEOM_GENERAL = 0, 2
GLSB_CELLS_ALL, GLSB_CELLS_INTERIOR, GLSB_CELLS_BOUNDARY = 0, 1, 2
GLSB_CONSTRAINED_BIT = 0x80000000


class GlsbDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("dim", C.c_int32), ("degree", C.c_int32), ("number_type", C.c_int32),
        ("increment_form", C.c_int32), ("consider_time_derivative", C.c_int32),
        ("cell_wise_stabilization", C.c_int32), ("time_order", C.c_int32),
        ("nu", C.c_double), ("c1", C.c_double), ("c2", C.c_double), ("theta", C.c_double),
        ("n_cells", C.c_uint64), ("n_owned", C.c_uint64), ("n_ghost", C.c_uint64),
        ("dof_indices", C.c_void_p),
        ("n_constraint_rows", C.c_uint32),
        ("row_dof", C.c_void_p), ("row_ptr", C.c_void_p), ("entry_col", C.c_void_p), ("entry_val", C.c_void_p),
        ("n_constrained_indices", C.c_uint32), ("constrained_indices", C.c_void_p),
        ("geometry_type", C.c_int32), ("inv_jac", C.c_void_p), ("jxw", C.c_void_p),
        ("cell_h_min", C.c_void_p), ("cell_measure", C.c_void_p),
        ("n_export", C.c_uint64), ("export_indices", C.c_void_p),
        ("n_edge_constrained_indices", C.c_uint32), ("edge_constrained_indices", C.c_void_p),
        ("has_edge_constrained_indices", C.c_int32),
        ("n_outflow_faces", C.c_uint32),
        ("face_cell", C.c_void_p), ("face_no", C.c_void_p), ("face_kind", C.c_void_p),
        ("face_normal", C.c_void_p), ("face_jxw", C.c_void_p), ("face_inv_jac", C.c_void_p),
        ("face_target_velocity", C.c_void_p),
    ]


class GlsbTransferDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("dim", C.c_int32), ("degree", C.c_int32),
        ("number_type", C.c_int32),
        ("n_coarse_cells", C.c_uint64), ("n_fine_dofs", C.c_uint64), ("n_coarse_dofs", C.c_uint64),
        ("coarse_dof_indices", C.c_void_p), ("fine_dof_indices", C.c_void_p),
        ("n_constraint_rows", C.c_uint32),
        ("row_ptr", C.c_void_p), ("entry_col", C.c_void_p), ("entry_val", C.c_void_p),
        ("weights", C.c_void_p),
    ]


# every symbol include/glsb200.h declares: name -> (restype, argtypes)
_P, _D, _I, _U64 = C.c_void_p, C.c_double, C.c_int, C.c_uint64
SYMBOLS = {
    "glsb_create": (_I, [C.POINTER(GlsbDesc), C.POINTER(_P)]),
    "glsb_destroy": (None, [_P]),
    "glsb_last_error": (C.c_char_p, [_P]),
    "glsb_invalidate_system": (_I, [_P]),
    "glsb_vmult": (_I, [_P, _P, _P, _D, _P]),
    "glsb_vmult_host": (_I, [_P, _P, _P, _D, _P]),
    "glsb_vmult_host_begin": (_I, [_P, _P, _P, _P, _P, _D, _P]),
    "glsb_vmult_host_finish": (_I, [_P, _P, _P, _P]),
    "glsb_host_register": (_I, [_P, _U64]),
    "glsb_host_unregister": (_I, [_P]),
    "glsb_edge_begin": (_I, [_P, _P, _P]),
    "glsb_edge_finish": (_I, [_P, _P, _P, _P]),
    "glsb_vmult_interface_down": (_I, [_P, _P, _P, _D, _P]),
    "glsb_vmult_interface_up": (_I, [_P, _P, _P, _D, _P]),
    "glsb_edge_extract": (_I, [_P, _P, _P, _P]),
    "glsb_vmult_begin": (_I, [_P, _P, _P]),
    "glsb_vmult_cells": (_I, [_P, _P, _P, _D, _I, _P]),
    "glsb_vmult_finish": (_I, [_P, _P, _P, _P]),
    "glsb_vmult_cells_part": (_I, [_P, _P, _P, _D, _I, _I, _I, _P]),
    "glsb_set_sm_reserve": (_I, [_P, _I]),
    "glsb_evaluate_residual": (_I, [_P, _P, _P, _D, _P]),
    "glsb_evaluate_residual_cells": (_I, [_P, _P, _P, _D, _I, _P]),
    "glsb_set_linearization_point": (_I, [_P, _P, _D, _P]),
    "glsb_set_previous_solution": (_I, [_P, C.POINTER(_P), C.POINTER(_D), _I, _P]),
    "glsb_compute_inverse_diagonal": (_I, [_P, _P, _D, _P]),
    "glsb_diagonal_cells": (_I, [_P, _P, _D, _P]),
    "glsb_diagonal_finish": (_I, [_P, _P, _P]),
    "glsb_get_system_matrix": (_I, [_P, _P, _D, _P]),
    "glsb_get_max_u": (_I, [_P, _P, C.POINTER(_D), _P]),
    "glsb_relaxation_vmult": (_I, [_P, _P, _P, _P, _D, _I, _D, _P]),
    "glsb_relaxation_step": (_I, [_P, _P, _P, _P, _D, _I, _D, _P]),
    "glsb_relaxation_update": (_I, [_P, _P, _P, _P, _P, _D, _P]),
    "glsb_estimate_relaxation": (_I, [_P, _P, _I, _D, _D, _U64, C.POINTER(_D), C.POINTER(_D), _P]),
    "glsb_transfer_create": (_I, [C.POINTER(GlsbTransferDesc), C.POINTER(_P)]),
    "glsb_transfer_destroy": (None, [_P]),
    "glsb_transfer_last_error": (C.c_char_p, [_P]),
    "glsb_transfer_prolongate_and_add": (_I, [_P, _P, _P, _P]),
    "glsb_transfer_restrict_and_add": (_I, [_P, _P, _P, _P]),
    "glsb_transfer_interpolate": (_I, [_P, _P, _P, _P]),
    "glsb_vec_multi_dot": (_I, [_P, _P, _U64, _I, _P, _U64, _I, _P]),
    "glsb_vec_multi_axpy": (_I, [_P, _P, _U64, _I, _P, _D, _U64, _I, _P]),
    "glsb_vec_axpby": (_I, [_P, _D, _P, _D, _U64, _I, _P]),
    "glsb_vec_convert": (_I, [_P, _I, _P, _I, _U64, _P]),
    "glsb_vec_set_zero_indexed": (_I, [_P, _P, _U64, _I, _P]),
    "glsb_dense_apply": (_I, [_P, _P, _P, C.c_uint32, C.c_uint32, _I, _P]),
    "glsb_pack_export": (_I, [_P, _P, _P, _P]),
    "glsb_unpack_add": (_I, [_P, _P, _P, _P]),
    "glsb_n_cells": (_U64, [_P]),
    "glsb_n_local": (_U64, [_P]),
    "glsb_n_interior_cells": (_U64, [_P]),
    "glsb_get_table": (_I, [_P, C.c_char_p, _P, _U64, _P]),
    "glsb_launch_count": (_U64, [_P]),
    "glsb_vmult_variant": (C.c_char_p, [_P]),
    "glsb_set_variant": (_I, [_P, _I]),
}

_lib = None


def load():
    """Load libglsb200.so and bind every declared symbol; raise if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C dealii_ns_gls_b200/csrc`. The operator has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class GlsbError(RuntimeError):
    pass


def check(lib, op, rc, what):
    if rc != 0:
        msg = lib.glsb_last_error(op)
        raise GlsbError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
