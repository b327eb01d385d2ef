"""ctypes wrapper of oracle/gls_oracle_c.c (TEST INFRASTRUCTURE ONLY; see that file's header).

``COracle.from_numpy_oracle(o, branch)`` takes mesh, geometry and q-point tables from a
``gls_oracle.OracleOperator`` and runs the cell loop in C (SIMD over cell batches, OpenMP)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _isa_level():
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return "v3"
    need = ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl")
    return "v4" if all(f in flags for f in need) else "v3"


def load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, f"libgls_oracle_c_{_isa_level()}.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE])
        lib = C.CDLL(path)
        P, D, I, L = C.c_void_p, C.c_double, C.c_int, C.c_int64
        lib.glso_create.restype = P
        lib.glso_create.argtypes = [I, I, L, L, P, P, P, P, I, P, P, D, D, I, I, I]
        lib.glso_set_tables.restype = None
        lib.glso_set_tables.argtypes = [P] * 9
        lib.glso_destroy.restype = None
        lib.glso_destroy.argtypes = [P]
        lib.glso_apply.restype = None
        lib.glso_apply.argtypes = [P, P, P, D, I, I]
        lib.glso_diagonal.restype = None
        lib.glso_diagonal.argtypes = [P, P, D, I, I]
        lib.glso_max_threads.restype = I
        lib.glso_max_threads.argtypes = []
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def max_threads():
    """All host threads this process may use (torchrun sets OMP_NUM_THREADS=1; the affinity mask is
    what the box really offers)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return load().glso_max_threads()


class COracle:
    BR_NEWTON, BR_FIXED_POINT, BR_RESIDUAL = 0, 1, 2

    def __init__(self, *, dim, degree, cell_dofs, n_dofs, S, D, w, cartesian, inv_jac, jxw, nu, theta,
                 branch, ctd, cell_wise):
        self.lib = load()
        self.n_dofs = int(n_dofs)
        self._keep = [np.ascontiguousarray(cell_dofs, dtype=np.uint32), np.ascontiguousarray(S, dtype=np.float64),
                      np.ascontiguousarray(D, dtype=np.float64), np.ascontiguousarray(w, dtype=np.float64),
                      np.ascontiguousarray(inv_jac, dtype=np.float64), np.ascontiguousarray(jxw, dtype=np.float64)]
        k = self._keep
        self.h = self.lib.glso_create(dim, degree, k[0].shape[0], self.n_dofs, _p(k[0]), _p(k[1]), _p(k[2]),
                                      _p(k[3]), int(cartesian), _p(k[4]), _p(k[5]), float(nu), float(theta),
                                      int(branch), int(ctd), int(cell_wise))

    def set_tables(self, U, H=None, P=None, O=None, Gold=None, gold_p=None, d1=None, d2=None):
        """tables as [cell, field, q] arrays"""
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64)
                for a in (U, H, P, O, Gold, gold_p, d1, d2)]
        self.lib.glso_set_tables(self.h, *[_p(a) for a in arrs])

    def apply(self, src, weight, n_threads=0):
        src = np.ascontiguousarray(src, dtype=np.float64)
        dst = np.empty(self.n_dofs, dtype=np.float64)
        self.lib.glso_apply(self.h, _p(dst), _p(src), float(weight), 1, int(n_threads))
        return dst

    def apply_into(self, dst, src, weight, n_threads=0):
        self.lib.glso_apply(self.h, _p(dst), _p(src), float(weight), 1, int(n_threads))

    def diagonal(self, weight, n_threads=0):
        """sum over the cells of the diagonal of the cell matrix (unit-vector applications, plain scatter)"""
        diag = np.empty(self.n_dofs, dtype=np.float64)
        self.lib.glso_diagonal(self.h, _p(diag), float(weight), 1, int(n_threads))
        return diag

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.glso_destroy(self.h)
            self.h = None

    @classmethod
    def from_numpy_oracle(cls, o, branch):
        """Share mesh/geometry/tables with a gls_oracle.OracleOperator (double precision)."""
        K = o.n_cells
        d = o.dim
        co = cls(dim=d, degree=o.degree, cell_dofs=o.cell_dofs, n_dofs=o.n_dofs, S=o.tb.b.S, D=o.tb.b.D,
                 w=o.tb.b.wq, cartesian=False, inv_jac=o.Jinv, jxw=o.JxW, nu=o.nu, theta=o.theta,
                 branch=branch, ctd=o.ctd, cell_wise=o.cell_wise)
        d1 = o.delta1_cell.reshape(K, 1) if o.cell_wise else o.delta1_q
        d2 = o.delta2_cell.reshape(K, 1) if o.cell_wise else o.delta2_q
        co.set_tables(o.U, o.H.reshape(K, d * d, -1), o.P, o.o,
                      None if o.Gold is None else o.Gold.reshape(K, d * d, -1), o.gold_p, d1, d2)
        return co
