"""ctypes binding of oracle/_ref/libref_time_integration.so: the reference's OWN include/time_integration.cc,
compiled unmodified from /root/reference by `make -C oracle _ref` (see oracle/Makefile and
oracle/ref_time_integration_wrap.cc).  TEST INFRASTRUCTURE ONLY.  `load()` returns None where the library is
absent and cannot be built (no reference tree): the tests then use the committed fixture
tests/golden/reference_time_integration.json, which tests/golden/make_golden_reference_ti.py wrote from this
library."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref_time_integration.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH) and os.path.exists("/root/reference/include/time_integration.cc"):
            subprocess.call(["make", "-C", _HERE, "_ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if not os.path.exists(_PATH):
            return None
        lib = C.CDLL(_PATH)
        lib.reft_create.restype = C.c_void_p
        lib.reft_create.argtypes = [C.c_int, C.c_int, C.c_double]
        lib.reft_destroy.argtypes = [C.c_void_p]
        lib.reft_update_dt.restype = C.c_int
        lib.reft_update_dt.argtypes = [C.c_void_p, C.c_double]
        lib.reft_query.restype = C.c_int
        lib.reft_query.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                                   C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint)]
        lib.reft_history_commit.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        _lib = lib
    return _lib


class ReferenceTimeIntegrator:
    """TimeIntegratorDataBDF / Theta / None of the reference, as compiled object code"""
    BDF, THETA, NONE = 0, 1, 2

    def __init__(self, kind, order=0, theta=1.0):
        self.lib = load()
        if self.lib is None:
            raise RuntimeError("oracle/_ref/libref_time_integration.so is not available")
        self.h = self.lib.reft_create(kind, order, theta)

    def update_dt(self, dt):
        """True if the reference accepted the step size, False if it threw"""
        return self.lib.reft_update_dt(self.h, float(dt)) == 0

    def query(self):
        w = (C.c_double * 8)()
        pw, dt, th, order = C.c_double(), C.c_double(), C.c_double(), C.c_uint()
        n = self.lib.reft_query(self.h, w, 8, C.byref(pw), C.byref(dt), C.byref(th), C.byref(order))
        return {"weights": [w[i] for i in range(n)], "primary_weight": pw.value, "current_dt": dt.value,
                "theta": th.value, "order": int(order.value)}

    def __del__(self):
        if getattr(self, "h", None) and self.lib is not None:
            self.lib.reft_destroy(self.h)
            self.h = None


def history_after_commits(values, commits):
    lib = load()
    n = len(values)
    a = (C.c_double * n)(*values)
    out = (C.c_double * n)()
    lib.reft_history_commit(n, a, commits, out)
    return [out[i] for i in range(n)]
