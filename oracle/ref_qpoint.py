"""ctypes binding of oracle/_ref/libref_qpoint.so: the reference's OWN quadrature-point kernel
(NavierStokesOperator::do_vmult_cell, both branches, include/operator_ns.cc:880-1182), compiled unmodified by
oracle/build_ref_qpoint.sh on stand-in types.  TEST INFRASTRUCTURE ONLY.  `load()` returns None where the library
is absent and cannot be built (no reference tree); tests then fall back to the committed fixture
tests/golden/reference_qpoint.npz, written from this library by tests/golden/make_golden_reference_qpoint.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref_qpoint.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH) and os.path.exists("/root/reference/include/operator_ns.cc"):
            subprocess.call(["make", "-C", _HERE, "_ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if not os.path.exists(_PATH):
            return None
        lib = C.CDLL(_PATH)
        P, D, I = C.c_void_p, C.c_double, C.c_int
        lib.refq_apply.restype = I
        lib.refq_apply.argtypes = [I, I, I, I, I, D, D, D, I] + [P] * 12
        lib.refq_apply_f32.restype = I
        lib.refq_apply_f32.argtypes = [I, I, I, I, I, D, D, D, I] + [P] * 12
        lib.refq_boundary.restype = I
        lib.refq_boundary.argtypes = [I, I, I, D, D, I] + [P] * 8 + [I]
        lib.refq_face_beta.restype = I
        lib.refq_face_beta.argtypes = [I, I, I, P, P]
        lib.refq_penalty.restype = I
        lib.refq_penalty.argtypes = [I, D, D, D, D, I, I, I] + [P] * 7
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def apply(*, dim, residual, increment_form, ctd, cell_wise, theta, nu, weight, value, grad, u_star, u_star_grad,
          p_star_grad, u_tdo=None, u_old_grad=None, p_old_grad=None, d1, d2, number="double"):
    """One cell.  Point-major arrays: value[q, C], grad[q, C, dim] (what get_value / get_gradient return),
    u_star[q, dim], u_star_grad[q, dim, dim], p_star_grad[q, dim], u_tdo[q, dim] or None (no
    u_time_derivative_old table), u_old_grad[q, dim, dim] / p_old_grad[q, dim] or None, d1 / d2: [1] (cell-wise)
    or [q].  Returns (value_out[q, C], grad_out[q, C, dim]): the arguments of submit_value / submit_gradient."""
    lib = load()
    if lib is None:
        raise RuntimeError("oracle/_ref/libref_qpoint.so is not available")
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    value, grad, u_star, u_star_grad, p_star_grad = f(value), f(grad), f(u_star), f(u_star_grad), f(p_star_grad)
    u_tdo, u_old_grad, p_old_grad, d1, d2 = f(u_tdo), f(u_old_grad), f(p_old_grad), f(d1), f(d2)
    n_q = value.shape[0]
    vo, go_ = np.empty_like(value), np.empty_like(grad)
    fn = lib.refq_apply if number == "double" else lib.refq_apply_f32   # float: Number = MGNumber (config.h:7)
    rc = fn(dim, int(residual), int(increment_form), int(ctd), int(cell_wise), float(theta), float(nu),
            float(weight), n_q, _p(value), _p(grad), _p(u_star), _p(u_star_grad), _p(p_star_grad),
            _p(u_tdo), _p(u_old_grad), _p(p_old_grad), _p(d1), _p(d2), _p(vo), _p(go_))
    if rc != 0:
        raise RuntimeError("refq_apply failed")
    return vo, go_


def penalty(*, dim, dt, nu, c1, c2, degree, u, h_min, measure):
    """The body of compute_penalty_parameters' cell loop (include/operator_ns.cc:348-421): u[cell, q, dim] velocity
    values at the quadrature points, h_min[cell] = minimum_vertex_distance(), measure[cell].  Returns
    (delta_1[cell], delta_2[cell], delta_1_q[cell, q], delta_2_q[cell, q])."""
    lib = load()
    if lib is None:
        raise RuntimeError("oracle/_ref/libref_qpoint.so is not available")
    u = np.ascontiguousarray(u, dtype=np.float64)
    h_min = np.ascontiguousarray(h_min, dtype=np.float64)
    measure = np.ascontiguousarray(measure, dtype=np.float64)
    K, nq = u.shape[0], u.shape[1]
    d1c, d2c, d1q, d2q = np.empty(K), np.empty(K), np.empty((K, nq)), np.empty((K, nq))
    rc = lib.refq_penalty(dim, float(dt), float(nu), float(c1), float(c2), int(degree), K, nq, _p(u), _p(h_min),
                          _p(measure), _p(d1c), _p(d2c), _p(d1q), _p(d2q))
    if rc != 0:
        raise RuntimeError("refq_penalty failed")
    return d1c, d2c, d1q, d2q


def boundary(*, dim, residual, kind, nu, beta, value, grad, normal, face_velocity, target, dof_values=None):
    """do_vmult_boundary on one face (include/operator_ns.cc:1195-1301).  kind 1 = cut, 2 = Nitsche, 0 = a face
    whose boundary id carries no outflow terms.  value[q, C], grad[q, C, dim], normal[q, dim],
    face_velocity[q, dim], target[q, C].  Returns (value_out[q, C], grad_out[q, C, dim], dof_values after the
    call: zeroed by the reference for kind 0)."""
    lib = load()
    if lib is None:
        raise RuntimeError("oracle/_ref/libref_qpoint.so is not available")
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    value, grad, normal, face_velocity, target = f(value), f(grad), f(normal), f(face_velocity), f(target)
    dv = f(np.ones(4) if dof_values is None else dof_values).copy()
    vo, go_ = np.zeros_like(value), np.zeros_like(grad)
    rc = lib.refq_boundary(dim, int(residual), int(kind), float(nu), float(beta), value.shape[0], _p(value), _p(grad),
                           _p(normal), _p(face_velocity), _p(target), _p(vo), _p(go_), _p(dv), len(dv))
    if rc != 0:
        raise RuntimeError("refq_boundary failed")
    return vo, go_, dv


def face_beta(*, dim, degree, measure):
    """effective_beta_face (include/operator_ns.cc:428-457) of faces whose cells have the given measures"""
    lib = load()
    if lib is None:
        raise RuntimeError("oracle/_ref/libref_qpoint.so is not available")
    measure = np.ascontiguousarray(measure, dtype=np.float64)
    out = np.empty_like(measure)
    if lib.refq_face_beta(dim, int(degree), len(measure), _p(measure), _p(out)) != 0:
        raise RuntimeError("refq_face_beta failed")
    return out
