"""Writes tests/golden/reference/qpoint.npz from the REFERENCE'S OWN object code (oracle/_ref/libref_qpoint.so =
lines 880-1182, 1195-1301 and 348-421 of /root/reference/include/operator_ns.cc compiled unmodified on stand-in types,
oracle/build_ref_qpoint.sh): for seeded random inputs, what do_vmult_cell hands to submit_value /
submit_gradient in every branch / flag combination the operator has, the same for do_vmult_boundary (cut and
Nitsche outflow faces), and the four stabilisation-parameter tables of compute_penalty_parameters.  Needs the reference tree; run from the repo root:
    python tests/golden/make_golden_reference_qpoint.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_qpoint as rq  # noqa: E402

# (dim, residual, increment_form, ctd, cell_wise, theta, has_u_time_derivative_old)
CASES = [(dim, res, inc, ctd, cw, th, old)
         for dim in (2, 3)
         for (res, inc, ctd, cw, th, old) in [
             (0, 1, 0, 1, 1.0, 1),   # Newton branch, performance.cc flags
             (0, 1, 1, 0, 1.0, 1),   # Newton branch, Turek flags (time derivative, q-point-wise delta)
             (0, 1, 1, 1, 1.0, 1),
             (0, 0, 1, 1, 1.0, 1),   # fixed-point branch (Picard / linearized)
             (0, 0, 1, 0, 0.5, 1),
             (0, 0, 0, 1, 1.0, 0),
             (1, 1, 1, 0, 1.0, 1),   # residual branch as Newton's right-hand side
             (1, 1, 1, 1, 1.0, 0),   # ... stationary: no u_time_derivative_old table
             (1, 0, 1, 1, 0.5, 1),   # residual branch of the theta scheme (old gradients)
             (1, 0, 0, 0, 0.75, 1),
         ]]
# (dim, kind: 1 cut / 2 Nitsche, residual)
BOUNDARY = [(dim, kind, res) for dim in (2, 3) for kind in (1, 2) for res in (0, 1)]
PENALTY = [(2, 2, 0.01, 0.001), (2, 1, 0.0, 0.5), (3, 2, 0.02, 0.001), (3, 3, 1.0, 0.05), (3, 2, 0.0, 0.2)]


def inputs(case_no, dim, n_q):
    rng = np.random.default_rng(1000 + case_no)
    r = lambda *s: rng.uniform(-1.0, 1.0, s)  # noqa: E731
    return dict(value=r(n_q, dim + 1), grad=r(n_q, dim + 1, dim), u_star=r(n_q, dim), u_star_grad=r(n_q, dim, dim),
                p_star_grad=r(n_q, dim), u_tdo=r(n_q, dim), u_old_grad=r(n_q, dim, dim), p_old_grad=r(n_q, dim),
                d1=rng.uniform(0.05, 1.0, n_q), d2=rng.uniform(0.05, 1.0, n_q))


def run_case(case_no, case, number="double"):
    dim, res, inc, ctd, cw, th, old = case
    n_q = 3 ** dim
    a = inputs(case_no, dim, n_q)
    return a, rq.apply(dim=dim, residual=res, increment_form=inc, ctd=ctd, cell_wise=cw, theta=th, nu=0.037,
                       weight=7.25, value=a["value"], grad=a["grad"], u_star=a["u_star"],
                       u_star_grad=a["u_star_grad"], p_star_grad=a["p_star_grad"], u_tdo=a["u_tdo"] if old else None,
                       u_old_grad=a["u_old_grad"] if th != 1.0 else None,
                       p_old_grad=a["p_old_grad"] if th != 1.0 else None,
                       d1=a["d1"][:1] if cw else a["d1"], d2=a["d2"][:1] if cw else a["d2"], number=number)


def boundary_inputs(no, dim):
    rng = np.random.default_rng(3000 + no)
    n_q = 3 ** (dim - 1)
    n = rng.uniform(-1.0, 1.0, (n_q, dim))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return dict(value=rng.uniform(-1.0, 1.0, (n_q, dim + 1)), grad=rng.uniform(-1.0, 1.0, (n_q, dim + 1, dim)), normal=n,
                face_velocity=rng.uniform(-1.0, 1.0, (n_q, dim)), target=rng.uniform(-1.0, 1.0, (n_q, dim + 1)))


def run_boundary(no, case):
    dim, kind, res = case
    a = boundary_inputs(no, dim)
    v, g, _ = rq.boundary(dim=dim, residual=res, kind=kind, nu=0.037, beta=3.7, **a)
    return a, (v, g)


def penalty_inputs(no, dim, degree):
    rng = np.random.default_rng(2000 + no)
    K, n_q = 12, (degree + 1) ** dim
    return dict(u=rng.uniform(-2.0, 2.0, (K, n_q, dim)), h_min=rng.uniform(0.01, 0.3, K),
                measure=rng.uniform(1e-4, 1e-2, K))


if __name__ == "__main__":
    assert rq.load() is not None, "build oracle/_ref first (needs /root/reference)"
    out = {"cases": np.array(CASES, dtype=np.float64), "penalty_cases": np.array(PENALTY, dtype=np.float64)}
    for i, case in enumerate(CASES):
        _, (v, g) = run_case(i, case)
        out[f"value_out_{i}"], out[f"grad_out_{i}"] = v, g
        _, (v, g) = run_case(i, case, number="float")     # Number = float (the multigrid level operators)
        out[f"value_out_f32_{i}"], out[f"grad_out_f32_{i}"] = v, g
    out["boundary_cases"] = np.array(BOUNDARY, dtype=np.float64)
    for i, case in enumerate(BOUNDARY):
        _, (v, g) = run_boundary(i, case)
        out[f"boundary_value_out_{i}"], out[f"boundary_grad_out_{i}"] = v, g
    for i, (dim, degree, dt, nu) in enumerate(PENALTY):
        a = penalty_inputs(i, dim, degree)
        r = rq.penalty(dim=dim, dt=dt, nu=nu, c1=4.0, c2=2.0, degree=degree, **a)
        for name, arr in zip(("d1_cell", "d2_cell", "d1_q", "d2_q"), r):
            out[f"{name}_{i}"] = arr
    # effective_beta_face (:428-457) for dim 2 / 3, degrees 1-4
    beta_measure = np.random.default_rng(4000).uniform(1e-5, 1e-1, 16)
    out["beta_measure"] = beta_measure
    for dim in (2, 3):
        for degree in (1, 2, 3, 4):
            out[f"beta_{dim}_{degree}"] = rq.face_beta(dim=dim, degree=degree, measure=beta_measure)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference", "qpoint.npz"), **out)
    print(len(CASES), "q-point cases,", len(BOUNDARY), "boundary cases,", len(PENALTY), "penalty cases written")
