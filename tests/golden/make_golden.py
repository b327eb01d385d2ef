"""Generates the golden fixtures under tests/golden/ from the numpy oracle.

  python tests/golden/make_golden.py

The reference ships no golden vectors and cannot be run here, so these
fixtures pin the ORACLE (tests/test_golden.py fails if a later change alters its results) and give
the CUDA path and the C++ host mirror a fixed, file-based target.  Each case stores the complete
operator input (mesh description exactly as the C ABI takes it, vectors, time-integrator scalars)
and the oracle's outputs."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from dealii_ns_gls_b200 import mesh as gm  # noqa: E402
from oracle import gls_oracle as go  # noqa: E402

CASES = {
    # name: (mesh factory, flags)
    "channel_2d_q1": (lambda: _channel(), dict(nu=0.0, ctd=True, cell_wise=True, order=1, weights=[40.0, -40.0],
                                               dt=0.025, theta=1.0, increment_form=True)),
    "turek_2d_q2_stat": (lambda: gm.cylinder_shell((3, 8), 2), dict(nu=0.001, ctd=True, cell_wise=False, order=0,
                                                                    weights=[], dt=1.0, theta=1.0,
                                                                    increment_form=True)),
    "turek_3d_q2_bdf2": (lambda: gm.cylinder_shell((2, 6, 2), 2), dict(nu=0.001, ctd=True, cell_wise=False, order=2,
                                                                       weights=[15.0, -20.0, 5.0], dt=0.1, theta=1.0,
                                                                       increment_form=True)),
    "cube_3d_q2_perf": (lambda: gm.hypercube(3, 3, 2), dict(nu=0.1, ctd=False, cell_wise=True, order=2,
                                                            weights=[10.0, -10.0, 0.0], dt=0.1, theta=1.0,
                                                            increment_form=True)),
    "cube_3d_q3_hanging": (lambda: gm.add_random_constraints(gm.hypercube(3, 2, 3), 10, 6, seed=4),
                           dict(nu=0.05, ctd=False, cell_wise=True, order=2, weights=[10.0, -10.0, 0.0], dt=0.1,
                                theta=1.0, increment_form=True)),
    "shell_2d_q2_theta": (lambda: gm.cylinder_shell((2, 6), 2), dict(nu=0.01, ctd=False, cell_wise=False, order=1,
                                                                     weights=[10.0, -10.0], dt=0.1, theta=0.5,
                                                                     increment_form=False)),
}


def _channel():
    """input_channel.json-like: 2-D Q1 channel, inflow/no-slip velocity rows and outflow pressure row zero."""
    def dirichlet(ref, c):
        if c < 2:
            return (np.abs(ref[:, 0]) < 1e-12) | (np.abs(ref[:, 1]) < 1e-12) | (np.abs(ref[:, 1] - 1) < 1e-12)
        return np.abs(ref[:, 0] - 4.0) < 1e-12
    return gm.structured_mesh(2, (8, 4), 1, extent=(4.0, 1.0), dirichlet=dirichlet)


def build_case(name):
    factory, f = CASES[name]
    mesh = factory()
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    o = go.OracleOperator(dim=mesh.dim, degree=mesh.degree, cell_dofs=mesh.cell_dofs, n_dofs=mesh.n_dofs,
                          cell_points=mesh.cell_points, mapping_degree=mesh.mapping_degree,
                          constraints=mesh.constraints, nu=f["nu"], c1=4.0, c2=2.0, theta=f["theta"],
                          order=f["order"], consider_time_derivative=f["ctd"], increment_form=f["increment_form"],
                          cell_wise_stabilization=f["cell_wise"], path="naive")
    hist = [rng.uniform(-1, 1, mesh.n_dofs) for _ in range(f["order"] + 1)]
    lin = rng.uniform(-1, 1, mesh.n_dofs)
    src = rng.uniform(-1, 1, mesh.n_dofs)
    if f["order"] > 0:
        o.set_previous_solution(hist, f["weights"])
    o.set_linearization_point(lin, f["dt"])
    w0 = f["weights"][0] if f["weights"] else 0.0
    sb = src.copy()
    sb[list(mesh.constraints.keys())] = 0.0
    cdofs = np.array(sorted(mesh.constraints.keys()), dtype=np.int64)
    row_ptr, ecol, ev = [0], [], []
    for d in cdofs:
        for m, w in mesh.constraints[int(d)]:
            ecol.append(m), ev.append(w)
        row_ptr.append(len(ecol))
    if mesh.geometry_type == 0:
        ij, jxw = mesh.cart_inv_jac, mesh.cart_det
    else:
        ij, jxw = gm.general_geometry(mesh)
    return dict(
        dim=mesh.dim, degree=mesh.degree, n_dofs=mesh.n_dofs, cell_dofs=mesh.cell_dofs.astype(np.uint32),
        geometry_type=mesh.geometry_type, inv_jac=np.asarray(ij), jxw=np.asarray(jxw), cell_points=mesh.cell_points,
        mapping_degree=mesh.mapping_degree, cell_h_min=mesh.cell_h_min, cell_measure=mesh.cell_measure,
        row_dof=cdofs.astype(np.uint32), row_ptr=np.array(row_ptr, dtype=np.uint32),
        entry_col=np.array(ecol, dtype=np.uint32), entry_val=np.array(ev, dtype=np.float64),
        nu=f["nu"], c1=4.0, c2=2.0, theta=f["theta"], order=f["order"], weights=np.array(f["weights"], dtype=np.float64),
        dt=f["dt"], ctd=int(f["ctd"]), cell_wise=int(f["cell_wise"]), increment_form=int(f["increment_form"]),
        history=np.array(hist), lin=lin, src=src, src_bc=sb,
        out_vmult=o.vmult(src, w0), out_residual=o.evaluate_residual(sb, w0),
        out_inv_diag=o.compute_inverse_diagonal(w0), out_max_u=o.get_max_u(src),
        delta_1=o.delta1_cell, delta_2=o.delta2_cell, delta_1_q=o.delta1_q, delta_2_q=o.delta2_q)


def write_cpp_dump(case, path):
    """Flat little-endian dump for the C++ host-mirror test (tests/cpp/test_operator_b200.cpp)."""
    with open(path, "wb") as f:
        def put(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            np.array([a.size], dtype=np.uint64).tofile(f)
            a.tofile(f)
        put([case["dim"], case["degree"], case["geometry_type"], case["order"], case["ctd"], case["cell_wise"],
             case["increment_form"], case["cell_dofs"].shape[0], case["n_dofs"]], np.int64)
        put([case["nu"], case["c1"], case["c2"], case["theta"], case["dt"]], np.float64)
        put(case["weights"], np.float64)
        for k, dt in (("cell_dofs", np.uint32), ("row_dof", np.uint32), ("row_ptr", np.uint32), ("entry_col", np.uint32),
                      ("entry_val", np.float64), ("inv_jac", np.float64), ("jxw", np.float64),
                      ("cell_h_min", np.float64), ("cell_measure", np.float64), ("history", np.float64),
                      ("lin", np.float64), ("src", np.float64), ("src_bc", np.float64), ("out_vmult", np.float64),
                      ("out_residual", np.float64), ("out_inv_diag", np.float64)):
            put(case[k], dt)
        put([case["out_max_u"]], np.float64)


if __name__ == "__main__":
    for name in CASES:
        c = build_case(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **c)
        print(name, "cells", c["cell_dofs"].shape[0], "dofs", c["n_dofs"])
    write_cpp_dump(build_case("turek_3d_q2_bdf2"), os.path.join(HERE, "turek_3d_q2_bdf2.bin"))
