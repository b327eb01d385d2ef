"""Golden record of the solver-level behaviour of the CPU restatement (oracle/gls_solver.py):

  python tests/golden/make_golden_solver.py     ->  tests/golden/solver_channel.json

For each channel configuration of tests/test_gpu_solver.py: Newton and GMRES iteration counts, time-step
sizes, first Newton residuals and the l2 norm of the solution after each of three time steps.  The reference
has no recorded iteration counts, so this pins the ORACLE's solver stack (a later change
that alters a count fails tests/test_solver_oracle.py) and gives the device path a file-based target."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from dealii_ns_gls_b200.driver import ChannelParameters, CylinderParameters  # noqa: E402
from tests.test_solver_oracle import _oracle_driver  # noqa: E402

CASES = {
    "q1": dict(n_global_refinements=0),
    "q1_r1": dict(n_global_refinements=1),
    "q2": dict(n_global_refinements=0, fe_degree=2),
    "3d_q1": dict(dim=3, n_global_refinements=0, fe_degree=1),
    "bdf2": dict(n_global_refinements=0, bdf_order=2),
    "qwise": dict(n_global_refinements=0, cell_wise_stabilization=False, nu=0.01),
    "stationary_q2": dict(n_global_refinements=0, time_integration="none", fe_degree=2, cell_wise_stabilization=False,
                          nu=0.05),
}


# the BASELINE configs that are not the channel, on the synthetic O-grid (driver.CylinderParameters): curved cells,
# no-slip cylinder rows, q-point-wise delta; level operators in float like the reference (config.h:7)
CYLINDER_CASES = {
    # input_turek_2D_Re20_stat.json:13-37: stationary ("none"), exact Newton, c1 = 1
    "turek2d_stat": (dict(dim=2, time_integration="none", c_1=1.0, u_max=0.3, n_global_refinements=2), 1),
    "turek2d_bdf2": (dict(dim=2, n_global_refinements=2, newton_inexact=True, u_max=0.3), 3),
    # input_turek_3D_Re100.json:12-31: BDF2, inexact Newton, no-slip walls
    "turek3d_bdf2": (dict(dim=3, n_global_refinements=1, base_shape=(2, 8, 2), newton_inexact=True, u_max=1.0,
                          no_slip_wall=True), 2),
    # input_hoffmann_3D_Re3900.json:37 / main.cc:285-287: slip walls.  Its impulsive first step needs 18-36 GMRES
    # iterations per Newton step, where the float-level V-cycle's round-off (summation order) moves a count by one;
    # recorded with double level operators, for which device and oracle agree to round-off
    "hoffmann3d_slip": (dict(dim=3, n_global_refinements=1, base_shape=(2, 8, 2), newton_inexact=True, u_max=1.0,
                             no_slip_wall=False, mg_number="double"), 2),
}


def run(kw, n_steps=3, params=ChannelParameters, **okw):
    d = _oracle_driver(params(**kw), **okw)
    out = []
    for _ in range(n_steps):
        r = d.step()
        out.append(dict(newton_iterations=r["newton_iterations"], linear_iterations=r["linear_iterations"],
                        dt=r["dt"], first_residuals=[float(x) for x in r["newton_residuals"][:2]],
                        solution_l2=float(np.linalg.norm(d.history[0]))))
    return out


if __name__ == "__main__":
    if "cylinder" in sys.argv[1:]:
        only = [a for a in sys.argv[2:]]
        rec = {name: dict(parameters=dict(kw, base_shape=list(kw["base_shape"])) if "base_shape" in kw else kw,
                          steps=run(kw, n, params=CylinderParameters,
                                    level_dtype=np.float64 if kw.get("mg_number") == "double" else np.float32))
               for name, (kw, n) in CYLINDER_CASES.items() if not only or name in only}
        if only:  # regenerate selected cases, keep the others
            with open(os.path.join(HERE, "solver_cylinder.json")) as f:
                old = json.load(f)
            old.update(rec)
            rec = old
        with open(os.path.join(HERE, "solver_cylinder.json"), "w") as f:
            json.dump(rec, f, indent=1)
        for name, r in rec.items():
            print(name, [(s["newton_iterations"], s["linear_iterations"]) for s in r["steps"]])
        sys.exit(0)
    rec = {name: dict(parameters=kw, steps=run(kw)) for name, kw in CASES.items()}
    with open(os.path.join(HERE, "solver_channel.json"), "w") as f:
        json.dump(rec, f, indent=1)
    for name, r in rec.items():
        print(name, [(s["newton_iterations"], s["linear_iterations"]) for s in r["steps"]])
