"""Writes tests/golden/turek_2d2.json (+ turek/turek_2d2_state.npz): the DFG benchmark 2D-2 (periodic flow around a
cylinder at Re = 100; input/input_turek_2D_Re100.json of the reference) on the CPU oracle -- BDF2, time-derivative
terms, q-point-wise stabilisation -- run from rest until the vortex shedding is periodic.  Records the force
history of the last periods, the maxima of drag and lift and the Strouhal number, and the three history vectors at
the end so that tests/test_turek_benchmark.py can continue the run for a few steps.  Takes 10 - 20 minutes:
    python tests/golden/make_golden_turek_unsteady.py [level] [t_final]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import gls_turek as gt  # noqa: E402

if __name__ == "__main__":
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    t_final = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
    sim = gt.UnsteadyTurek(level=level)
    t0 = time.perf_counter()
    while sim.t < t_final - 1e-12:
        r = sim.step()
        if sim.n_steps % 50 == 0:
            print(f"t = {r['t']:.4f}  c_D = {r['drag']:.5f}  c_L = {r['lift']: .5f}  Newton {r['newton_iterations']}  "
                  f"wall {time.perf_counter() - t0:.0f} s", flush=True)
    stats = gt.shedding_statistics(sim.records, t_final - 1.5)
    print(stats)
    tail = [r for r in sim.records if r["t"] >= t_final - 1.5]
    out = {"literature": gt.LITERATURE_2D2, "level": level, "dt": sim.dt, "t_final": sim.t, "n_steps": sim.n_steps,
           "n_cells": int(sim.mesh.n_cells), "n_dofs": int(sim.mesh.n_dofs), "statistics": stats, "tail": tail}
    os.makedirs(os.path.join(ROOT, "tests", "golden", "turek"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "turek", "turek_2d2_state.npz"),
                        history=np.stack(sim.history).astype(np.float64), t=sim.t, dt=sim.dt,
                        bdf_dt=np.array(sim.bdf.dt, dtype=np.float64))
    # three more steps from the saved state: what the test re-computes
    out["continuation"] = [sim.step() for _ in range(3)]
    with open(os.path.join(ROOT, "tests", "golden", "turek_2d2.json"), "w") as f:
        json.dump(out, f, indent=1)
