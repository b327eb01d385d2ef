"""Writes tests/golden/turek_2d1.json: the DFG benchmark 2D-1 (input/input_turek_2D_Re20_stat.json of the reference)
on the CPU oracle at three mesh levels -- drag, lift and pressure difference as SimulationCylinder::postprocess
evaluates them, the residual-based forces, the Newton residual histories.  Run from the repo root:
    python tests/golden/make_golden_turek.py          (about 35 s)
tests/test_turek_benchmark.py re-runs levels 1 and 2 against this file and checks all three levels against the
published values."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import gls_turek as gt  # noqa: E402

if __name__ == "__main__":
    out = {"literature": gt.LITERATURE, "levels": {}}
    for level in (1, 2, 3):
        out["levels"][str(level)] = gt.run(level)
        print(level, {k: v for k, v in out["levels"][str(level)].items() if k != "newton_residuals"})
    with open(os.path.join(ROOT, "tests", "golden", "turek_2d1.json"), "w") as f:
        json.dump(out, f, indent=1)
