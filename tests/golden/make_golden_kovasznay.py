"""Writes tests/golden/kovasznay.json: nodal errors of the oracle's stationary solutions against the analytic Kovasznay
flow (oracle/gls_exact.py) on sequences of meshes -- 2-D straight and curved cells (8^2 .. 32^2), 3-D rotated
(3^3, 6^3, 12^3; the last one takes about 7 minutes of sparse LU).  Run from the repo root:
    python tests/golden/make_golden_kovasznay.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import gls_exact as ge  # noqa: E402

if __name__ == "__main__":
    out = {"Re": ge.RE, "cases": {}}
    for name, dim, curved, ns in (("2d_straight", 2, False, (8, 16, 32)), ("2d_curved", 2, True, (8, 16, 32)),
                                  ("3d_rotated", 3, False, (3, 6, 12))):
        rows = []
        for n in ns:
            r = ge.solve(dim, n, curved=curved)
            rows.append({"n": n, "n_dofs": r["n_dofs"], "err_u": r["err_u"], "err_p": r["err_p"],
                         "newton_steps": len(r["newton_residuals"]) - 1})
            print(name, rows[-1], flush=True)
        out["cases"][name] = rows
    with open(os.path.join(ROOT, "tests", "golden", "kovasznay.json"), "w") as f:
        json.dump(out, f, indent=1)
