"""Writes tests/golden/reference_time_integration.json from the REFERENCE'S OWN object code
(oracle/_ref/libref_time_integration.so = /root/reference/include/time_integration.cc compiled unmodified,
`make -C oracle _ref`): BDF weights of orders 1-3 over sequences of varying step sizes, the theta and the
stationary scheme, SolutionHistory::commit_solution.  Needs the reference tree; run from the repo root:
    python tests/golden/make_golden_reference_ti.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_time_integration as rt  # noqa: E402


def sequences():
    rng = np.random.default_rng(2024)
    out = [[0.1] * 5, [0.1, 0.1, 0.05, 0.2, 0.2], [1e-3, 2e-3, 4e-3, 8e-3]]
    for _ in range(5):
        out.append([float(x) for x in 0.01 * rng.uniform(0.5, 2.0, 6)])
    return out


if __name__ == "__main__":
    assert rt.load() is not None, "build oracle/_ref first (needs /root/reference)"
    rec = {"source": "include/time_integration.cc of the reference, compiled unmodified (oracle/Makefile: _ref)",
           "bdf": [], "theta": [], "none": None, "history": []}
    for order in (1, 2, 3):
        for seq in sequences():
            t = rt.ReferenceTimeIntegrator(rt.ReferenceTimeIntegrator.BDF, order=order)
            steps = []
            for dt in seq:
                ok = t.update_dt(dt)
                steps.append(dict(t.query(), accepted=ok))
            rec["bdf"].append({"order": order, "dts": seq, "after_each_update": steps})
    for theta in (1.0, 0.5, 0.75):
        t = rt.ReferenceTimeIntegrator(rt.ReferenceTimeIntegrator.THETA, theta=theta)
        steps = []
        for dt in (0.1, 0.025):
            t.update_dt(dt)
            steps.append(t.query())
        rec["theta"].append({"theta": theta, "dts": [0.1, 0.025], "after_each_update": steps})
    t = rt.ReferenceTimeIntegrator(rt.ReferenceTimeIntegrator.NONE)
    t.update_dt(0.3)
    rec["none"] = t.query()
    for size, commits in ((2, 1), (3, 1), (3, 2), (4, 3)):
        vals = [float(10 + i) for i in range(size)]
        rec["history"].append({"values": vals, "commits": commits, "after": rt.history_after_commits(vals, commits)})
    with open(os.path.join(ROOT, "tests", "golden", "reference_time_integration.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print("bdf cases:", len(rec["bdf"]), " first:", rec["bdf"][8]["after_each_update"][2])
