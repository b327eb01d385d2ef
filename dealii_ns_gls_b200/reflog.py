"""The reference's console log as data: parse the per-time-step lines `gls-app` prints and compare two runs.

    cycle\\t<k> at time t = <t> with delta_t = <dt> and u_max = <u>          main.cc:921-923
        [N] step <i>; residual = <r>   /   [N] step <i> ; residual = <r>    solver_nl.cc:53, :79
        [L] solved in <n> iterations.                                       solver_l.cc:70
        [N] solved in <n> iterations.                                       solver_nl.cc:88
        [S] l2-norm of solution: <x>                                        main.cc:971

The device driver (driver.Driver with verbose = True) prints the same lines, so the north-star criterion
"identical GMRES / Newton iteration counts" becomes a diff of two text files the day a deal.II build of the
reference exists:

    python -m dealii_ns_gls_b200.reflog reference.log device.log
"""
from __future__ import annotations

import re
import sys

_NUM = r"([-+0-9.eE]+|nan|inf)"
_CYCLE = re.compile(r"cycle\s+(\d+) at time t = " + _NUM + r" with delta_t = " + _NUM + r" and u_max = " + _NUM)
_NSTEP = re.compile(r"\[N\] step (\d+)\s*; residual = " + _NUM)
_LSOLVE = re.compile(r"\[L\] solved in (\d+) iterations")
_NSOLVE = re.compile(r"\[N\] solved in (\d+) iterations")
_SNORM = re.compile(r"\[S\] l2-norm of solution: " + _NUM)


def parse(text: str) -> list:
    """list of dicts, one per time step: cycle, t, dt, u_max, newton_residuals, linear_iterations,
    newton_iterations, solution_l2"""
    steps, cur = [], None
    for line in text.splitlines():
        m = _CYCLE.search(line)
        if m:
            cur = dict(cycle=int(m.group(1)), t=float(m.group(2)), dt=float(m.group(3)), u_max=float(m.group(4)),
                       newton_residuals=[], linear_iterations=[], newton_iterations=None, solution_l2=None)
            steps.append(cur)
            continue
        if cur is None:
            continue
        m = _NSTEP.search(line)
        if m:
            cur["newton_residuals"].append(float(m.group(2)))
            continue
        m = _LSOLVE.search(line)
        if m:
            cur["linear_iterations"].append(int(m.group(1)))
            continue
        m = _NSOLVE.search(line)
        if m:
            cur["newton_iterations"] = int(m.group(1))
            continue
        m = _SNORM.search(line)
        if m:
            cur["solution_l2"] = float(m.group(1))
    return steps


def format_step(rec: dict) -> str:
    """the lines of one time step in the reference's format (what Driver prints with verbose = True)"""
    out = [f"\ncycle\t{rec['cycle']} at time t = {rec['t']:g} with delta_t = {rec['dt']:g} and u_max = {rec['u_max']:g}"]
    res, lin = rec["newton_residuals"], rec["linear_iterations"]
    out.append(f"    [N] step 0; residual = {res[0]:g}")
    for i, n in enumerate(lin):
        out.append(f"    [L] solved in {n} iterations.")
        out.append(f"    [N] step {i + 1} ; residual = {res[i + 1]:g}")
    out.append(f"    [N] solved in {rec['newton_iterations']} iterations.")
    if rec.get("solution_l2") is not None:
        out.append(f"    [S] l2-norm of solution: {rec['solution_l2']:g}")
    return "\n".join(out)


def compare(a: list, b: list, rtol=1e-6) -> list:
    """differences between two parsed runs: iteration counts must be identical, norms agree to rtol"""
    diffs = []
    if len(a) != len(b):
        diffs.append(f"number of time steps: {len(a)} vs {len(b)}")
    for x, y in zip(a, b):
        k = x["cycle"]
        if x["newton_iterations"] != y["newton_iterations"]:
            diffs.append(f"cycle {k}: Newton iterations {x['newton_iterations']} vs {y['newton_iterations']}")
        if x["linear_iterations"] != y["linear_iterations"]:
            diffs.append(f"cycle {k}: GMRES iterations {x['linear_iterations']} vs {y['linear_iterations']}")
        for name in ("dt", "solution_l2"):
            if x[name] is not None and y[name] is not None and abs(x[name] - y[name]) > rtol * max(abs(x[name]), 1e-300):
                diffs.append(f"cycle {k}: {name} {x[name]:.10g} vs {y[name]:.10g}")
    return diffs


def main(argv):
    if len(argv) != 3:
        print(__doc__)
        return 2
    with open(argv[1]) as f:
        a = parse(f.read())
    with open(argv[2]) as f:
        b = parse(f.read())
    diffs = compare(a, b)
    for d in diffs:
        print(d)
    print(f"{len(a)} / {len(b)} time steps, {len(diffs)} differences")
    return 1 if diffs else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
