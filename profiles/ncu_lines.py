#!/usr/bin/env python
"""Per-source-line stall samples of an .ncu-rep captured with --import-source on (kernels built
with -lineinfo).  usage: python profiles/ncu_lines.py rep.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    d = dict(zip(hdr, r))
    try:
        s = int(d["# Samples"])
    except ValueError:
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
    lines.append((s, cur_file, r[0], r[1].strip()[:90], int(d["Instructions Executed"]), stalls))
tot = sum(l[0] for l in lines)
print(f"total samples {tot}")
for s, f, ln, src, ex, st in sorted(lines, key=lambda x: -x[0])[:top]:
    top3 = " ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * s / tot:5.1f}% {s:6d} ex={ex:9d} {f}:{ln:>4s}  {src}\n        [{top3}]")
