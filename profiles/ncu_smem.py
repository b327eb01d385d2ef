#!/usr/bin/env python
"""Shared-memory wavefronts per instruction class (and per source line) of an .ncu-rep.
usage: python profiles/ncu_smem.py rep [n_cells]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur = None; agg = {}; lines = collections.defaultdict(lambda: [0, 0, 0])
for r in rows:
    if not r: continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if r[0] != "":
        cur = (r[0], r[1].strip()[:70]); continue
    ws = d["L1 Wavefronts Shared"]
    w = int(ws) if ws.isdigit() else 0
    if not w: continue
    s = d["Source"].split(); op = s[1] if s[0].startswith("@") else s[0]
    e = int(d["Instructions Executed"]); i = int(d["L1 Wavefronts Shared Ideal"]) if d["L1 Wavefronts Shared Ideal"].isdigit() else 0
    a = agg.setdefault(op, [0, 0, 0, 0]); a[0] += e; a[1] += w; a[2] += i; a[3] += 1
    l = lines[cur]; l[0] += e; l[1] += w; l[2] += i
tw = sum(a[1] for a in agg.values())
print(f"shared wavefronts total {tw}" + (f" = {tw / n_cells:.1f} per cell" if n_cells else ""))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:12s} n={a[3]:4d} executed={a[0]:10d} wavefronts={a[1]:11d} ideal={a[2]:11d} wf/exec={a[1] / a[0]:.2f}")
print("-- per source line --")
for k, l in sorted(lines.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{100 * l[1] / tw:5.1f}% wf/exec={l[1] / l[0]:.2f} ideal={l[2] / l[0]:.2f}  {k[0]:>4s} {k[1]}")
