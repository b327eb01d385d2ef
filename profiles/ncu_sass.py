#!/usr/bin/env python
"""SASS listing of an .ncu-rep with samples per instruction; prints windows around the hottest instructions.
usage: python profiles/ncu_sass.py rep [min_samples] [context]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = int(sys.argv[2]) if len(sys.argv) > 2 else 500; ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
hot = [i for i, r in enumerate(data) if int(r[iS]) >= thr]
shown = set()
for h in hot:
    lo, hi = max(0, h - ctx), min(len(data), h + 3)
    if any(i in shown for i in range(lo, hi)):
        lo = max(lo, max(shown) + 1)
    else:
        print("-----")
    for i in range(lo, hi):
        r = data[i]; shown.add(i)
        s = {hdr[k][6:]: int(r[k]) for k in st if r[k].isdigit() and int(r[k]) > 0}
        top = " ".join(f"{k}:{v}" for k, v in sorted(s.items(), key=lambda kv: -kv[1])[:2])
        print(f"{i:5d} {int(r[iS]):6d} ex={int(r[iEx]):8d}  {r[iSrc].strip()[:80]:80s} {top}")
