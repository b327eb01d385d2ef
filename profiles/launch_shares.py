#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, total ms, share.
usage: python profiles/launch_shares.py gpurun_out/launches_<tag>.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ui]]
    k = r[ki][:110]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total GPU time in the captured launches: {tot:.3f} ms (cold-cache, serialised: read the shares, not the absolutes)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
    print(f"{v[1]:10.3f} ms {v[0]:4d}x {100 * v[1] / tot:5.1f}%  avg {v[1] / v[0]:8.4f} ms  {k}")
