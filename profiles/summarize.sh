#!/bin/bash
# usage: bash profiles/summarize.sh <tag> <cells in the ncu-full capture>   (reads gpurun_out/, writes profiles/<tag>_*.txt)
tag=$1; cells=$2
{
  echo "# $tag: bench line (plain run, no profiler)"; cat gpurun_out/bench_$tag.json
  echo; echo "# $tag: ncu launch list of bench.py --steps 3 --warmup 3 (shares)"; python profiles/launch_shares.py gpurun_out/launches_$tag.csv
  echo; echo "# $tag: ncu --set full of one vmult launch ($cells cells)"; python profiles/ncu_summary.py gpurun_out/prof_$tag.ncu-rep $cells
  echo; echo "# $tag: stall samples per source line"; python profiles/ncu_lines.py gpurun_out/prof_$tag.ncu-rep 30
} > profiles/${tag}_summary.txt 2>&1
cp gpurun_out/launches_$tag.csv profiles/${tag}_launches.csv
# dram bytes of the captured launch -> profiles/traffic.json (bench.py reports it as roofline.traffic)
python - <<PY
import csv, io, json, subprocess
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_$tag.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h, u, v = rows[0], rows[1], rows[2]
sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
tot = sum(float(v[h.index(k)]) * sc[u[h.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
json.dump({"kernel": v[h.index("Kernel Name")], "cells": $cells, "dram_bytes_per_launch": tot,
           "source": "profiles/${tag}_summary.txt (ncu --set full, one launch of bench.py's workload)"},
          open("profiles/traffic.json", "w"), indent=1)
print("traffic", tot, "B per launch =", tot / $cells, "B/cell")
PY
