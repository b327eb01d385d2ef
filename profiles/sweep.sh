#!/bin/bash
# env-knob sweep of the Q2 kernel: usage bash profiles/sweep.sh "<VAR=val ...>" "<VAR=val ...>" ...
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --steps 10 --warmup 3 --no-cpu-baseline --time-step-refinements -1 > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_$i.json"))
    print("$cfg", "->", round(d["value"],2), "GDoF/s kernel_ms", round(d["roofline"]["kernel_ms"],3), "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("$cfg", "FAILED", e)
PY
done
