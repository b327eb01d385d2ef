#!/bin/bash
# sweep of bench configurations: each argument is "ENV=.. ENV=.. -- bench args"
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  envs="${cfg%%--*}"; args="${cfg#*--}"
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu-baseline --time-step-refinements -1 $args > gpurun_out/sweep2_$i.json 2> gpurun_out/sweep2_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep2_$i.json")); r=d["roofline"]
    print("$cfg", "->", round(d["value"],2), "GDoF/s kernel_ms", round(r["kernel_ms"],3), "frac", round(r["frac"],3), d["config"]["kernel_variant"])
except Exception as e:
    print("$cfg", "FAILED", e)
PY
done
