#!/bin/bash
# env-knob + argument sweep of bench.py: usage bash profiles/sweep3.sh <tag> "<VAR=val ...> -- <bench args>" ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  envs="${cfg%%--*}"; args=""
  case "$cfg" in *--*) args="${cfg#*--}";; esac
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --time-step-refinements -1 $args > gpurun_out/sweep_${tag}_$i.json 2> gpurun_out/sweep_${tag}_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_${tag}_$i.json"))
    print("[$cfg]", "->", round(d["value"],2), "GDoF/s kernel_ms", round(d["roofline"]["kernel_ms"],3), "frac", round(d["roofline"]["frac"],3), d["config"].get("kernel_variant"), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("[$cfg]", "FAILED", e)
PY
done
