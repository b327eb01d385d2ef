#!/bin/bash
# quick GPU check: parity tests + a short bench (no CPU baseline); usage: bash profiles/quick.sh <tag> [bench args]
tag=$1; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$tag.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --time-step-refinements -1 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$tag.json"))
    print("value", round(d["value"],2), "GDoF/s  frac", round(d["roofline"]["frac"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "e2e", round(d["e2e"]["value"],2), d["config"]["kernel_variant"])
except Exception as e:
    print("no bench json", e); print(open("gpurun_out/bench_$tag.err").read()[-2000:])
PY
