#!/bin/bash
# one ncu --set full capture of the vmult kernel; usage: bash profiles/prof.sh <tag> [bench args]
tag=$1; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_vmult -s 4 -c 1 -f -o gpurun_out/prof_$tag \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --time-step-refinements -1 "$@" > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
