#!/bin/bash
# One gpurun call: GPU parity tests, bench (plain), ncu launch list, one --set full capture of the
# top kernel.  Usage (from the repo root, on the GPU box): bash profiles/run_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cat gpurun_out/bench_$tag.json
# launch list of the same command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --time-step-refinements -1 > gpurun_out/ncu_launch_$tag.log 2>&1; echo "ncu list rc=$?"
# full capture of the vmult kernel (one launch) at the bench's own size, so that dram bytes per launch
# (profiles/traffic.json, read by bench.py) refer to the same launch as roofline.achieved
ncu --set full --clock-control none --import-source on -k regex:k_vmult -s 4 -c 1 -f -o gpurun_out/prof_$tag \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --time-step-refinements -1 > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
