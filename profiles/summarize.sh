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
