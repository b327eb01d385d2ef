#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU): key metrics + SASS-level stall/opcode mix.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [n_cells]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",  # ALL LSU wavefronts: shared + global
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__lsuin_requests.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_active.avg",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]
print("== metrics ==")
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:80s} {vals[i]:>16s} {units[i]}")
d = dict(zip(hdr, vals))
if n_cells:
    rd = float(d["dram__bytes_read.sum"]); wr = float(d["dram__bytes_write.sum"])
    ur = units[hdr.index("dram__bytes_read.sum")]; uw = units[hdr.index("dram__bytes_write.sum")]
    sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    tot = rd * sc[ur] + wr * sc[uw]
    print(f"dram traffic per launch {tot:.4e} B = {tot / n_cells:.1f} B/cell (read {rd*sc[ur]/n_cells:.1f}, write {wr*sc[uw]/n_cells:.1f})")
print("== stall samples ==")
st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(float(v)) for h, v in zip(hdr, vals)
      if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
tot = sum(st.values())
for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {k:24s} {v:8d} {100 * v / tot:5.1f}%")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr2, data = rows[1], rows[2:]
iS, iSrc, iEx = hdr2.index("# Samples"), hdr2.index("Source"), hdr2.index("Instructions Executed")
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    if len(r) <= iEx:
        continue
    t = r[iSrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += int(r[iEx]); samp[op] += int(r[iS])
te, ts = sum(ops.values()), sum(samp.values())
print(f"== SASS: {len(data)} instructions, {te} warp-instr executed" + (f", {te * 8 / n_cells:.0f} per cell-thread" if n_cells else "") + " ==")
for k, v in ops.most_common(18):
    extra = f" per-cell-thread {v * 8 / n_cells:7.1f}" if n_cells else ""
    print(f"  {k:10s} {100 * v / te:5.1f}%{extra}   samples {100 * samp[k] / ts:5.1f}%")
print("== top sampled instructions ==")
for i in sorted(range(len(data)), key=lambda i: -int(data[i][iS]) if len(data[i]) > iS else 0)[:14]:
    print(f"  {int(data[i][iS]):7d}  {data[i][iSrc][:100]}")
