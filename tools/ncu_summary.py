#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + per-source-line instruction / stall shares.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fmaheavy.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:72s} {r[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
agg, cur, h = {}, None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 2 and r[0] == "Line No":
        h = r
    elif len(r) > 2 and r[0] != "" and h:
        try:
            inst, samp = int(r[h.index("Instructions Executed")]), int(r[h.index("# Samples")])
        except ValueError:
            continue
        k = (cur, int(r[0]))
        a = agg.get(k, (0, 0, r[1]))
        agg[k] = (a[0] + inst, a[1] + samp, r[1])
ti, ts = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
print(f"total warp-instructions {ti}, stall samples {ts}")
for (f, ln), (i, s, code) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{f}:{ln:4d} inst {100 * i / ti:5.1f}% samp {100 * s / ts:5.1f}%  {code.strip()[:100]}")

# optional: instruction / sample share by line range of a file:  FILE:lo-hi,...  in env NCU_RANGES
import os
rng = os.environ.get("NCU_RANGES")
if rng:
    for spec in rng.split(","):
        f, lh = spec.split(":")
        lo, hi = map(int, lh.split("-"))
        i = sum(v[0] for (ff, ln), v in agg.items() if ff == f and lo <= ln <= hi)
        sm = sum(v[1] for (ff, ln), v in agg.items() if ff == f and lo <= ln <= hi)
        print(f"range {spec:32s} inst {100 * i / ti:5.1f}%  samp {100 * sm / ts:5.1f}%")
