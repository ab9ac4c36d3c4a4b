#!/usr/bin/env python3
"""Print the handful of raw ncu metrics we track per kernel capture.  usage: tools/ncu_summary.py <report.ncu-rep>"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for w in want:
        if w in d:
            print(f"{w} = {d[w]} {units[hdr.index(w)]}")
    st = sorted(((float(d[h] or 0), h) for h in stall), reverse=True)[:8]
    print("top stalls (warps per issue):", ", ".join(f"{h.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, h in st))
    print()
