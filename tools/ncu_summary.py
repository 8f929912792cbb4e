#!/usr/bin/env python3
"""Summarises an .ncu-rep (raw page) into the handful of counters DESIGN.md / profiles/ cite.
usage: python tools/ncu_summary.py report.ncu-rep [kernel-substring]"""
import csv, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for d in data:
    name = d[hdr.index("Kernel Name")]
    if pat and pat not in name: continue
    print("kernel:", name)
    for k in want:
        if k in hdr: print("  %-75s %-12s %s" % (k, units[hdr.index(k)], d[hdr.index(k)]))
    st = sorted(((float(d[hdr.index(k)] or 0), k) for k in stall), reverse=True)
    print("  stall reasons (warps stalled per issue-active cycle):")
    for v, k in st[:8]: print("    %-40s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
