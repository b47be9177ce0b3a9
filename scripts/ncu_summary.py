#!/usr/bin/env python
"""Key figures of every kernel in an .ncu-rep (needs `ncu` on PATH): ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[2:]:
    print("----", r[idx["Kernel Name"]][:60])
    for w in want:
        if w in idx:
            print(f"  {w:62s} {r[idx[w]]:>16s} {rows[1][idx[w]]}")
    for i, h in enumerate(hdr):
        # pipe utilisation (which pipe binds: fma / alu / xu / fp64 / lsu ...)
        if (h.startswith("sm__inst_executed_pipe_") or h.startswith("sm__pipe_")) and "pct_of_peak_sustained_active" in h:
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if v > 5:
                print(f"  pipe {h:57s} {v:16.2f} %")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if v > 0.15:
                print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:.2f}")
