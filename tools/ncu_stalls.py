#!/usr/bin/env python
"""Key raw metrics + warp stall reasons of every kernel in an ncu report.  usage: tools/ncu_stalls.py X.ncu-rep"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(txt.splitlines())); h = r[0]
for row in r[2:]:
    d = dict(zip(h, row))
    print(d.get("Kernel Name", "")[:70], "dur(us)", d.get("gpu__time_duration.sum"))
    for k in ("launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
              "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct"):
        if k in d: print("   ", k, d[k], r[1][h.index(k)])
    items = [(float(d[k]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
             for k in h if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and d[k] not in ("", "n/a")]
    print("    stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(items, reverse=True)[:8]))
