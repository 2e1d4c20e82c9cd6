#!/usr/bin/env python3
"""Print the handful of ncu raw-page metrics this project tracks.
usage: ncu_summary.py <raw.csv>   (from `ncu -i x.ncu-rep --page raw --csv`)"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum",
    "sm__cycles_elapsed.avg",
]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("kernel:", name[:60])
    for k in KEYS:
        for i, h in enumerate(hdr):
            if h == k:
                print("  %-68s %s %s" % (k, r[i], units[i]))
    for i, h in enumerate(hdr):
        if "average_warps_issue_stalled" in h and h.endswith(
                "per_issue_active.ratio"):
            v = float(r[i] or 0)
            if v >= 0.15:
                print("  stall %-60s %.2f" % (h.replace(
                    "smsp__average_warps_issue_stalled_", "").replace(
                        "_per_issue_active.ratio", ""), v))
