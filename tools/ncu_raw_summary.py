#!/usr/bin/env python3
"""key metrics per kernel launch of an ncu report: ncu -i R.ncu-rep --page raw --csv > raw.csv; python tools/ncu_raw_summary.py raw.csv"""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem thr %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram thr %"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 thr %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 thr %"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("sm__inst_executed_pipe_fp64.sum", "fp64 inst"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps/cyc"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp cycles per inst"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 ld sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 ld requests"),
    ("lts__t_sectors_op_atom.sum", "L2 atom sectors"),
    ("lts__t_sectors_op_red.sum", "L2 red sectors"),
]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("==", name[:70])
    for key, label in WANT:
        if key in hdr:
            i = hdr.index(key)
            print(f"   {label:<22} {r[i]} {units[i]}")
