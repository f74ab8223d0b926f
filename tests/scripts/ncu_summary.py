"""One-kernel summary (the metrics the profiles/ files quote) from `ncu -i X.ncu-rep --page raw --csv`.
usage: ncu -i X.ncu-rep --page raw --csv | python ncu_summary.py"""
import csv
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
ix = {h: i for i, h in enumerate(hdr)}
for k in KEYS:
    if k in ix:
        print(f"{k},{units[ix[k]]},{vals[ix[k]][:100]}")
