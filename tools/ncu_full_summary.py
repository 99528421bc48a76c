#!/usr/bin/env python
"""Summarise one `ncu --set full` capture: ncu -i X.ncu-rep --page raw --csv > X.csv; python tools/ncu_full_summary.py X.csv"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg"]


def main(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        print("kernel:", d.get("Kernel Name", ("?", ""))[0][:100])
        for k in KEYS:
            if k in d:
                print(f"  {k:85s} {d[k][0]:>16s} {d[k][1]}")
        for h in hdr:
            if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct"):
                v = float(d[h][0].replace(",", "") or 0)
                if v >= 5.0:
                    print(f"  {h:85s} {v:16.2f} %")


if __name__ == "__main__":
    main(sys.argv[1])
