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


def traffic_json(path, kernel, batch, source):
    """append {batch, dram_bytes, source} for `kernel` to profiles/ncu_traffic.json (bench.py reads roofline.traffic from it)"""
    import json
    import os
    rows = [r for r in csv.reader(open(path)) if r]
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total = 0.0
    for vals in rows[2:3]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = d[k]
            total += float(v.replace(",", "")) * scale.get(u, 1.0)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    db = json.load(open(out)) if os.path.exists(out) else {}
    ents = [e for e in db.get(kernel, []) if int(e["batch"]) != int(batch)]
    ents.append({"batch": int(batch), "dram_bytes": total, "source": source})
    db[kernel] = ents
    json.dump(db, open(out, "w"), indent=1)
    print("wrote", out, kernel, batch, total)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "--traffic-json":
        traffic_json(sys.argv[1], sys.argv[3], sys.argv[4], sys.argv[5])
    else:
        main(sys.argv[1])
