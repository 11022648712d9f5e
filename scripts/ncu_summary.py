#!/usr/bin/env python
"""Tabulate the metrics that matter from an ncu report: python scripts/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__inst_executed_pipe_tensor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg"]


def main(path, extra=()):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", d.get("Kernel Name", "")[:80], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for h, u in zip(hdr, units):
            if any(h.startswith(k) for k in KEYS) or any(e in h for e in extra):
                print(f"   {h:75s} {d[h]:>18s} {u}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
