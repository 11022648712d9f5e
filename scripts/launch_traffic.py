#!/usr/bin/env python
"""Per-family DRAM traffic and time from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` launch list of one denoising step (scripts/gpu_round.sh):

  python scripts/launch_traffic.py gpurun_out/launches.csv > profiles/rNN_traffic.json

bench.py reads the newest profiles/r*_traffic.json to fill `roofline.traffic` (bytes per launch of the dominant kernel).
"""
import collections
import csv
import json
import re
import sys

FAMILIES = {  # bench.py family -> substrings of the kernel names that implement it
    "linear_tcgen05": ["gemm_pair_kernel", "gemm_kernel", "gemm_tcgen05"],
    "groupnorm_silu": ["gn_ring_kernel", "gn_slab_kernel", "gn_split", "gn_bfhwc", "gn_ncfhw", "gn_finalize"],
    "layernorm_pe": ["layernorm_ring_kernel", "layernorm_pe_kernel", "layernorm_flat_kernel"],
    "row_stats": ["row_stats_kernel"],
    "temporal_attn_core": ["temporal_attn_kernel"],
    "cross_attn_core": ["cross_attn_kernel"],
    "bias_act_residual": ["bias_act_residual_kernel"],
    "residual_merge": ["residual_merge_kernel"],
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(path):
    with open(path) as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    per_launch = collections.defaultdict(dict)
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
        per_launch[(row["ID"], row["Kernel Name"])][row["Metric Name"]] = v
    fam = collections.defaultdict(lambda: dict(launches=0, us=0.0, dram_read=0.0, dram_write=0.0))
    for (_, name), m in per_launch.items():
        for f, keys in FAMILIES.items():
            if "ca::" in name and any(k in name for k in keys):
                d = fam[f]
                d["launches"] += 1
                d["us"] += m.get("gpu__time_duration.sum", 0.0)
                d["dram_read"] += m.get("dram__bytes_read.sum", 0.0)
                d["dram_write"] += m.get("dram__bytes_write.sum", 0.0)
                break
    out = {}
    for f, d in fam.items():
        n = max(d["launches"], 1)
        out[f] = dict(launches=d["launches"], us_per_launch=d["us"] / n, dram_read_per_launch=d["dram_read"] / n,
                      dram_write_per_launch=d["dram_write"] / n, traffic_per_launch=(d["dram_read"] + d["dram_write"]) / n)
    json.dump(dict(source=re.sub(r".*/", "", path), note="ncu launch list of one eager denoising step: cold-cache, serialised launches",
                   families=out), sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1])
