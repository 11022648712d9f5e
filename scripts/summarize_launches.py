#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time, share and launch count per kernel.

  python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_step_launches.md
"""
import collections
import csv
import re
import sys


def family(name: str) -> str:
    if "ca::" in name:
        m = re.search(r"ca::(?:<unnamed>::)?(\w+)(<[^>(]*>)?", name)
        return "OWN  " + (m.group(1) + (m.group(2) or "") if m else name[:60])
    if "sdpa" in name or "flash" in name:
        return "LIB  attention (cuDNN/torch SDPA): " + re.sub(r"<.*", "", name)[:60]
    if "cutlass" in name or "conv" in name or "implicit_gemm" in name or "nhwcAddPadding" in name:
        return "LIB  cuDNN conv: " + re.sub(r"<.*", "", name)[:70]
    if "nvjet" in name or "gemm" in name.lower():
        return "LIB  cuBLAS (time-embedding MLPs): " + name[:40]
    m = re.search(r"at::(?:native::)?(?:<unnamed>::)?(\w+)", name)
    sub = re.search(r"(CUDAFunctor_add|direct_copy_kernel|silu_kernel|MulFunctor|bfloat16_copy|sin_kernel|cos_kernel|exp_kernel)", name)
    return "TORCH " + (m.group(1) if m else name[:40]) + (f" [{sub.group(1)}]" if sub else "")


def main(path):
    with open(path) as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        k = family(row["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += v
        total += v
    n = sum(a[0] for a in agg.values())
    print(f"# ncu launch list summary: {n} launches, {total / 1e3:.2f} ms serialised device time\n")
    cls = collections.defaultdict(float)
    for k, (_, t) in agg.items():
        cls[k.split()[0]] += t
    print("| class | ms | share |\n|---|---|---|")
    for k, t in sorted(cls.items(), key=lambda kv: -kv[1]):
        print(f"| {k} | {t / 1e3:.2f} | {100 * t / total:.1f}% |")
    print("\n| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / total < 0.0005:
            continue
        print(f"| {k} | {c} | {t / 1e3:.3f} | {100 * t / total:.1f}% | {t / c:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
