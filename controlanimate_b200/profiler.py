"""Launch counting and optional CUDA-event timing of every C-ABI kernel call (used by bench.py).

Counting is always on (an integer add).  With `enable(True)` each call is bracketed by CUDA events recorded on the
launching (current) stream; `summary()` turns them into per-kernel-family time, algorithmic GB/s or TFLOP/s and the
fraction of the measured roofline.  Timing is OFF inside bench.py's timed region.
"""
from __future__ import annotations

import contextlib
from collections import defaultdict

import torch

_enabled = False
_count = 0
_spans = []
_wall = None


def reset():
    global _count, _spans
    _count, _spans = 0, []


def launch_count() -> int:
    return _count


def enable(flag: bool, backlog_ms: float = 0.0):
    """backlog_ms > 0: first park the GPU on a spin kernel for that long so the host runs ahead and every timed span
    measures back-to-back DEVICE time (otherwise the interval between two in-stream events also contains the host's
    launch latency whenever the GPU out-runs Python)."""
    global _enabled, _spans, _wall
    _enabled = flag
    if flag:
        _spans = []
        if backlog_ms > 0:
            torch.cuda._sleep(int(backlog_ms * 1e-3 * 1.9e9))
        _wall = torch.cuda.Event(enable_timing=True)
        _wall.record()


@contextlib.contextmanager
def span(name: str, launches: int = 1, nbytes: float = 0.0, flops: float = 0.0):
    global _count
    _count += launches
    if not _enabled:
        yield
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    yield
    e1.record()
    _spans.append((name, e0, e1, nbytes, flops, launches))


def summary(peaks: dict) -> dict:
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    torch.cuda.synchronize()
    wall_ms = _wall.elapsed_time(end)
    fam = defaultdict(lambda: dict(launches=0, ms=0.0, bytes=0.0, flops=0.0))
    for name, e0, e1, nb, fl, n in _spans:
        f = fam[name]
        f["launches"] += n
        f["ms"] += e0.elapsed_time(e1)
        f["bytes"] += nb
        f["flops"] += fl
    out = {}
    for name, f in fam.items():
        sec = f["ms"] * 1e-3
        gbs = f["bytes"] / sec / 1e9 if sec > 0 else 0.0
        tfs = f["flops"] / sec / 1e12 if sec > 0 else 0.0
        tensor_bound = f["flops"] > 0 and (f["flops"] / max(f["bytes"], 1.0)) > peaks["tf_sustained"] * 1e12 / (peaks["hbm"] * 1e9)
        entry = dict(launches=f["launches"], ms_total=f["ms"], us_per_launch=f["ms"] * 1e3 / max(f["launches"], 1),
                     share_of_step=f["ms"] / wall_ms, bytes_per_launch=f["bytes"] / max(f["launches"], 1))
        if tensor_bound:
            entry.update(bound="tensor", achieved=tfs, peak=peaks["tf_sustained"], unit="TFLOP/s", frac=tfs / peaks["tf_sustained"],
                         flops_per_launch=f["flops"] / max(f["launches"], 1), hbm_gbs=gbs)
        else:
            entry.update(bound="hbm", achieved=gbs, peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"])
        out[name] = entry
    own_ms = sum(f["ms"] for f in fam.values())
    dominant_name = max(out, key=lambda k: out[k]["ms_total"]) if out else None
    dom = None
    if dominant_name:
        d = out[dominant_name]
        dom = dict(kernel=dominant_name, bound=d["bound"], achieved=d["achieved"], peak=d["peak"], unit=d["unit"], frac=d["frac"],
                   traffic=None, peak_source=f'{peaks["source"]} ({"sustained bf16 GEMM" if d["bound"] == "tensor" else "HBM copy"})',
                   avg_launch_us=d["us_per_launch"], launches=d["launches"], share_of_step=d["share_of_step"],
                   timing="CUDA events around each launch on the launching stream, instrumented pass after the timed region")
    return dict(families=out, dominant=dom, own_share=own_ms / wall_ms, wall_ms=wall_ms,
                launches_per_step=sum(f["launches"] for f in fam.values()))
