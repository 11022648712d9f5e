"""python scripts/ln_fold_lab.py: LayerNorm(+PE) -> projection as two launches (ca_layernorm_pe + ca_linear) against the folded
form (ca_row_stats + ca_linear_ln) at the config-2 shapes; each pair timed as ONE region (median of 10, L2 flushed before)."""
import json, os, statistics, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlanimate_b200 import _lib as L, ops
L.load(build_if_missing=False)
dev, bt = torch.device("cuda"), torch.bfloat16
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def timeit(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.fill_(1.0); flush[: flush.numel() // 2].sum(); torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return round(statistics.median(ts), 1)


rows = []
f = 16
for T, c in ((131072, 320), (32768, 640), (8192, 1280), (2048, 1280)):
    sites = T // (2 * f)
    x = torch.randn(T, c, device=dev, dtype=bt)
    gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.1
    pe = torch.randn(32, c, device=dev)
    for nm, n, geglu, use_pe in (("qkv", 3 * c, False, True), ("q", c, False, False), ("geglu", 8 * c, True, False)):
        w = torch.randn(n, c, device=dev, dtype=bt) * c ** -0.5
        bias = torch.randn(n, device=dev) if geglu else None
        y = torch.empty(T, n // 2 if geglu else n, device=dev, dtype=bt)
        ln_out = torch.empty_like(x)
        st = torch.empty(T, 2, device=dev)
        fold = ops.fold_layernorm(w, gamma, beta, bias=bias, pe=pe if use_pe else None)

        def two():
            ops.layernorm_pe(x, gamma, beta, 1e-5, pe=pe if use_pe else None, frames=f, sites=sites, out=ln_out)
            return ops.linear(ln_out, w, bias, geglu=geglu, out=y)

        def folded():
            ops.row_stats(x, 1e-5, out=st)
            return ops.linear_ln(x, st, *fold, frames=f, sites=sites, geglu=geglu, out=y)

        row = dict(shape=f"{nm} m{T} n{n} k{c}", ln_plus_linear_us=timeit(two), stats_plus_linear_ln_us=timeit(folded),
                   layernorm_us=timeit(lambda: ops.layernorm_pe(x, gamma, beta, 1e-5, pe=pe if use_pe else None, frames=f, sites=sites, out=ln_out)),
                   linear_us=timeit(lambda: ops.linear(ln_out, w, bias, geglu=geglu, out=y)),
                   row_stats_us=timeit(lambda: ops.row_stats(x, 1e-5, out=st)),
                   linear_ln_us=timeit(lambda: ops.linear_ln(x, st, *fold, frames=f, sites=sites, geglu=geglu, out=y)))
        rows.append(row); print(json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/ln_fold_lab.json", "w"), indent=1)
