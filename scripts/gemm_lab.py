"""python scripts/gemm_lab.py [tag]: time every projection shape of config 2 (median of 10, L2 flushed); with
CA_GEMM_TIMING=1 the library also prints the per-role cycle attribution of each launch to stderr (development aid)."""
import json, os, statistics, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlanimate_b200 import _lib as L, ops
L.load(build_if_missing=False)
dev, bt = torch.device("cuda"), torch.bfloat16
tag = sys.argv[1] if len(sys.argv) > 1 else "lab"
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None   # e.g. "qkv m8192,geglu m131072"
timing = os.environ.get("CA_GEMM_TIMING") is not None
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
rows = []
for T, c in ((131072, 320), (32768, 640), (8192, 1280), (2048, 1280)):
    for nm, n, k, res, geglu in (("proj_in", c, c, False, False), ("qkv", 3 * c, c, False, False), ("out+res", c, c, True, False),
                                 ("geglu", 8 * c, c, False, True), ("ff_out+res", c, 4 * c, True, False)):
        if only and not any(o in f"{nm} m{T}" for o in only):
            continue
        x = torch.randn(T, k, device=dev, dtype=bt)
        w = torch.randn(n, k, device=dev, dtype=bt) * k ** -0.5
        bias = torch.randn(n, device=dev)
        n_out = n // 2 if geglu else n
        r = torch.randn(T, n_out, device=dev, dtype=bt) if res else None
        y = torch.empty(T, n_out, device=dev, dtype=bt)
        fn = lambda: ops.linear(x, w, bias, residual=r, geglu=geglu, out=y)
        try:
            fn(); torch.cuda.synchronize()
        except ValueError as e:
            print(json.dumps(dict(tag=tag, shape=f"{nm} m{T} n{n} k{k}", error=str(e)[:60])), flush=True)
            continue
        if timing:
            print(f"== {nm} m={T} n={n} k={k}", file=sys.stderr, flush=True)
            fn(); torch.cuda.synchronize()
            continue
        ts = []
        for _ in range(10):
            flush.fill_(1.0); flush[: flush.numel() // 2].sum(); torch.cuda._sleep(300000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = statistics.median(ts)
        tc = []
        if os.environ.get('LAB_CUBLAS'):
            F = torch.nn.functional
            def cub():
                if geglu:
                    a_, g_ = F.linear(x, w, bias.to(bt)).chunk(2, dim=-1); return a_ * F.gelu(g_)
                if res: return torch.addmm(r, x, w.t()).add_(bias.to(bt))
                return F.linear(x, w, bias.to(bt))
            cub(); torch.cuda.synchronize()
            for _ in range(10):
                flush.fill_(1.0); flush[: flush.numel() // 2].sum(); torch.cuda._sleep(300000)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); cub(); e1.record(); torch.cuda.synchronize()
                tc.append(e0.elapsed_time(e1) * 1e3)
        row = dict(tag=tag, shape=f"{nm} m{T} n{n} k{k}", us=round(us, 1), tflops=round(2.0 * T * n * k / us / 1e6, 0))
        if tc: row['cublas_fused_us'] = round(statistics.median(tc), 1)
        rows.append(row); print(json.dumps(row), flush=True)
        del x, w, r, y
if rows:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(f"gpurun_out/gemm_lab_{tag}.json", "w"), indent=1)
