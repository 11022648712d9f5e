"""CA_GEMM_TIMING=1 python scripts/gemm_timing_set.py: per-role cycle attribution of the CTA-pair GEMM on the 64x64-level shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlanimate_b200 import _lib as L, ops
L.load(build_if_missing=False)
dev, bt = torch.device("cuda"), torch.bfloat16
cases = [("proj_in", 131072, 320, 320, False, False), ("qkv", 131072, 960, 320, False, False), ("out+res", 131072, 320, 320, True, False),
         ("geglu", 131072, 2560, 320, False, True), ("ff_out+res", 131072, 320, 1280, True, False),
         ("qkv32", 32768, 1920, 640, False, False), ("ff_out32", 32768, 640, 2560, True, False), ("out32", 32768, 640, 640, True, False)]
for name, m, n, k, res, geglu in cases:
    x = torch.randn(m, k, device=dev, dtype=bt)
    w = torch.randn(n, k, device=dev, dtype=bt) * k ** -0.5
    bias = torch.randn(n, device=dev)
    r = torch.randn(m, n, device=dev, dtype=bt) if res else None
    fn = lambda: ops.linear(x, w, bias, residual=r, geglu=geglu)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sys.stderr.flush()
    print(f"== {name} m={m} n={n} k={k}", flush=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"   {e0.elapsed_time(e1)*1e3:.1f} us, {2.0*m*n*k/e0.elapsed_time(e1)/1e9:.0f} TFLOP/s", flush=True)
