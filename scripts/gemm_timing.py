"""CA_GEMM_TIMING=1 python scripts/gemm_timing.py [n] [k] [m]: per-role cycle attribution of the CTA-pair GEMM."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlanimate_b200 import _lib as L, ops
L.load(build_if_missing=False)
dev, bt = torch.device("cuda"), torch.bfloat16
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
k = int(sys.argv[2]) if len(sys.argv) > 2 else 320
m = int(sys.argv[3]) if len(sys.argv) > 3 else 131072
x = torch.randn(m, k, device=dev, dtype=bt)
w = torch.randn(n, k, device=dev, dtype=bt)
for _ in range(2):
    ops.linear(x, w)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.linear(x, w); e1.record(); torch.cuda.synchronize()
print(f"n={n} k={k} m={m}: {e0.elapsed_time(e1)*1e3:.1f} us (sync'd launch), {2.0*m*n*k/e0.elapsed_time(e1)/1e9:.0f} TFLOP/s")
