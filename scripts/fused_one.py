import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlanimate_b200 import _lib as L, ops
L.load(build_if_missing=False)
dev, bt = torch.device("cuda"), torch.bfloat16
b, f, d, c, heads = 2, 16, 4096, 320, 8
T = b * f * d
x = torch.randn(T, c, device=dev, dtype=bt)
wq, wk, wv, wo = (torch.randn(c, c, device=dev, dtype=bt) * c ** -0.5 for _ in range(4))
g, be, bo = torch.ones(c, device=dev), torch.zeros(c, device=dev), torch.zeros(c, device=dev)
pe = torch.randn(32, c, device=dev)
perm = ops.pack_qkv_per_head(wq, wk, wv, heads)
y = torch.empty_like(x)
for _ in range(3):
    ops.temporal_attention_fused(x, g, be, pe, perm, wo, bo, batch=b, frames=f, sites=d, heads=heads, out=y)
torch.cuda.synchronize()
