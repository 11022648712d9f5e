"""Launch each hand-written kernel a few times at config-2 shapes (for `ncu --set full -k regex:...`)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlanimate_b200 import _lib as L, layers as Ly, ops

which = sys.argv[1] if len(sys.argv) > 1 else "all"
L.load(build_if_missing=False)
dev, bt = torch.device("cuda"), torch.bfloat16
b, c, f, s = 2, 320, 16, 64
x = torch.randn(b, c, f, s, s, device=dev, dtype=bt)
xn = Ly.to_native(x)
g, be, te = torch.ones(c, device=dev), torch.zeros(c, device=dev), torch.randn(b, c, device=dev)
tok = xn.permute(0, 2, 3, 4, 1).reshape(-1, c)
T = tok.shape[0]
for _ in range(3):
    if which in ("all", "gn"):
        ops.groupnorm_silu(x, g, be, 32, 1e-5, temb=te)
        ops.groupnorm_silu(xn, g, be, 32, 1e-5, temb=te)
    if which in ("all", "ln"):
        ops.layernorm_pe(tok, g, be, 1e-5, pe=torch.randn(32, c, device=dev), frames=f, sites=s * s)
    if which in ("all", "attn"):
        qkv = torch.randn(T, 3 * c, device=dev, dtype=bt)
        ops.temporal_attention_core(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], batch=b, frames=f, sites=s * s, heads=8)
    if which in ("all", "gemm"):
        w3 = torch.randn(3 * c, c, device=dev, dtype=bt)
        wg = torch.randn(8 * c, c, device=dev, dtype=bt)
        ops.linear(tok, w3)
        ops.linear(tok, wg, torch.randn(8 * c, device=dev), geglu=True)
torch.cuda.synchronize()
