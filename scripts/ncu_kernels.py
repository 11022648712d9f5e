"""Launch each hand-written kernel at config-2 shapes, three rounds in a fixed order (for `ncu --set full -k regex:<family>
-s <2 rounds> -c <1 round>`): round r launches every case of the family once, so skipping 2*len(cases) launches profiles
a warm third round."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlanimate_b200 import _lib as L, layers as Ly, ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
L.load(build_if_missing=False)
dev, bt = torch.device("cuda"), torch.bfloat16
b, f = 2, 16
levels = [(320, 64), (640, 32), (1280, 16)]
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for _ in range(3):
    for c, s in levels:
        x = torch.randn(b, c, f, s, s, device=dev, dtype=bt)
        xn = Ly.to_native(x)
        g, be, te = torch.ones(c, device=dev), torch.zeros(c, device=dev), torch.randn(b, c, device=dev)
        tok = xn.permute(0, 2, 3, 4, 1).reshape(-1, c)
        T = tok.shape[0]
        res = torch.randn_like(tok)
        flush.fill_(0.0)
        if which in ("all", "gn"):
            ops.groupnorm_silu(x, g, be, 32, 1e-5, temb=te)
            ops.groupnorm_silu(xn, g, be, 32, 1e-5, temb=te)
        if which in ("all", "ln"):
            ops.layernorm_pe(tok, g, be, 1e-5, pe=torch.randn(32, c, device=dev), frames=f, sites=s * s)
        if which in ("all", "attn"):
            qkv = torch.randn(T, 3 * c, device=dev, dtype=bt)
            ops.temporal_attention_core(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], batch=b, frames=f, sites=s * s, heads=8)
        if which in ("all", "xattn"):
            qx = torch.randn(T, c, device=dev, dtype=bt)
            kvx = torch.randn(b, 77, 2 * c, device=dev, dtype=bt)
            ops.cross_attention_core(qx, kvx[:, :, :c], kvx[:, :, c:], frames=b * f, sites=s * s, heads=8)
        if which in ("all", "fused") and c == 320:
            wq = torch.randn(c, c, device=dev, dtype=bt) * c ** -0.5
            perm = ops.pack_qkv_per_head(wq, wq, wq, 8)
            ops.temporal_attention_fused(tok, g, be, torch.randn(32, c, device=dev), perm, wq, be, batch=b, frames=f, sites=s * s, heads=8)
        if which in ("all", "gemm"):
            w3 = torch.randn(3 * c, c, device=dev, dtype=bt)
            w1 = torch.randn(c, c, device=dev, dtype=bt)
            wg = torch.randn(8 * c, c, device=dev, dtype=bt)
            w2 = torch.randn(c, 4 * c, device=dev, dtype=bt)
            bias = torch.randn(8 * c, device=dev)
            u = torch.randn(T, 4 * c, device=dev, dtype=bt)
            ops.linear(tok, w3)                                   # qkv
            ops.linear(tok, w1, bias[:c], residual=res)           # out + bias + residual
            ops.linear(tok, wg, bias, geglu=True)                 # GEGLU
            ops.linear(u, w2, bias[:c], residual=res)             # ff out + residual
        torch.cuda.synchronize()
torch.cuda.synchronize()
