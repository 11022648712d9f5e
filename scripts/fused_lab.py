"""Time the fused temporal-attention block against the four-launch path at config-2's 64x64 level (b2 f16 d4096 C320)."""
import json, os, statistics, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlanimate_b200 import _lib as L, ops
L.load(build_if_missing=False)
dev, bt = torch.device("cuda"), torch.bfloat16
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.fill_(1.0); flush[: flush.numel() // 2].sum(); torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)
rows = []
for (b, f, d, c) in ((2, 16, 4096, 320), (2, 8, 4096, 320), (1, 32, 4096, 320), (2, 16, 1024, 320)):
    T, heads = b * f * d, 8
    x = torch.randn(T, c, device=dev, dtype=bt)
    wq, wk, wv, wo = (torch.randn(c, c, device=dev, dtype=bt) * c ** -0.5 for _ in range(4))
    g, be, bo = torch.ones(c, device=dev), torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    pe = torch.randn(32, c, device=dev)
    perm = ops.pack_qkv_per_head(wq, wk, wv, heads)
    w3 = torch.cat([wq, wk, wv]).contiguous()
    y = torch.empty_like(x)
    def fused():
        ops.temporal_attention_fused(x, g, be, pe, perm, wo, bo, batch=b, frames=f, sites=d, heads=heads, out=y)
    def unfused():
        n = ops.layernorm_pe(x, g, be, 1e-5, pe=pe, frames=f, sites=d)
        qkv = ops.linear(n, w3)
        o = ops.temporal_attention_core(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], batch=b, frames=f, sites=d, heads=heads)
        return ops.linear(o, wo, bo, residual=x)
    tf, tu = timeit(fused), timeit(unfused)
    yu = unfused(); fused()
    err = float((y.float() - yu.float()).abs().max())
    flops = T * (8.0 * c * c + 4.0 * f * c)
    row = dict(shape=f"b{b} f{f} d{d} c{c}", fused_us=round(tf, 1), unfused_us=round(tu, 1), fused_tflops=round(flops / tf / 1e6, 0),
               hbm_roofline_us=round(2.0 * T * c * 2 / 6534e3, 1), tensor_roofline_us=round(flops / 1369.7e6, 1), max_abs_diff_vs_unfused=err)
    rows.append(row); print(json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/fused_lab.json", "w"), indent=1)
