"""python scripts/fmha_lab.py: the spatial self-attention core (ca_spatial_attn_core) against torch SDPA (cuDNN) at the config-2
shapes; median of 10, L2 flushed."""
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
for frames, sites, heads, hd in ((32, 4096, 8, 40), (16, 4096, 8, 40), (32, 1024, 8, 40)):
    c = heads * hd
    qkv = torch.randn(frames * sites, 3 * c, device=dev, dtype=bt)
    out = torch.empty(frames * sites, c, device=dev, dtype=bt)
    q5 = qkv.reshape(frames, sites, 3, heads, hd)
    q, k, v = (q5[:, :, i].transpose(1, 2) for i in range(3))
    own = lambda: ops.spatial_attention_core(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], frames=frames, sites=sites, heads=heads, out=out)
    lib = lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v)
    o1 = own().float().reshape(frames, sites, heads, hd)
    o2 = lib().transpose(1, 2).float()
    err = float((o1 - o2).abs().max() / o2.abs().max())
    flops = 4.0 * frames * heads * sites * sites * hd
    t_own, t_lib = timeit(own), timeit(lib)
    row = dict(shape=f"frames{frames} sites{sites} heads{heads} hd{hd}", own_us=t_own, cudnn_us=t_lib, own_tflops=round(flops / t_own / 1e6),
               cudnn_tflops=round(flops / t_lib / 1e6), max_rel_diff=err)
    rows.append(row); print(json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/fmha_lab.json", "w"), indent=1)
