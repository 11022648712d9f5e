#!/usr/bin/env python
"""BASELINE config 5: microbenchmark sweep of the hand-written kernels against the measured rooflines.

frames 8-32 x channels 320/640/1280 x latents 32^2-96^2 (b=2), both memory layouts for GroupNorm.  Each timing is the
median of `--iters` launches, every launch preceded by an L2 flush (256 MB write, then a 256 MB read so that the timed
kernel starts on a cold L2 WITHOUT inheriting ~126 MB of dirty lines whose write-back would be billed to it) and
bracketed by CUDA events on the launching stream.  Writes one JSON document (default gpurun_out/microbench.json).
"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlanimate_b200 import _lib as L, layers as Ly, ops  # noqa: E402


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        flush[: flush.numel() // 2].sum()
        # park the GPU (~150 us) so that the launch below is already queued when e0 fires: the event interval then holds
        # device time only, not the host's launch latency
        torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="gpurun_out/microbench.json")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="", help="comma list of kernel families: gn,ln,attn,xattn,fused,copy,gemm,merge")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    want = lambda fam: not only or fam in only  # noqa: E731
    L.load(build_if_missing=False)
    dev = torch.device("cuda")
    hbm, tf, src = peaks()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    rows = []
    bt = torch.bfloat16
    b = 2
    frames = [16] if args.quick else [8, 16, 24, 32]
    levels = [(320, 64), (640, 32), (1280, 16), (1280, 8)] if args.quick else \
        [(320, 32), (320, 48), (320, 64), (320, 96), (640, 32), (640, 48), (1280, 16), (1280, 24), (1280, 8)]

    def add(kernel, shape, us, us_min, nbytes, flops=0.0, **extra):
        gbs = nbytes / us / 1e3
        tfs = flops / us / 1e6
        tensor_bound = flops > 0 and extra.pop("tensor", False)
        r = dict(kernel=kernel, shape=shape, us=round(us, 2), us_min=round(us_min, 2), alg_bytes=nbytes, gbs=round(gbs, 1),
                 frac_hbm=round(gbs / hbm, 4))
        if flops:
            r.update(tflops=round(tfs, 1), frac_tensor=round(tfs / tf, 4))
        r["bound"] = "tensor" if tensor_bound else "hbm"
        r.update(extra)
        rows.append(r)
        print(json.dumps(r), flush=True)

    for f in frames:
        for c, s in levels:
            n = b * c * f * s * s
            if n * 2 > 1.5e9:
                continue
            x = torch.randn(b, c, f, s, s, device=dev, dtype=bt)
            xn = Ly.to_native(x)
            g, be = torch.ones(c, device=dev), torch.zeros(c, device=dev)
            te = torch.randn(b, c, device=dev)
            for name, inp in (("ncfhw", x), ("bfhwc", xn)) if want("gn") else ():
                y = torch.empty_like(inp)
                us, mn = timeit(lambda: ops.groupnorm_silu(inp, g, be, 32, 1e-5, temb=te, out=y), args.iters, flush)
                add("groupnorm_silu", f"b{b} c{c} f{f} {s}x{s} {name} +temb", us, mn, 2.0 * n * 2)
            if want("copy"):
                # yardsticks on the same tensor and the same timing harness: a plain device copy and the elementwise conv epilogue
                y = torch.empty_like(xn)
                us, mn = timeit(lambda: y.copy_(xn), args.iters, flush)
                add("torch_copy(yardstick)", f"b{b} c{c} f{f} {s}x{s}", us, mn, 2.0 * n * 2)
                x4 = xn.permute(0, 2, 1, 3, 4).reshape(b * f, c, s, s)
                bz = te[0].contiguous()
                us, mn = timeit(lambda: ops.bias_act_residual(x4, bz, silu=True, inplace=True), args.iters, flush)
                add("bias_act_residual(silu)", f"n{b * f} c{c} {s}x{s}", us, mn, 2.0 * n * 2)
            # LayerNorm + PE on tokens
            tok = xn.permute(0, 2, 3, 4, 1).reshape(-1, c)
            pe = torch.randn(32, c, device=dev)
            yt = torch.empty_like(tok)
            if want("ln"):
                us, mn = timeit(lambda: ops.layernorm_pe(tok, g, be, 1e-5, pe=pe, frames=f, sites=s * s, out=yt), args.iters, flush)
                add("layernorm_pe", f"T{tok.shape[0]} c{c} f{f}", us, mn, 2.0 * n * 2)
            # temporal attention core on a packed QKV buffer
            T = tok.shape[0]
            qkv = torch.randn(T, 3 * c, device=dev, dtype=bt)
            o = torch.empty(T, c, device=dev, dtype=bt)
            if want("attn"):
                us, mn = timeit(lambda: ops.temporal_attention_core(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], batch=b, frames=f,
                                                                    sites=s * s, heads=8, out=o), args.iters, flush)
                add("temporal_attn_core", f"b{b} f{f} d{s * s} c{c} (hd {c // 8})", us, mn, 4.0 * T * c * 2, 4.0 * f * c * T)
            if f == 16 and want("xattn") and c // 8 in (40, 80, 160):
                # cross-attention of the spatial transformer: every site against 77 prompt tokens, 2 prompts (CFG halves)
                qx = torch.randn(T, c, device=dev, dtype=bt)
                kvx = torch.randn(b, 77, 2 * c, device=dev, dtype=bt)
                us, mn = timeit(lambda: ops.cross_attention_core(qx, kvx[:, :, :c], kvx[:, :, c:], frames=b * f, sites=s * s, heads=8,
                                                                 out=o), args.iters, flush)
                add("cross_attn_core", f"frames{b * f} d{s * s} c{c} (hd {c // 8}) L77", us, mn, 2.0 * T * c * 2, 4.0 * 77 * c * T)
                qh = qx.reshape(b * f, s * s, 8, c // 8).transpose(1, 2)
                kh = kvx[:, :, :c].reshape(b, 77, 8, c // 8).transpose(1, 2).repeat_interleave(f, dim=0)
                vh = kvx[:, :, c:].reshape(b, 77, 8, c // 8).transpose(1, 2).repeat_interleave(f, dim=0)
                us, mn = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh), args.iters, flush)
                add("torch_sdpa(yardstick)", f"frames{b * f} d{s * s} c{c} (hd {c // 8}) L77", us, mn, 2.0 * T * c * 2, 4.0 * 77 * c * T)
            if c in ops.FUSED_TEMPORAL_WIDTHS and want("fused"):
                # the whole temporal-attention block in one launch vs the four launches it replaces
                wq_, wk_, wv_, wo_ = (torch.randn(c, c, device=dev, dtype=bt) * c ** -0.5 for _ in range(4))
                perm = ops.pack_qkv_per_head(wq_, wk_, wv_, 8)
                w3_ = torch.cat([wq_, wk_, wv_]).contiguous()
                flops = T * (8.0 * c * c + 4.0 * f * c)
                us, mn = timeit(lambda: ops.temporal_attention_fused(tok, g, be, pe, perm, wo_, be, batch=b, frames=f, sites=s * s, heads=8,
                                                                     out=yt), args.iters, flush)
                add("temporal_attn_fused", f"b{b} f{f} d{s * s} c{c}", us, mn, 2.0 * T * c * 2 + 8.0 * c * c, flops, tensor=True)

                def four():
                    n_ = ops.layernorm_pe(tok, g, be, 1e-5, pe=pe, frames=f, sites=s * s)
                    qkv_ = ops.linear(n_, w3_)
                    o_ = ops.temporal_attention_core(qkv_[:, :c], qkv_[:, c:2 * c], qkv_[:, 2 * c:], batch=b, frames=f, sites=s * s, heads=8)
                    return ops.linear(o_, wo_, be, residual=tok)
                us, mn = timeit(four, args.iters, flush)
                add("temporal_attn_4launch", f"b{b} f{f} d{s * s} c{c}", us, mn, 2.0 * T * c * 2 + 8.0 * c * c, flops, tensor=True)
            if f == 16 and want("gemm"):
                # the motion module's GEMMs: fused QKV, out-proj + residual, GEGLU, FF out
                w3 = torch.randn(3 * c, c, device=dev, dtype=bt) * c ** -0.5
                w1 = torch.randn(c, c, device=dev, dtype=bt) * c ** -0.5
                wg = torch.randn(8 * c, c, device=dev, dtype=bt) * c ** -0.5
                w2 = torch.randn(c, 4 * c, device=dev, dtype=bt) * (4 * c) ** -0.5
                bias = torch.randn(8 * c, device=dev)
                u = torch.randn(T, 4 * c, device=dev, dtype=bt)
                for nm, fn, (m_, n_, k_) in (
                        ("proj_in", lambda: ops.linear(tok, w1, bias[:c]), (T, c, c)),
                        ("qkv", lambda: ops.linear(tok, w3), (T, 3 * c, c)),
                        ("out+bias+res", lambda: ops.linear(tok, w1, bias[:c], residual=yt), (T, c, c)),
                        ("geglu", lambda: ops.linear(tok, wg, bias, geglu=True), (T, 8 * c, c)),
                        ("ff_out+res", lambda: ops.linear(u, w2, bias[:c], residual=yt), (T, c, 4 * c))):
                    us, mn = timeit(fn, args.iters, flush)
                    nb = (m_ * k_ + n_ * k_ + m_ * (n_ // 2 if nm == "geglu" else n_) * (2 if "res" in nm else 1)) * 2.0
                    add("linear_tcgen05", f"{nm} m{m_} n{n_} k{k_}", us, mn, nb, 2.0 * m_ * n_ * k_, tensor=True)
                    # cuBLAS yardstick for the same GEMM (library; not used on the product path)
                    a_ = u if nm == "ff_out+res" else tok
                    w_ = {"proj_in": w1, "qkv": w3, "out+bias+res": w1, "geglu": wg, "ff_out+res": w2}[nm]
                    us_c, mn_c = timeit(lambda: torch.nn.functional.linear(a_, w_), args.iters, flush)
                    add("cublas_linear(yardstick)", f"{nm} m{m_} n{n_} k{k_}", us_c, mn_c, nb, 2.0 * m_ * n_ * k_, tensor=True)
                    # the same arithmetic the fused kernel does, on cuBLAS + torch elementwise ops
                    bb = bias[:w_.shape[0]].to(bt)
                    if nm == "geglu":
                        def cub_epi():
                            av, gv = torch.nn.functional.linear(a_, w_, bb).chunk(2, dim=-1)
                            return av * torch.nn.functional.gelu(gv)
                    elif "res" in nm:
                        def cub_epi():
                            return torch.addmm(yt, a_, w_.t()).add_(bb)
                    else:
                        def cub_epi():
                            return torch.nn.functional.linear(a_, w_, bb)
                    us_e, mn_e = timeit(cub_epi, args.iters, flush)
                    add("cublas+epilogue(yardstick)", f"{nm} m{m_} n{n_} k{k_}", us_e, mn_e, nb, 2.0 * m_ * n_ * k_, tensor=True)
            del x, xn, tok, qkv, o, yt
            torch.cuda.empty_cache()

    # residual merge at config 2 / config 1 sizes (native layout, in place on skips)
    from oracle import synth
    for nets, (f, lat) in ((2, (16, 64)), (4, (16, 64)), (1, (8, 32))) if want("merge") else ():
        shapes = synth.residual_shapes()
        raw = [[torch.randn(b * f, ch, lat // d, lat // d, device=dev, dtype=bt).contiguous(memory_format=torch.channels_last)
                for ch, d in shapes] for _ in range(nets)]
        skips = [torch.randn_like(t) for t in raw[0]]
        sc = [[1.0] * 13 for _ in range(nets)]
        E = sum(t.numel() for t in skips)
        us, mn = timeit(lambda: ops.residual_merge(raw, sc, skips, frames=f, add_into_dst=True, layout=L.CA_LAYOUT_BFHWC), args.iters, flush)
        add("residual_merge", f"{nets} nets, 13 tensors, E={E} (b{b} f{f} lat{lat}) in-place", us, mn, (nets + 2.0) * E * 2)
        del raw, skips
        torch.cuda.empty_cache()

    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(dict(peaks=dict(hbm_gbs=hbm, bf16_tflops=tf, source=src), rows=rows), open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
