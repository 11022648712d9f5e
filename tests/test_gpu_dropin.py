"""The drop-in route on a GPU: the REFERENCE'S OWN modules (animatediff/models/*.py, modules/*.py, imported unmodified through
oracle/diffusers_shim from /root/reference or the shipped baseline/_ref) with `controlanimate_b200.install` applied.

* B1: `B200TemporalAttnProcessor` driven by the reference's `VersatileAttention.forward` (which hands the processor
  `encoder_hidden_states = hidden_states`, motion_module.py:309,321).
* B1-B4 together: the reference `UNet3DConditionModel.forward` after `install(unet, controlnet_pipeline)`, fed by the wrapped
  `MultiControlNetResidualsPipeline.__call__` (lazy residual proxies -> kernel (3) at the reference's own `skip + residual`),
  against the reference's fp32 CPU forward and next to its own bf16 eager GPU forward.
* IP-Adapter: `B200IPAttnProcessor` against the oracle (pinned to the reference's IPAttnProcessor2_0 by tests/golden/ip_adapter.npz),
  alone and inside the native UNet at config-4 shape (f = 32, odd pyramid, 77 + 4 tokens); the temporal processors survive
  `set_ip_adapter`-style overwrites.
"""
import types

import pytest
import torch

from oracle import ref_import
from oracle import ref_ops as R
from oracle import ref_unet3d as U
from oracle import synth

pytestmark = pytest.mark.gpu
SEED = 55


def cosine(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()))


def r16(t):
    return t.bfloat16().float()


@pytest.fixture(scope="module")
def ref():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    if ref_import.reference_root() is None:
        pytest.fail("the reference sources were not shipped: run __graft_entry__.build() in the build container (baseline/_ref)")
    ref_import.import_reference()
    from controlanimate_b200 import _lib
    _lib.load(build_if_missing=False)
    import animatediff.models.unet as ref_unet
    import modules.attention_processor as ref_proc
    import modules.controlresiduals_pipeline as ref_cn
    return types.SimpleNamespace(unet=ref_unet, proc=ref_proc, cn=ref_cn)


def small_cfg():
    cfg = synth.unet_config(tiny=True)
    cfg.update(block_out_channels=(64, 128, 256, 256), cross_attention_dim=64)
    return cfg


def _residual_shapes(cfg, hh, ww):
    out, sh, sw, div_prev = [], hh, ww, 1
    for ch, div in synth.residual_shapes(cfg["block_out_channels"]):
        while div_prev < div:
            sh, sw = (sh + 1) // 2, (sw + 1) // 2
            div_prev *= 2
        out.append((ch, sh, sw))
    return out


class _FakeDiffusersNet:
    """Stands in for a diffusers ControlNetModel (third party, not installed): returns pre-made RAW residuals, multiplied by
    the conditioning scale it is called with, as diffusers does."""

    def __init__(self, raw):
        self.raw = raw

    def __call__(self, sample, t, encoder_hidden_states=None, controlnet_cond=None, conditioning_scale=1.0, guess_mode=False,
                 return_dict=False):
        assert not guess_mode
        res = [r * conditioning_scale for r in self.raw]
        return res[:-1], res[-1]


class _FakeMulti:
    """diffusers MultiControlNetModel.forward: sum of the nets' scaled residuals."""

    def __init__(self, nets):
        self.nets = nets

    def __call__(self, sample, t, encoder_hidden_states=None, controlnet_cond=None, conditioning_scale=None, guess_mode=False,
                 return_dict=False):
        down = mid = None
        for k, net in enumerate(self.nets):
            d, m = net(sample, t, encoder_hidden_states, controlnet_cond[k], conditioning_scale[k], guess_mode)
            down = d if down is None else [a + b for a, b in zip(down, d)]
            mid = m if mid is None else mid + m
        return down, mid


def test_processor_through_reference_versatile_attention(ref):
    """install(unet, motion_modules=False, resnets=False): only the processors are swapped, so the reference's own
    VersatileAttention.forward (rearranges, PE, dead Q/K/V, `encoder_hidden_states = hidden_states`) drives the B200 processor."""
    from controlanimate_b200.install import install, verify_installed
    cfg = small_cfg()
    unet = ref.unet.UNet3DConditionModel(**cfg)
    synth.fill_module_(unet, SEED)
    with torch.no_grad():
        for p in unet.parameters():
            p.copy_(r16(p))
    unet.set_attn_processor(ref.proc.AttnProcessor2_0())
    unet.eval()
    sample = r16(synth.tensor(SEED, "va.sample", (2, 4, 8, 16, 16)))
    ctx = r16(synth.tensor(SEED, "va.ctx", (2, 7, 64)))
    with torch.no_grad():
        want = unet(sample, 501, encoder_hidden_states=ctx).sample
        g = unet.cuda().bfloat16()
        counts = install(g, motion_modules=False, resnets=False)
        assert counts["processors"] == 42 and counts["motion_modules"] == 0 and verify_installed(g)
        got = g(sample.cuda().bfloat16(), 501, encoder_hidden_states=ctx.cuda().bfloat16()).sample
    c = cosine(got, want)
    print(f"[drop-in B1] reference UNet3D with B200 temporal processors only: cosine {c:.6f}")
    assert c >= 0.999, c


def test_reference_unet_forward_with_install(ref):
    """Boundaries B1-B4 through the reference's own UNet3DConditionModel.forward and MultiControlNetResidualsPipeline."""
    import time
    from controlanimate_b200.install import install
    from controlanimate_b200.residuals import LazyResidual
    cfg = small_cfg()
    b, f, hh, ww = 2, 8, 16, 16
    unet = ref.unet.UNet3DConditionModel(**cfg)
    synth.fill_module_(unet, SEED)
    with torch.no_grad():
        for p in unet.parameters():
            p.copy_(r16(p))
    unet.set_attn_processor(ref.proc.AttnProcessor2_0())
    unet.eval()
    sample = r16(synth.tensor(SEED, "di.sample", (b, 4, f, hh, ww)))
    ctx = r16(synth.tensor(SEED, "di.ctx", (b, 7, 64)))
    raws = [[r16(synth.tensor(SEED, f"di.raw{k}.{i}", (b * f, ch, sh, sw), 0.1)) for i, (ch, sh, sw) in enumerate(_residual_shapes(cfg, hh, ww))]
            for k in range(2)]
    scales = [1.0, 0.5]

    def pipe_on(device, dtype):
        p = object.__new__(ref.cn.MultiControlNetResidualsPipeline)        # no checkpoints to load offline
        p.controlnet = _FakeMulti([_FakeDiffusersNet([t.to(device, dtype) for t in raw]) for raw in raws])
        p.prep_images, p.cond_scale = [None, None], scales
        return p

    with torch.no_grad():
        # the reference end to end, fp32 on the CPU: its pipeline's __call__ (scale, sum, rearrange) + its UNet forward
        cpu_pipe = pipe_on("cpu", torch.float32)
        cpu_pipe.controlnet = _Unhalf(cpu_pipe.controlnet)
        d_cpu, m_cpu = cpu_pipe(sample, 501, ctx, f, guess_mode=False)
        want = unet(sample, 501, encoder_hidden_states=ctx, down_block_additional_residuals=d_cpu, mid_block_additional_residual=m_cpu).sample
        # the reference on the GPU in bf16 (its stock path), then the same objects with install() applied
        g = unet.cuda().bfloat16()
        sample_g, ctx_g = sample.cuda().bfloat16(), ctx.cuda().bfloat16()
        gp = pipe_on("cuda", torch.bfloat16)
        gp.controlnet = _Unhalf(gp.controlnet)
        d16, m16 = gp(sample_g, 501, ctx_g, f, guess_mode=False)
        stock = g(sample_g, 501, encoder_hidden_states=ctx_g, down_block_additional_residuals=d16, mid_block_additional_residual=m16).sample
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            d16, m16 = gp(sample_g, 501, ctx_g, f, guess_mode=False)
            g(sample_g, 501, encoder_hidden_states=ctx_g, down_block_additional_residuals=d16, mid_block_additional_residual=m16)
        torch.cuda.synchronize()
        t_stock = (time.perf_counter() - t0) / 3

        gp2 = pipe_on("cuda", torch.bfloat16)
        counts = install(g, gp2)
        assert counts == dict(processors=42, motion_modules=21, resnets=22, controlnet_pipeline=1)
        d_lazy, m_lazy = gp2(sample_g, 501, ctx_g, f, guess_mode=False)
        assert len(d_lazy) == 12 and all(isinstance(r, LazyResidual) for r in d_lazy) and isinstance(m_lazy, LazyResidual)
        got = g(sample_g, 501, encoder_hidden_states=ctx_g, down_block_additional_residuals=d_lazy, mid_block_additional_residual=m_lazy).sample
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            d_lazy, m_lazy = gp2(sample_g, 501, ctx_g, f, guess_mode=False)
            g(sample_g, 501, encoder_hidden_states=ctx_g, down_block_additional_residuals=d_lazy, mid_block_additional_residual=m_lazy)
        torch.cuda.synchronize()
        t_own = (time.perf_counter() - t0) / 3
        # the lazy proxies hold what the reference's merged tensors hold
        assert cosine(d_lazy[3].materialize(), d_cpu[3]) >= 0.9999 and cosine(m_lazy.materialize(), m_cpu) >= 0.9999
    c_own, c_stock = cosine(got, want), cosine(stock, want)
    print(f"[drop-in B1-B4] reference UNet3D.forward + MultiControlNetResidualsPipeline with install(): cosine {c_own:.6f} vs the "
          f"reference's fp32 CPU forward (its own bf16 eager GPU forward: {c_stock:.6f}); small-size wall time per forward "
          f"{t_own * 1e3:.1f} ms vs {t_stock * 1e3:.1f} ms stock eager")
    assert got.shape == want.shape
    assert c_own >= 0.999, c_own


class _Unhalf:
    """The reference hard-codes `.half()` on the ControlNet inputs (controlresiduals_pipeline.py:295,297); the recorded nets
    ignore the inputs, so this only keeps the dtypes of the returned residuals as they are."""

    def __init__(self, multi):
        self.multi, self.nets = multi, multi.nets

    def __call__(self, *a, **kw):
        return self.multi(*a, **kw)


@pytest.mark.parametrize("c,cross,L,scale", [(320, 768, 81, 0.6), (640, 768, 81, 1.0), (1280, 768, 20, 0.3)])
def test_ip_adapter_processor(ref, c, cross, L, scale):
    """B200IPAttnProcessor through the AttentionProcessor protocol == IPAttnProcessor2_0 (attention_processor.py:367-492)."""
    from controlanimate_b200 import layers
    n, d, heads, ntok = 4, 96, 8, 4
    attn = layers._SpatialAttention(c, heads, cross)
    sd = {k: r16(v) for k, v in U.synth_state_dict({"to_q.weight": (c, c), "to_k.weight": (c, cross), "to_v.weight": (c, cross),
                                                     "to_out.0.weight": (c, c), "to_out.0.bias": (c,)}, SEED).items()}
    attn.load_state_dict(sd, strict=False)
    attn = attn.cuda().bfloat16()
    proc = layers.B200IPAttnProcessor(hidden_size=c, cross_attention_dim=cross, scale=scale, num_tokens=ntok)
    ip = {k: r16(v) for k, v in U.synth_state_dict({"to_k_ip.weight": (c, cross), "to_v_ip.weight": (c, cross)}, SEED + 1).items()}
    proc.load_state_dict(ip)
    proc = proc.cuda().bfloat16()
    x = r16(synth.tensor(SEED, f"ipp.x{c}", (n, d, c)))
    ctx = r16(synth.tensor(SEED, f"ipp.ctx{c}", (n, L, cross)))
    want = R.ip_attention_processor(x, ctx, sd["to_q.weight"], sd["to_k.weight"], sd["to_v.weight"], sd["to_out.0.weight"],
                                    sd["to_out.0.bias"], ip["to_k_ip.weight"], ip["to_v_ip.weight"], heads, ntok, scale)
    with torch.no_grad():
        got = proc(attn, x.cuda().bfloat16(), encoder_hidden_states=ctx.cuda().bfloat16())
    assert got.shape == want.shape and cosine(got, want) >= 0.9999, cosine(got, want)
    # the image tokens matter: without them the result differs measurably
    plain = R.attention_processor(x, sd["to_q.weight"], sd["to_k.weight"], sd["to_v.weight"], sd["to_out.0.weight"], sd["to_out.0.bias"],
                                  heads, ctx[:, :L - ntok])
    assert cosine(want, plain) < 0.9995


def test_unet3d_config4_with_ip_adapter(ref):
    """Config 4 end to end on the native UNet: f = 32, 12x7 latents (odd pyramid), b = 1, prompt of 77 + 4 image tokens, every
    cross-attention on B200IPAttnProcessor (install.set_ip_adapter), temporal processors overwritten by a plain AttnProcessor2_0
    as modules/ip_adapter.py:95-126 does — they must survive — against the oracle with the same IP weights."""
    from controlanimate_b200 import unet as un
    from controlanimate_b200.install import set_ip_adapter
    from controlanimate_b200.layers import B200TemporalAttnProcessor
    cfg = synth.unet_config(tiny=False)
    net = un.UNet3DConditionModel(**cfg)
    sd = {k: (v if k.endswith(".pe") else r16(v)) for k, v in U.synth_state_dict(U.unet3d_shapes(cfg), SEED).items()}
    net.load_state_dict(sd, strict=True)
    net = net.cuda().bfloat16().eval()
    procs = set_ip_adapter(net, scale=0.7, num_tokens=4)
    assert len(procs) == 16
    ip_sd = {k: r16(v) for k, v in U.synth_state_dict(U.ip_adapter_shapes(cfg), SEED + 1).items()}
    net.load_state_dict(ip_sd, strict=False)
    # what set_ip_adapter does to the other 74 processors
    plain = ref.proc.AttnProcessor2_0()
    for name, m in net.named_modules():
        if hasattr(m, "set_processor") and not name.endswith("attn2"):
            m.set_processor(plain)
    temporal = [p for k, p in net.attn_processors.items() if "motion_modules" in k]
    assert len(temporal) == 42 and all(isinstance(p, B200TemporalAttnProcessor) for p in temporal)
    b, f, hh, ww = 1, 32, 7, 12
    sample = r16(synth.tensor(SEED, "c4ip.sample", (b, 4, f, hh, ww)))
    ctx = r16(synth.tensor(SEED, "c4ip.ctx", (b, 81, 768)))
    want = U.unet3d_forward(sd, cfg, sample, 251, ctx, ip=dict(sd=ip_sd, num_tokens=4, scale=0.7))
    want_noip = U.unet3d_forward(sd, cfg, sample, 251, ctx)
    with torch.no_grad():
        got = net(sample.cuda(), 251, ctx.cuda()).sample
    c = cosine(got, want)
    print(f"[config 4 + IP-Adapter] native UNet3D f=32 12x7 81 tokens: cosine {c:.6f} (oracle with vs without IP: {cosine(want, want_noip):.4f})")
    assert c >= 0.999, c
