"""GPU parity at FULL WIDTH (block_out_channels 320/640/1280/1280: head_dim 40/80/160, the specialised attention kernels,
`cross_attn.cu`, every tcgen05 tiling of the real projections) against the CPU oracle on the same seeded inputs.

Spatial sizes are reduced (16x16 latents for config 2, 12x7 for config 4's odd pyramid) so that the fp32 CPU oracle
finishes in seconds; the frame count, widths, head sizes, prompt lengths, ControlNet count and CFG batch are the
BASELINE.json ones.  Tolerances are north_star's: UNet noise-prediction cosine >= 0.999 per step.  Each test prints the
plain numbers it measured (`pytest -s` / the captured log) so that DESIGN.md can quote them.

Also here: the model-level fp16 case (the reference runs `.half()`, modules/controlanimate_pipeline.py:108-110) and the LCM
`timestep_cond` path (unet.py:526-534), at tiny width.
"""
import pytest
import torch

from oracle import ref_ops as R
from oracle import ref_unet3d as U
from oracle import synth

pytestmark = pytest.mark.gpu
SEED = 91


@pytest.fixture(scope="module")
def ca():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    import controlanimate_b200.layers as Ly
    import controlanimate_b200.pipeline as P
    import controlanimate_b200.unet as Un
    from controlanimate_b200 import _lib
    _lib.load(build_if_missing=False)

    class NS:
        layers, pipeline, unet = Ly, P, Un
    return NS


def r16(t, dtype=torch.bfloat16):
    return t.to(dtype).float()


def cosine(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()))


def rel_rms(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm())


def load_synth(module, shapes, seed, dtype=torch.bfloat16):
    sd = U.synth_state_dict(shapes, seed)
    sd = {k: (v if k.endswith(".pe") else r16(v, dtype)) for k, v in sd.items()}  # oracle and product see the SAME weights
    module.load_state_dict(sd, strict=True)
    return sd


@pytest.fixture(scope="module")
def full(ca):
    """Full-width UNet3D + two ControlNets, built once for the module (1.3 G + 2 x 0.36 G parameters)."""
    cfg = synth.unet_config(tiny=False)
    unet = ca.unet.UNet3DConditionModel(**cfg)
    sd_u = load_synth(unet, U.unet3d_shapes(cfg), SEED)
    unet = unet.cuda().bfloat16().eval()
    nets, sds = [], []
    for k in range(2):
        cn = ca.unet.ControlNetModel()
        sds.append(load_synth(cn, U.controlnet_shapes(cfg), SEED + 1 + k))
        nets.append(cn.cuda().bfloat16().eval())
    return cfg, unet, sd_u, nets, sds


def test_unet3d_full_width_config2(ca, full):
    """UNet3D forward (unet.py:458-621), b=2 CFG halves, f=16, 77 prompt tokens, C=(320,640,1280,1280), 16x16 latents."""
    cfg, unet, sd, _, _ = full
    b, f, hh = 2, 16, 16
    sample = r16(synth.tensor(SEED, "fw.sample", (b, 4, f, hh, hh)))
    ctx = r16(synth.tensor(SEED, "fw.ctx", (b, 77, 768)))
    ref = U.unet3d_forward(sd, cfg, sample, 501, ctx)
    with torch.no_grad():
        y = unet(sample.cuda(), 501, ctx.cuda()).sample
    c, e = cosine(y, ref), rel_rms(y, ref)
    print(f"[full width] UNet3D config-2 noise prediction: cosine {c:.6f}, relative rms error {e:.4f}")
    assert y.shape == ref.shape
    assert c >= 0.999, c
    assert e <= 4e-2, e


def test_denoising_step_full_width_config2(ca, full):
    """One step of the hot loop (controlanimation_pipeline.py:793-849) at full width: 2 ControlNets (scales 1.0 / 0.5) ->
    single-pass residual merge -> UNet3D -> CFG 7.5 -> DDIM.  The UNet prediction is held to the 0.999 contract; the
    CFG-combined tensor is REPORTED next to what plain torch bf16 eager (the oracle's own op sequence run in bf16 on the
    GPU, i.e. the reference's 16-bit path) achieves against the same fp32 oracle: eps_u + 7.5 (eps_c - eps_u) amplifies the
    16-bit rounding of the two UNet rows for every 16-bit implementation."""
    cfg, unet, sd_u, nets, sds = full
    f, hh, t, g = 16, 16, 501, 7.5
    cond_scale = [1.0, 0.5]
    latents = r16(synth.tensor(SEED, "fs.lat", (1, 4, f, hh, hh)))
    prompt = r16(synth.tensor(SEED, "fs.ctx", (2, 77, 768)))
    images = [r16(synth.tensor(SEED, f"fs.img{k}", (2 * f, 3, hh * 8, hh * 8), 0.5)) for k in range(2)]

    def oracle(sd_unet, sd_nets, lat, prm, imgs):
        model_in = torch.cat([lat] * 2)
        x2d = model_in.permute(0, 2, 1, 3, 4).reshape(2 * f, 4, hh, hh)
        ctx_tiled = torch.cat([prm] * f)                       # controlresiduals_pipeline.py:292
        per_net = [U.controlnet_forward(sd_nets[k], cfg, x2d, t, ctx_tiled, imgs[k]) for k in range(2)]
        down, mid = R.merge_controlnet_residuals(per_net, cond_scale, f)
        return per_net, U.unet3d_forward(sd_unet, cfg, model_in, t, prm, down, mid)

    per_net, noise = oracle(sd_u, sds, latents, prompt, images)
    ref_eps = R.cfg_combine(noise, g)

    mc = ca.pipeline.MultiControlNetResiduals(nets, cond_scale)
    mc.prep_images = [im.cuda().bfloat16() for im in images]
    sched = ca.pipeline.DDIMScheduler()
    sched.set_timesteps(20)
    loop = ca.pipeline.DenoisingLoop(unet, mc, sched, guidance_scale=g)
    lat2 = torch.cat([latents] * 2).cuda().bfloat16()
    with torch.no_grad():
        raw = mc.raw(lat2, t, prompt.cuda().bfloat16(), f)
        worst_raw = min(cosine(raw[k][i], per_net[k][i]) for k in range(2) for i in range(13))
        rs, _ = mc(lat2, t, prompt.cuda().bfloat16(), f, guess_mode=False)
        n_own = unet(lat2, t, prompt.cuda().bfloat16(), down_block_additional_residuals=rs).sample
        eps_own = loop.predict_noise(latents.cuda().bfloat16(), t, prompt.cuda().bfloat16())
    c_noise, c_eps = cosine(n_own, noise), cosine(eps_own, ref_eps)

    # the same op sequence in stock torch bf16 on the GPU (the 16-bit yardstick)
    to16 = lambda sd: {k: (v.cuda() if k.endswith(".pe") else v.cuda().bfloat16()) for k, v in sd.items()}  # noqa: E731
    R.USE_SDPA = True
    try:
        with torch.no_grad():
            _, n_t16 = oracle(to16(sd_u), [to16(s) for s in sds], latents.cuda().bfloat16(), prompt.cuda().bfloat16(),
                              [im.cuda().bfloat16() for im in images])
    finally:
        R.USE_SDPA = False
    c_noise_t16, c_eps_t16 = cosine(n_t16, noise), cosine(R.cfg_combine(n_t16.float().cpu(), g), ref_eps)
    print(f"[full width] step config 2: worst raw ControlNet residual cosine {worst_raw:.6f}; UNet noise cosine own {c_noise:.6f} "
          f"(torch bf16 eager {c_noise_t16:.6f}); CFG-combined (g = 7.5) cosine own {c_eps:.6f} (torch bf16 eager {c_eps_t16:.6f})")
    assert worst_raw >= 0.999, worst_raw
    assert c_noise >= 0.999, c_noise                       # north_star: noise-prediction cosine >= 0.999 per step
    # the guided combination: at least as good as stock torch bf16 (minus a small margin), and never below 0.995
    assert c_eps >= min(0.999, c_eps_t16 - 2e-3) and c_eps >= 0.995, (c_eps, c_eps_t16)


def test_unet3d_full_width_config4_odd_pyramid(ca, full):
    """BASELINE config 4 at full width: f = 32 (= PE max_len), non-square 12x7 latents (pyramid 7 -> 4 -> 2 -> 1 and
    12 -> 6 -> 3 -> 2: `forward_upsample_size`, unet.py:491-499,596-597; resnet.py:68-69), b = 1, IP-Adapter-length prompt of
    81 tokens (plain cross-attention here; the dual-KV processor has its own test)."""
    cfg, unet, sd, _, _ = full
    b, f, hh, ww = 1, 32, 7, 12
    sample = r16(synth.tensor(SEED, "c4.sample", (b, 4, f, hh, ww)))
    ctx = r16(synth.tensor(SEED, "c4.ctx", (b, 81, 768)))
    ref = U.unet3d_forward(sd, cfg, sample, 251, ctx)
    with torch.no_grad():
        y = unet(sample.cuda(), 251, ctx.cuda()).sample
    c, e = cosine(y, ref), rel_rms(y, ref)
    print(f"[full width] UNet3D config-4 (f=32, 12x7, 81 tokens): cosine {c:.6f}, relative rms error {e:.4f}")
    assert y.shape == ref.shape and c >= 0.999 and e <= 4e-2, (c, e)


def _tiny_cfg(**kw):
    cfg = synth.unet_config(tiny=True)
    cfg.update(block_out_channels=(64, 128, 256, 256), cross_attention_dim=64)
    cfg.update(kw)
    return cfg


def test_unet3d_fp16_model(ca):
    """The reference's own GPU dtype (`.half()`): whole UNet3D in fp16 storage / fp32 accumulate against the fp32 oracle."""
    cfg = _tiny_cfg()
    unet = ca.unet.UNet3DConditionModel(**cfg)
    sd = load_synth(unet, U.unet3d_shapes(cfg), SEED, torch.float16)
    unet = unet.cuda().half().eval()
    sample = r16(synth.tensor(SEED, "h.sample", (2, 4, 8, 16, 12)), torch.float16)
    ctx = r16(synth.tensor(SEED, "h.ctx", (2, 77, 64)), torch.float16)
    ref = U.unet3d_forward(sd, cfg, sample, 751, ctx)
    with torch.no_grad():
        y = unet(sample.cuda(), 751, ctx.cuda()).sample
    c, e = cosine(y, ref), rel_rms(y, ref)
    print(f"[fp16] UNet3D tiny width: cosine {c:.6f}, relative rms error {e:.4f}")
    assert y.dtype == torch.float16 and c >= 0.9995 and e <= 2e-2, (c, e)


def test_unet3d_timestep_cond_lcm(ca):
    """LCM guidance embedding: `timestep_cond` -> time_embedding.cond_proj (unet.py:526-534; pipeline :477-498,823-833),
    b = 1 (config 3 runs without CFG duplication)."""
    cfg = _tiny_cfg(time_cond_proj_dim=256)
    unet = ca.unet.UNet3DConditionModel(**cfg)
    sd = load_synth(unet, U.unet3d_shapes(cfg), SEED)
    unet = unet.cuda().bfloat16().eval()
    sample = r16(synth.tensor(SEED, "l.sample", (1, 4, 16, 8, 8)))
    ctx = r16(synth.tensor(SEED, "l.ctx", (1, 77, 64)))
    w = r16(synth.tensor(SEED, "l.w", (1, 256)))
    ref = U.unet3d_forward(sd, cfg, sample, 759, ctx, timestep_cond=w)
    ref_nocond = U.unet3d_forward(sd, cfg, sample, 759, ctx)
    with torch.no_grad():
        y = unet(sample.cuda(), 759, ctx.cuda(), timestep_cond=w.cuda().bfloat16()).sample
    c = cosine(y, ref)
    print(f"[lcm] UNet3D with timestep_cond: cosine {c:.6f} (cosine of the oracle with vs without the embedding: {cosine(ref, ref_nocond):.4f})")
    assert c >= 0.999, c
    assert cosine(ref, ref_nocond) < 0.9999   # the guidance embedding matters, so the agreement above is not vacuous


def test_denoising_step_config3_lcm_four_controlnets(ca):
    """BASELINE config 3, one window-step: LCM branch (b = 1, guidance through timestep_cond; controlanimation_pipeline.py:
    770-771, 823-833), FOUR ControlNets with SampleConfig's scales [1.0, 0.35, 1.0, 0.4] merged by kernel (3) in one pass,
    16 frames — against the oracle (tiny width: four full-width nets would only repeat test_denoising_step_full_width)."""
    cfg = _tiny_cfg(time_cond_proj_dim=256)
    f, hh, t, w_guid = 16, 16, 759, 1.1
    scales = [1.0, 0.35, 1.0, 0.4]
    unet = ca.unet.UNet3DConditionModel(**cfg)
    sd_u = load_synth(unet, U.unet3d_shapes(cfg), SEED)
    unet = unet.cuda().bfloat16().eval()
    nets, sds = [], []
    for k in range(4):
        cn = ca.unet.ControlNetModel(block_out_channels=cfg["block_out_channels"], cross_attention_dim=cfg["cross_attention_dim"])
        sds.append(load_synth(cn, U.controlnet_shapes(cfg), SEED + 10 + k))
        nets.append(cn.cuda().bfloat16().eval())
    latents = r16(synth.tensor(SEED, "c3.lat", (1, 4, f, hh, hh)))
    prompt = r16(synth.tensor(SEED, "c3.ctx", (1, 77, cfg["cross_attention_dim"])))
    images = [r16(synth.tensor(SEED, f"c3.img{k}", (f, 3, hh * 8, hh * 8), 0.5)) for k in range(4)]
    w_emb = ca.pipeline.get_w_embedding(torch.full((1,), w_guid), 256)
    # oracle
    x2d = latents.permute(0, 2, 1, 3, 4).reshape(f, 4, hh, hh)
    per_net = [U.controlnet_forward(sds[k], cfg, x2d, t, torch.cat([prompt] * f), images[k]) for k in range(4)]
    down, mid = R.merge_controlnet_residuals(per_net, scales, f)
    want = U.unet3d_forward(sd_u, cfg, latents, t, prompt, down, mid, timestep_cond=r16(w_emb))
    # product
    mc = ca.pipeline.MultiControlNetResiduals(nets, scales)
    mc.prep_images = [im.cuda().bfloat16() for im in images]
    sched = ca.pipeline.DDIMScheduler()
    sched.set_timesteps(4)
    loop = ca.pipeline.DenoisingLoop(unet, mc, sched, guidance_scale=w_guid, use_lcm=True)
    with torch.no_grad():
        got = loop.predict_noise(latents.cuda().bfloat16(), t, prompt.cuda().bfloat16())
        nxt = loop.step(latents.cuda(), t, prompt.cuda().bfloat16())
    c = cosine(got, want)
    print(f"[config 3] LCM step, 4 ControlNets, b=1, f=16: model-output cosine {c:.6f}")
    assert got.shape == (1, 4, f, hh, hh) and c >= 0.999, c
    assert torch.isfinite(nxt).all() and nxt.shape == latents.shape
