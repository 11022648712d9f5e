"""CPU tests: the C-ABI library loads and exports every symbol the header declares; host-side logic."""
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    """No compute calls (there is no GPU here): dlopen + symbol/signature table against include/controlanimate_b200.h."""
    from controlanimate_b200 import _lib, build
    build.build()
    lib = _lib.load(build_if_missing=False)
    header = open(os.path.join(ROOT, "include", "controlanimate_b200.h")).read()
    declared = set(re.findall(r"\b(ca_[a-z0-9_]+)\s*\(", header))
    declared -= {"ca_status_t", "ca_dtype_t", "ca_layout_t", "ca_epilogue_t"}
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.ca_version()
    # argument validation happens before any CUDA call -> exercisable on CPU
    assert lib.ca_groupnorm_silu(None, None, None, None, None, 0, 1, 32, 1, 1, 1, 32, 1e-5, 1, 1, 0, 0, None, 0, None) == 1
    assert b"null pointer" in lib.ca_last_error()
    assert lib.ca_groupnorm_workspace_bytes(2, 320, 16, 64, 64, 32, 1, 0, 0) > 0        # 80 KB groups -> 2 chunks + counters
    assert lib.ca_groupnorm_workspace_bytes(2, 1280, 16, 8, 8, 32, 1, 0, 0) == 0        # 5 KB groups -> single chunk


def test_product_never_imports_the_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "controlanimate_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
            assert "ref_ops" not in src and "ref_unet3d" not in src, fn


def test_ops_refuse_cpu_tensors():
    from controlanimate_b200 import ops
    x = torch.zeros(1, 32, 1, 4, 4, dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="no CPU fallback"):
        ops.groupnorm_silu(x, torch.ones(32), torch.zeros(32), 32, 1e-5)
    with pytest.raises(ValueError, match="no CPU fallback"):
        ops.linear(torch.zeros(4, 16, dtype=torch.bfloat16), torch.zeros(16, 16, dtype=torch.bfloat16))


def test_layout_helpers_are_views():
    from controlanimate_b200 import _lib as L, layers as Ly, ops
    x = torch.randn(2, 8, 3, 4, 5)
    n = Ly.to_native(x)
    assert ops.video_layout(x) == L.CA_LAYOUT_NCFHW and ops.video_layout(n) == L.CA_LAYOUT_BFHWC and torch.equal(n, x)
    x4 = Ly.frames4(n)
    assert x4.data_ptr() == n.data_ptr() and x4.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(x4, x.permute(0, 2, 1, 3, 4).reshape(6, 8, 4, 5))
    tok = Ly.tokens(x4)
    assert tok.data_ptr() == n.data_ptr() and tok.is_contiguous() and tok.shape == (2 * 3 * 4 * 5, 8)
    assert Ly.video5(x4, 3).data_ptr() == n.data_ptr() and torch.equal(Ly.video5(x4, 3), x)
    assert Ly.from_tokens(tok, 6, 4, 5).data_ptr() == n.data_ptr()
    with pytest.raises(ValueError):
        ops.video_layout(x[:, :, :, ::2])


def test_scheduler_and_scales_match_oracle():
    from controlanimate_b200 import pipeline, residuals
    from oracle import ref_ops as R
    s = pipeline.DDIMScheduler()
    assert s.set_timesteps(4) == R.ddim_timesteps(4) == [751, 501, 251, 1]
    assert s.set_timesteps(20)[:2] == [951, 901] and s.timesteps[-1] == 1
    g = torch.Generator().manual_seed(0)
    x, e = torch.randn(1, 4, 2, 3, 3, generator=g), torch.randn(1, 4, 2, 3, 3, generator=g)
    for t in s.timesteps:
        assert torch.allclose(s.step(e, t, x), R.ddim_step(e, t, x, R.ddim_alphas_cumprod(), 20), atol=1e-5)
    sc = residuals._scales([1.0, 0.5], 13, guess_mode=True)
    lv = torch.logspace(-1, 0, 13)
    assert abs(sc[1][0] - 0.05) < 1e-7 and abs(sc[0][12] - 1.0) < 1e-7 and abs(sc[1][6] - 0.5 * float(lv[6])) < 1e-7
    assert residuals._scales([0.35], 13, guess_mode=False) == [[0.35] * 13]


def test_state_dict_keys_match_reference_tables():
    """Same keys/shapes as the reference modules (so load_weights, util.py:117-120, keeps working)."""
    from controlanimate_b200 import layers as Ly, unet as Un
    from oracle import ref_unet3d as U, synth
    mm = Ly.B200MotionModule(in_channels=64, **synth.MOTION_MODULE_KWARGS_V2)
    want = U.motion_module_shapes("", 64, 32)
    assert {k: tuple(v.shape) for k, v in mm.state_dict().items()} == want
    assert float(mm.temporal_transformer.proj_out.weight.abs().max()) == 0.0        # zero_initialize (motion_module.py:76-77)
    cfg = synth.unet_config(tiny=True)
    un = Un.UNet3DConditionModel(**cfg)
    assert {k: tuple(v.shape) for k, v in un.state_dict().items()} == U.unet3d_shapes(cfg)
    procs = un.attn_processors
    assert len(procs) == 90 and sum("motion_modules" in k for k in procs) == 42
    assert "down_blocks.0.motion_modules.0.temporal_transformer.transformer_blocks.0.attention_blocks.1.processor" in procs
    un.set_attn_processor(dict(procs))
    with pytest.raises(ValueError):
        un.set_attn_processor({k: v for k, v in list(procs.items())[:5]})
    # PE table stays fp32 when the model is cast (the kernel adds it in fp32)
    un = un.to(torch.bfloat16)
    pe = un.mid_block.motion_modules[0].temporal_transformer.transformer_blocks[0].attention_blocks[0].pos_encoder.pe
    assert pe.dtype == torch.float32 and un.conv_in.weight.dtype == torch.bfloat16


def _have_reference():
    from oracle import ref_import
    return ref_import.reference_root() is not None


@pytest.mark.skipif(not _have_reference(), reason="needs the reference sources (/root/reference or baseline/_ref)")
def test_install_into_unmodified_reference_unet():
    """install() swaps the reference's own modules for B200 ones with identical state_dicts and keeps the 90-entry
    processor protocol (boundary B1/B2/B4).  Structural only: there is no GPU here and no CPU fallback to run."""
    from oracle import ref_import
    shim = os.path.join(ROOT, "oracle", "diffusers_shim")
    ref_root = ref_import.import_reference()
    try:
        from animatediff.models.unet import UNet3DConditionModel as RefUNet
        from controlanimate_b200.install import install, verify_installed
        from controlanimate_b200.layers import B200MotionModule, B200ResnetBlock3D, B200TemporalAttnProcessor
        from oracle import synth
        ref = RefUNet(**synth.unet_config(tiny=True))
        synth.fill_module_(ref, 1)
        before = {k: v.clone() for k, v in ref.state_dict().items()}
        counts = install(ref)
        assert counts == dict(processors=42, motion_modules=21, resnets=22, controlnet_pipeline=0)
        after = ref.state_dict()
        assert set(after) == set(before) and all(torch.equal(after[k], before[k]) for k in before)
        assert len(ref.attn_processors) == 90 and verify_installed(ref)
        assert sum(isinstance(m, B200MotionModule) for m in ref.modules()) == 21
        assert sum(isinstance(m, B200ResnetBlock3D) for m in ref.modules()) == 22
        temporal = [p for k, p in ref.attn_processors.items() if "motion_modules" in k]
        assert all(isinstance(p, B200TemporalAttnProcessor) for p in temporal)
    finally:
        for p in (shim, ref_root):
            if p in sys.path:
                sys.path.remove(p)
        for m in [m for m in sys.modules if m.split(".")[0] in ("diffusers", "animatediff", "modules", "controlnet_aux")]:
            del sys.modules[m]


def test_conv_weight_is_relaid_once_per_version():
    """layers._weight_cl hands cuDNN a channels_last filter without re-laying it on every call (r01c: 158 copies per step),
    and notices in-place weight updates (load_state_dict / LoRA merges bump the version counter)."""
    from controlanimate_b200 import layers
    conv = torch.nn.Conv2d(8, 16, 3, padding=1, bias=False)
    w1 = layers._weight_cl(conv)
    assert w1.is_contiguous(memory_format=torch.channels_last) and torch.equal(w1, conv.weight)
    assert layers._weight_cl(conv) is w1                                  # cached
    with torch.no_grad():
        conv.weight.mul_(2.0)                                             # in-place update -> new version
    w2 = layers._weight_cl(conv)
    assert w2 is not w1 and torch.equal(w2, conv.weight)
    conv.load_state_dict({"weight": torch.ones_like(conv.weight)})
    assert torch.equal(layers._weight_cl(conv), torch.ones_like(conv.weight))
    one = torch.nn.Conv2d(8, 16, 1, bias=False)                           # 1x1 filters already satisfy channels_last
    assert layers._weight_cl(one) is one.weight


def test_ctx_map_is_converted_once():
    """layers._ctx_i32: the frame -> prompt index tensor reaches the C ABI as int32 and is converted once per tensor."""
    from controlanimate_b200 import layers
    assert layers._ctx_i32(None) is None
    m = torch.arange(6) % 3
    a = layers._ctx_i32(m)
    assert a.dtype == torch.int32 and a.tolist() == [0, 1, 2, 0, 1, 2]
    assert layers._ctx_i32(m) is a


def test_fold_layernorm_algebra():
    """ops.fold_layernorm (host side of ca_linear_ln): with w_gain = w * gamma, colsum = sum_k w_gain and shift = (beta + pe) w^T
    + bias, rstd * (x w_gain^T - mean * colsum) + shift[frame] equals Linear(LayerNorm(x) + pe[frame]) (motion_module.py:214-215,
    285-288 -> :321) — checked in fp32 on the CPU, where no kernel is involved."""
    import torch
    from controlanimate_b200 import ops
    g = torch.Generator().manual_seed(0)
    T, k, n, frames, sites = 96, 64, 48, 4, 8
    x = torch.randn(T, k, generator=g) * 2 + 3
    w, bias = torch.randn(n, k, generator=g) * k ** -0.5, torch.randn(n, generator=g)
    gamma, beta, pe = torch.rand(k, generator=g) + 0.5, torch.randn(k, generator=g), torch.randn(frames + 2, k, generator=g)
    w_gain, colsum, shift = ops.fold_layernorm(w, gamma, beta, bias=bias, pe=pe)
    assert w_gain.shape == (n, k) and colsum.shape == (n,) and shift.shape == (frames + 2, n)
    mean, var = x.mean(1, keepdim=True), x.var(1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + 1e-5)
    frame_of = (torch.arange(T) // sites) % frames
    got = rstd * (x @ w_gain.t() - mean * colsum[None, :]) + shift[frame_of]
    want = torch.nn.functional.linear(torch.nn.functional.layer_norm(x, (k,), gamma, beta, 1e-5) + pe[frame_of], w, bias)
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4)
    # without a positional-encoding table the shift has one row
    assert ops.fold_layernorm(w, gamma, beta)[2].shape == (1, n)


def test_hoisted_cond_embedding_cache_follows_images_and_weights():
    """MultiControlNetResiduals.cond_embedding (host logic of the hoisted ControlNet conditioning embedding): one evaluation per
    (net, image buffer); re-evaluated when the buffer is written in place, replaced (even at a recycled address), sliced
    differently, or when the embedding weights change; `into` pins the value to a caller-owned buffer (CUDA-graph pointer)."""
    from controlanimate_b200.pipeline import MultiControlNetResiduals

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.controlnet_cond_embedding = torch.nn.Linear(1, 1)
            self.calls = 0

        def embed_condition(self, image):
            self.calls += 1
            return image * float(self.controlnet_cond_embedding.weight.detach().reshape(()))

    net = Net()
    mc = MultiControlNetResiduals([net], [1.0])
    img = torch.arange(8.0).reshape(4, 2).clone()
    a = mc.cond_embedding(0, img)
    assert net.calls == 1 and mc.cond_embedding(0, img) is a and net.calls == 1          # cached
    assert mc.cond_embedding(0, img[0:4]) is a and net.calls == 1                         # a full-range view of the same storage
    img.add_(1.0)                                                                         # next window written in place
    b = mc.cond_embedding(0, img)
    assert net.calls == 2 and torch.equal(b, img * float(net.controlnet_cond_embedding.weight.detach()))
    half = mc.cond_embedding(0, img[2:])                                                  # a CFG half's rows: its own entry
    assert net.calls == 3 and half.shape[0] == 2 and mc.cond_embedding(0, img[2:]) is half and net.calls == 3
    with torch.no_grad():
        net.controlnet_cond_embedding.weight.mul_(2.0)                                    # weights changed in place
    c = mc.cond_embedding(0, img)
    assert net.calls == 4 and torch.equal(c, 2 * b)
    other = img.clone()                                                                   # another buffer with equal contents
    mc.cond_embedding(0, other)
    assert net.calls == 5
    buf = torch.empty_like(c)                                                             # graph-owned destination
    d = mc.cond_embedding(0, img, into=buf)
    assert d is buf and net.calls == 6 and torch.equal(buf, c)
    assert mc.cond_embedding(0, img, into=buf) is buf and net.calls == 6                  # refreshed only when needed
    img.mul_(3.0)
    assert mc.cond_embedding(0, img, into=buf) is buf and net.calls == 7 and torch.equal(buf, img * float(net.controlnet_cond_embedding.weight.detach()))
    # the entry owns the image: a new tensor can not reuse its address while it is cached
    owner = mc._cond_cache[(0, img.data_ptr(), tuple(img.shape), img.dtype)][2]
    assert owner is img
