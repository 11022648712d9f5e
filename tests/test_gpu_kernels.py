"""GPU parity tests: every CUDA kernel, called through the C ABI (ctypes), against the CPU oracle
(oracle/ref_ops.py, itself pinned to the reference's own Python by tests/test_oracle_golden.py).

Tolerances (BASELINE.json north_star): per-op max relative error <= 2e-3 for bf16 storage with fp32
accumulation.  `ref` is the oracle evaluated in fp32 on the SAME (bf16-rounded) inputs and the error is
max(|y - ref| - q(ref)) / max|ref|, where q(ref) = half an ulp of the storage type at |ref| is the
unavoidable quantisation of writing the result as bf16/f16 (bf16 keeps 8 significant bits, so storing
alone costs up to 2^-8 = 3.9e-3 relative; the 2e-3 budget is for the arithmetic).  fp32 storage has
q = 0 and is held to 2e-5, which pins the arithmetic itself.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_ops as R
from oracle import synth

pytestmark = pytest.mark.gpu

REL = {torch.bfloat16: 2e-3, torch.float16: 1e-3, torch.float32: 2e-5}


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from controlanimate_b200 import _lib, ops as O
    lib = _lib.load(build_if_missing=False)
    assert lib.ca_device_sm() == 100, "these kernels are built for sm_100a (B200) only"
    return O


MANT = {torch.bfloat16: 8, torch.float16: 11}


PLAIN = []   # (test id, dtype, relerr, plain max-normalised error, worst per-element relative error): see conftest.pytest_sessionfinish


def relerr(y, ref):
    """The asserted number (docstring above).  Next to it two PLAIN numbers — no half-ulp subtraction — are recorded for every
    call and written to gpurun_out/kernel_relerr.tsv at session end: max|y - ref| / max|ref|, and the worst PER-ELEMENT relative
    error max(|y - ref| / max(|ref|, max|ref| / 64)) (elements within 1/64 of the largest magnitude are held to their own size)."""
    dt = y.dtype
    y = y.detach().float().cpu()
    assert torch.isfinite(y).all()
    err = (y - ref).abs()
    top = ref.abs().max().clamp_min(1e-12)
    PLAIN.append((os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0].split("::", 1)[-1], str(dt).replace("torch.", ""), None,
                  float(err.max() / top), float((err / ref.abs().clamp_min(top / 64)).max())))
    if dt in MANT:  # half an ulp of the storage type at |ref|
        expo = torch.floor(torch.log2(ref.abs().clamp_min(1e-30)))
        err = (err - 0.5 * torch.pow(2.0, expo - (MANT[dt] - 1)) * 1.0001).clamp_min(0)
    out = float(err.max() / ref.abs().max().clamp_min(1e-12))
    PLAIN[-1] = PLAIN[-1][:2] + (out,) + PLAIN[-1][3:]
    return out


def to_layout(x, layout):
    """Return a [b,c,f,h,w] CUDA tensor with the requested memory layout."""
    x = x.cuda()
    if layout == "ncfhw":
        return x.contiguous()
    return x.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)


GN_CASES = [
    # b, c, f, h, w, groups, per_frame, temb
    (2, 64, 3, 6, 5, 32, True, True),      # ragged h*w -> scalar path in NCFHW
    (2, 64, 3, 8, 4, 8, True, False),
    (1, 320, 4, 16, 16, 32, True, True),   # cpg = 10 (vectors straddle groups in BFHWC)
    (2, 96, 2, 8, 8, 32, False, True),     # v1 (non-inflated) statistics over frames too
    (1, 640, 2, 64, 64, 32, True, True),   # 160 KB groups -> multi-chunk domains + domain barrier
    (1, 32, 16, 64, 64, 32, False, False), # one channel per group, long domains (split launch path)
    (2, 1280, 2, 8, 8, 32, True, True),    # tiny groups
    (2, 320, 16, 32, 32, 32, True, True),  # 32 domains x 640 KB: several domains per team (rounds > 1), team barrier
    (1, 960, 2, 64, 64, 32, True, False),  # 7.9 MB domains, cpg = 30: teams of ~80 CTAs
    (3, 2560, 1, 16, 16, 32, True, True),  # one row lane per CTA (c/8 = 320 channel vectors)
    (1, 64, 40, 96, 54, 32, False, True),  # v1 statistics over 40 frames: one 26 MB domain spread over the whole grid
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.float16])
@pytest.mark.parametrize("layout", ["ncfhw", "bfhwc"])
@pytest.mark.parametrize("case", GN_CASES)
def test_groupnorm_silu(ops, case, layout, dtype):
    b, c, f, h, w, groups, per_frame, use_temb = case
    x = (synth.tensor(7, f"gn.{case}", (b, c, f, h, w)) * 1.5 + 0.4).to(dtype)
    gamma = 1 + synth.tensor(7, "gn.gamma", (c,), 0.1)
    beta = synth.tensor(7, "gn.beta", (c,), 0.1)
    temb = synth.tensor(7, "gn.temb", (b, c)) if use_temb else None
    for silu in (True, False):
        ref = R.groupnorm_silu(x.float(), gamma, beta, groups, 1e-5, per_frame, temb, silu)
        y = ops.groupnorm_silu(to_layout(x, layout), gamma.cuda(), beta.cuda(), groups, 1e-5, per_frame=per_frame, silu=silu,
                               temb=None if temb is None else temb.cuda())
        assert y.shape == x.shape
        assert relerr(y, ref) <= REL[dtype], (case, layout, dtype, silu)


@pytest.mark.parametrize("env", [{"CA_GN_SLAB": "0"}, {"CA_GN_SLAB_CLUSTER": "8"}, {"CA_GN_SLAB": "0", "CA_GN_RING": "0"}],
                         ids=["ring", "slab-clusters", "split"])
def test_groupnorm_alternative_paths(env):
    """The native-layout GroupNorm picks among several kernels behind ca_groupnorm_silu (small-domain slab kernel, slice ring,
    split statistics / normalise launches); the path is chosen once per process from the environment, so the ones the
    default run does not reach for a given shape are re-checked in child processes against the same oracle cases."""
    import os
    import subprocess
    import sys
    if os.environ.get("CA_GN_CHILD") == "1":
        pytest.skip("already inside a child run")
    child_env = dict(os.environ, CA_GN_CHILD="1", **env)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", __file__, "-k",
                        "test_groupnorm_silu and bfhwc or test_groupnorm_row_strided"], env=child_env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_groupnorm_row_strided_temb_and_determinism(ops):
    """temb rows may be column slices of one wide [b, sum C] buffer (layers.TembBank); results are bit-reproducible."""
    b, c, f, h, w = 2, 320, 4, 32, 32
    x = to_layout(synth.tensor(9, "gn.ts", (b, c, f, h, w)).bfloat16(), "bfhwc")
    gamma, beta = 1 + synth.tensor(9, "g", (c,), 0.1), synth.tensor(9, "b", (c,), 0.1)
    wide = synth.tensor(9, "wide", (b, 3 * c)).cuda()
    temb = wide[:, c:2 * c]
    assert not temb.is_contiguous()
    ref = R.groupnorm_silu(x.float().cpu(), gamma, beta, 32, 1e-5, True, temb.cpu(), True)
    y0 = ops.groupnorm_silu(x, gamma.cuda(), beta.cuda(), 32, 1e-5, temb=temb)
    assert relerr(y0, ref) <= 2e-3
    for _ in range(3):
        assert torch.equal(ops.groupnorm_silu(x, gamma.cuda(), beta.cuda(), 32, 1e-5, temb=temb), y0)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("n,c,h,w", [(3, 320, 9, 7), (2, 16, 64, 64), (1, 1280, 8, 8)])
def test_bias_act_residual(ops, n, c, h, w, dtype):
    """Convolution epilogue y = (act(x + bias) + residual) * scale against plain fp32 torch on the same inputs."""
    x = synth.tensor(11, f"bar.x{n}{c}", (n, c, h, w)).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    r = synth.tensor(11, f"bar.r{n}{c}", (n, c, h, w)).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    bias = synth.tensor(11, "bar.b", (c,), 0.5)
    xf, rf = x.float().cpu(), r.float().cpu()
    bb = bias.view(1, -1, 1, 1)
    assert relerr(ops.bias_act_residual(x, bias.cuda()), xf + bb) <= REL[dtype]
    assert relerr(ops.bias_act_residual(x, bias.cuda(), silu=True), torch.nn.functional.silu(xf + bb)) <= REL[dtype]
    assert relerr(ops.bias_act_residual(x, bias.cuda(), r, scale=0.5), (xf + bb + rf) * 0.5) <= REL[dtype]
    assert relerr(ops.bias_act_residual(x, None, r), xf + rf) <= REL[dtype]
    y = ops.bias_act_residual(x.clone(), bias.cuda(), r, inplace=True)
    assert relerr(y, xf + bb + rf) <= REL[dtype]
    t = x.permute(0, 2, 3, 1).reshape(-1, c)                       # token-matrix view of the same memory
    assert relerr(ops.bias_act_residual(t, bias.cuda()), (xf + bb).permute(0, 2, 3, 1).reshape(-1, c)) <= REL[dtype]
    with pytest.raises(ValueError):
        ops.bias_act_residual(x.contiguous(), bias.cuda())          # NCHW-contiguous is not channels-last rows
    with pytest.raises(ValueError):
        ops.bias_act_residual(x.cpu(), bias)


def test_groupnorm_in_place_and_errors(ops):
    x = synth.tensor(3, "gn.ip", (1, 64, 2, 8, 8)).cuda().bfloat16()
    g, b = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")
    ref = R.groupnorm_silu(x.float().cpu(), g.cpu(), b.cpu(), 32, 1e-5)
    y = ops.groupnorm_silu(x, g, b, 32, 1e-5, out=x)
    assert y.data_ptr() == x.data_ptr() and relerr(y, ref) <= 2e-3
    with pytest.raises(ValueError):
        ops.groupnorm_silu(x, g, b, 7, 1e-5)                      # c not divisible by groups
    with pytest.raises(ValueError):
        ops.groupnorm_silu(x.cpu(), g, b, 32, 1e-5)               # no CPU fallback
    with pytest.raises(ValueError):
        ops.groupnorm_silu(x[:, :, :, ::2], g, b, 32, 1e-5)       # strided view


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("c,f,d,b", [(320, 16, 12, 2), (64, 5, 7, 1), (1280, 8, 3, 2), (640, 32, 4, 1)])
def test_layernorm_pe(ops, c, f, d, b, dtype):
    x = (synth.tensor(5, f"ln.{c}.{f}", (b * f, d, c)) * 2 + 0.3).to(dtype)
    gamma = 1 + synth.tensor(5, "ln.g", (c,), 0.1)
    beta = synth.tensor(5, "ln.b", (c,), 0.1)
    pe = R.positional_encoding(32, c)
    ref = torch.nn.functional.layer_norm(x.float(), (c,), gamma, beta, 1e-5)
    y = ops.layernorm_pe(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5)
    assert relerr(y, ref) <= REL[dtype]
    # + PE in token order == the reference's "(b f) d c -> (b d) f c ; + pe[:, :f]" (motion_module.py:285-288)
    ref_pe = (ref.reshape(b, f, d, c) + pe[0, :f][None, :, None, :]).reshape(b * f, d, c)
    y = ops.layernorm_pe(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5, pe=pe.cuda(), frames=f, sites=d)
    assert relerr(y, ref_pe) <= REL[dtype]


ATTN_CASES = [
    # b, f, d, heads, hd
    (2, 16, 40, 8, 40),    # SD1.5 level 0 head_dim (k8 tail step)
    (1, 16, 24, 8, 80),
    (1, 16, 9, 8, 160),    # d not a multiple of the site tile
    (2, 8, 33, 8, 40),     # config-1 frame count
    (1, 32, 10, 8, 40),    # PE max_len (two query m-tiles)
    (1, 24, 5, 8, 80),     # v1 max_len
    (2, 5, 7, 8, 8),       # tiny golden-shaped case (f not a multiple of 8)
    (1, 12, 6, 4, 16),
    (1, 1, 3, 2, 8),       # single frame: softmax over one key
    # the remaining (frames, head_dim) compile-time specialisations: 8/24/32 frames x 40/80/160
    (1, 8, 10, 8, 80), (1, 8, 12, 8, 160), (1, 24, 9, 8, 40), (1, 24, 6, 8, 160), (1, 32, 7, 8, 80), (1, 32, 5, 8, 160),
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("case", ATTN_CASES)
def test_temporal_attention_core(ops, case, dtype):
    b, f, d, heads, hd = case
    C = heads * hd
    T = b * f * d
    qkv = synth.tensor(11, f"attn.{case}", (T, 3 * C)).to(dtype)
    q, k, v = (qkv[:, i * C:(i + 1) * C].float() for i in range(3))

    def seq(t):  # token-major [(b f d), C] -> [(b d), f, C]
        return t.reshape(b, f, d, C).permute(0, 2, 1, 3).reshape(b * d, f, C)

    ref = R.attention_core(seq(q), seq(k), seq(v), heads).reshape(b, d, f, C).permute(0, 2, 1, 3).reshape(T, C)
    g = qkv.cuda()
    o = ops.temporal_attention_core(g[:, :C], g[:, C:2 * C], g[:, 2 * C:], batch=b, frames=f, sites=d, heads=heads)
    assert relerr(o, ref) <= REL[dtype], case
    # separate (unpacked) buffers take the same path with ld = C
    o2 = ops.temporal_attention_core(g[:, :C].contiguous(), g[:, C:2 * C].contiguous(), g[:, 2 * C:].contiguous(),
                                     batch=b, frames=f, sites=d, heads=heads)
    assert torch.equal(o, o2)


X_CASES = [
    # frames, sites, heads, head_dim, n_ctx, kv_len, explicit frame->prompt map
    (4, 200, 8, 40, 2, 77, False),    # ragged last query tile (200 = 128 + 72), two prompts x two frames
    (6, 64, 8, 80, 3, 77, True),      # interleaved frame -> prompt map (ControlNet: frame % n_prompts)
    (2, 256, 8, 160, 2, 81, False),   # IP-Adapter length (77 + 4 image tokens): 96-key variant, two P V passes
    (3, 130, 4, 40, 1, 5, False),     # very short prompt, one prompt for every frame
]


@pytest.mark.parametrize("mode", ["ring", "persist"])
def test_layernorm_alternative_paths(mode):
    """ca_layernorm_pe defaults to the flat kernel for 16-bit rows; the smem-ring and the persistent register-resident
    kernels stay selectable (CA_LN_MODE) and are re-checked in child processes against the same oracle cases."""
    import os
    import subprocess
    import sys
    if os.environ.get("CA_LN_CHILD") == "1":
        pytest.skip("already inside a child run")
    child_env = dict(os.environ, CA_LN_CHILD="1", CA_LN_MODE=mode)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", __file__, "-k",
                        "layernorm and not alternative"], env=child_env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("case", X_CASES)
def test_cross_attention_core(ops, case, dtype):
    """Row N2: cross-attention of the spatial transformer against the oracle's attention_core (attention_processor.py:56-62)."""
    frames, d, heads, hd, n_ctx, L, use_map = case
    C = heads * hd
    q = synth.tensor(13, f"xattn.q{case}", (frames * d, C)).to(dtype)
    kv = synth.tensor(13, f"xattn.kv{case}", (n_ctx, L, 2 * C)).to(dtype)
    cmap = (torch.arange(frames) % n_ctx) if use_map else (torch.arange(frames) // (frames // n_ctx))
    kf, vf = kv[:, :, :C].float()[cmap], kv[:, :, C:].float()[cmap]
    ref = R.attention_core(q.float().reshape(frames, d, C), kf, vf, heads).reshape(frames * d, C)
    g = kv.cuda()
    o = ops.cross_attention_core(q.cuda(), g[:, :, :C], g[:, :, C:], frames=frames, sites=d, heads=heads,
                                 ctx_of_frame=cmap.to(torch.int32).cuda() if use_map else None)
    assert relerr(o, ref) <= REL[dtype], case


def test_cross_attention_rejects_unbuilt_variants(ops):
    q = torch.zeros(64, 8 * 64, device="cuda", dtype=torch.bfloat16)
    kv = torch.zeros(1, 77, 8 * 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        ops.cross_attention_core(q, kv, kv, frames=1, sites=64, heads=8)        # head_dim 64 is not an SD1.5 width
    q = torch.zeros(64, 320, device="cuda", dtype=torch.bfloat16)
    kv = torch.zeros(1, 200, 320, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        ops.cross_attention_core(q, kv, kv, frames=1, sites=64, heads=8)        # prompt longer than 96 tokens


def test_temporal_attention_rejects_long_sequences(ops):
    q = torch.zeros(33 * 2, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        ops.temporal_attention_core(q, q, q, batch=1, frames=33, sites=2, heads=8)


def _residual_sets(seed, n_nets, b, f, hw0, dtype, chans=(32, 64, 128, 128)):
    shapes = synth.residual_shapes(chans)
    sets = []
    for k in range(n_nets):
        cur = []
        for i, (c, div) in enumerate(shapes):
            s = max(hw0 // div, 1)
            cur.append(synth.tensor(seed, f"res.{k}.{i}", (b * f, c, s, s + (1 if i % 2 else 0))).to(dtype))
        sets.append(cur)
    return sets


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("n_nets,b_res,b_dst,guess", [(1, 2, 2, False), (2, 2, 2, False), (4, 1, 2, True), (5, 1, 1, False)])
def test_residual_merge_reference_layout(ops, n_nets, b_res, b_dst, guess, dtype):
    """B3 contract in the reference's layouts: (b f) c h w per net -> b c f h w, optionally += into skips."""
    from controlanimate_b200 import _lib as L
    f = 3
    sets = _residual_sets(21, n_nets, b_res, f, 8, dtype)
    cond = [1.0, 0.5, 0.35, 0.4, 0.8][:n_nets]
    level = torch.logspace(-1, 0, 13) if guess else torch.ones(13)
    scales = [[float(level[i]) * cond[k] for i in range(13)] for k in range(n_nets)]
    down_ref, mid_ref = R.merge_controlnet_residuals([[t.float() for t in s] for s in sets], cond, f, guess_mode=guess)
    ref = list(down_ref) + [mid_ref]
    # producer mode (dst = sum)
    dst = [torch.empty((b_res, r.shape[1], f, r.shape[3], r.shape[4]), dtype=dtype, device="cuda") for r in ref]
    ops.residual_merge([[t.cuda() for t in s] for s in sets], scales, dst, frames=f, add_into_dst=False, layout=L.CA_LAYOUT_NCFHW)
    for d, r in zip(dst, ref):
        assert relerr(d, r) <= REL[dtype]
    # in-place mode on skips with batch broadcast (unet.py:567-585)
    skips = [synth.tensor(22, f"skip.{i}", (b_dst, r.shape[1], f, r.shape[3], r.shape[4])).to(dtype) for i, r in enumerate(ref)]
    want = [s.float() + r for s, r in zip(skips, ref)]
    got = [s.cuda() for s in skips]
    ops.residual_merge([[t.cuda() for t in s] for s in sets], scales, got, frames=f, add_into_dst=True, layout=L.CA_LAYOUT_NCFHW)
    for g, wv in zip(got, want):
        assert relerr(g, wv) <= REL[dtype]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_residual_merge_native_layout(ops, dtype):
    from controlanimate_b200 import _lib as L
    f, b = 4, 2
    sets = _residual_sets(23, 2, b, f, 8, dtype)
    scales = [[1.0] * 13, [0.5] * 13]
    ref = [a.float() + 0.5 * c.float() for a, c in zip(*sets)]
    cl = [[t.cuda().contiguous(memory_format=torch.channels_last) for t in s] for s in sets]
    skips = [synth.tensor(24, f"skipn.{i}", tuple(r.shape)).to(dtype) for i, r in enumerate(ref)]
    got = [s.cuda().contiguous(memory_format=torch.channels_last) for s in skips]
    ops.residual_merge(cl, scales, got, frames=f, add_into_dst=True, layout=L.CA_LAYOUT_BFHWC)
    for g, s, r in zip(got, skips, ref):
        assert relerr(g, s.float() + r) <= REL[dtype]


LINEAR_CASES = [
    # m, n, k
    (256, 64, 64), (128, 320, 320), (1000, 960, 320), (4096, 640, 640), (2048, 1280, 2560), (77, 192, 128), (300, 2560, 320),
    # large-m cases: the CTA-pair kernel switches to its B-stationary schedule (K = 320) / many tiles per pair
    (8000, 960, 320), (9001, 2560, 320), (19000, 320, 320), (5000, 320, 1280),
    # round 2 tilings: contiguous tile ranges that straddle n-blocks with a resident B block
    # (bn = 240); 512-column single-stage tiles (two UMMA sub-tiles per A stage) for long K with narrow N
    (3000, 96, 64), (1500, 160, 320), (40000, 1920, 640), (33000, 640, 2560), (70000, 320, 1280), (20000, 1280, 1280),
    # A-stationary schedule (K <= 320, at least one m-block of 256 rows per CTA pair, several n-blocks): the A tile of an
    # m-block stays in shared memory while the pair walks its n-blocks; ragged last m-block, 1..5 resident k-blocks
    (40000, 960, 320), (25000, 2560, 320), (19001, 640, 256), (30000, 192, 64), (21000, 384, 136), (50000, 320, 192),
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,n,k", LINEAR_CASES)
def test_linear_tcgen05(ops, m, n, k, dtype):
    """nn.Linear semantics (motion_module.py:147,155 and the processor's to_q/k/v/to_out) on tcgen05."""
    x = synth.tensor(31, f"lin.x.{m}.{k}", (m, k)).to(dtype)
    w = synth.tensor(31, f"lin.w.{n}.{k}", (n, k), k ** -0.5).to(dtype)
    bias = synth.tensor(31, "lin.b", (n,), 0.1)
    res = synth.tensor(31, f"lin.r.{m}.{n}", (m, n)).to(dtype)
    ref = torch.nn.functional.linear(x.float(), w.float(), bias)
    y = ops.linear(x.cuda(), w.cuda(), bias.cuda())
    assert relerr(y, ref) <= REL[dtype]
    y = ops.linear(x.cuda(), w.cuda(), None, residual=res.cuda())
    assert relerr(y, torch.nn.functional.linear(x.float(), w.float()) + res.float()) <= REL[dtype]
    if n % 64 == 0:
        a, g = ref.chunk(2, dim=-1)
        y = ops.linear(x.cuda(), w.cuda(), bias.cuda(), geglu=True)
        assert y.shape == (m, n // 2)
        assert relerr(y, a * torch.nn.functional.gelu(g)) <= REL[dtype]


LN_FOLD_CASES = [
    # T rows, k (LayerNorm width), n, frames, sites   (frames > 1: per-frame positional-encoding shift rows)
    (4096, 320, 960, 16, 128), (2048, 320, 960, 5, 32), (1000, 320, 320, 1, 1), (6144, 640, 1920, 16, 64),
    (2048, 1280, 3840, 16, 64), (3000, 64, 192, 1, 1), (20000, 320, 2560, 1, 1), (70000, 320, 960, 16, 4096),
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("T,k,n,frames,sites", LN_FOLD_CASES)
def test_layernorm_folded_into_projection(ops, T, k, n, frames, sites, dtype):
    """ca_row_stats + ca_linear_ln == Linear(LayerNorm(x) + pe[frame]) (motion_module.py:214-215, 285-288 -> :321's to_q/k/v;
    attention.py:271-297's norm1-3 -> projections), also in front of GEGLU, against the fp32 formula on the same bf16 rows.
    Rows get a large common offset so that the epilogue's mean * colsum term really cancels something."""
    T = T // (frames * sites) * frames * sites if frames > 1 else T
    x = (synth.tensor(41, f"lnf.x.{T}.{k}", (T, k)) * 1.5 + synth.tensor(41, f"lnf.o.{T}", (T, 1)) * 4.0).to(dtype)
    w = synth.tensor(41, f"lnf.w.{n}.{k}", (n, k), k ** -0.5).to(dtype)
    gamma = 1.0 + synth.tensor(41, "lnf.g", (k,), 0.3)
    beta = synth.tensor(41, "lnf.b", (k,), 0.2)
    bias = synth.tensor(41, "lnf.bias", (n,), 0.1)
    pe = synth.tensor(41, "lnf.pe", (24, k), 0.5) if frames > 1 else None
    ln = torch.nn.functional.layer_norm(x.float(), (k,), gamma, beta, 1e-5)
    if pe is not None:
        frame_of = (torch.arange(T) // sites) % frames
        ln = ln + pe[frame_of]
    ref = torch.nn.functional.linear(ln, w.float(), bias)
    xc = x.cuda()
    st = ops.row_stats(xc, 1e-5)
    mean, var = x.float().mean(1), x.float().var(1, unbiased=False)
    assert torch.allclose(st[:, 0].cpu(), mean, atol=1e-5, rtol=1e-5)
    assert torch.allclose(st[:, 1].cpu(), torch.rsqrt(var + 1e-5), atol=0, rtol=1e-4)
    w_gain, colsum, shift = ops.fold_layernorm(w.cuda(), gamma.cuda(), beta.cuda(), bias=bias.cuda(), pe=None if pe is None else pe.cuda())
    assert shift.shape == (24 if pe is not None else 1, n)
    y = ops.linear_ln(xc, st, w_gain, colsum, shift, frames=frames, sites=sites)
    assert relerr(y, ref) <= 2 * REL[dtype]           # two roundings on the weights' side (w, then w * gamma)
    if n % 64 == 0:
        a, g = ref.chunk(2, dim=-1)
        y = ops.linear_ln(xc, st, w_gain, colsum, shift, frames=frames, sites=sites, geglu=True)
        assert y.shape == (T, n // 2)
        assert relerr(y, a * torch.nn.functional.gelu(g)) <= 2 * REL[dtype]
    # a strided view of a wider buffer (row stride > k) gives the same statistics
    wide = torch.zeros((min(T, 512), k + 64), dtype=dtype, device="cuda")
    wide[:, :k] = xc[:wide.shape[0]]
    assert torch.equal(ops.row_stats(wide[:, :k], 1e-5), st[:wide.shape[0]])


def test_layernorm_fold_contract_errors(ops):
    x = torch.zeros((64, 64), dtype=torch.bfloat16, device="cuda")
    st = ops.row_stats(x)
    w_gain, colsum, shift = ops.fold_layernorm(torch.zeros((64, 64), dtype=torch.bfloat16, device="cuda"), torch.ones(64, device="cuda"),
                                               torch.zeros(64, device="cuda"), pe=torch.zeros((8, 64), device="cuda"))
    with pytest.raises((RuntimeError, ValueError)):
        ops.linear_ln(x, st, w_gain, colsum, shift, frames=16, sites=32)      # 16 frames but an 8-row table
    with pytest.raises((RuntimeError, ValueError)):
        ops.linear_ln(x, st, w_gain, colsum, shift, frames=2, sites=24)       # sites not a multiple of 32
    with pytest.raises(ValueError):
        ops.linear_ln(x, st[:32], w_gain, colsum, shift)
    with pytest.raises(ValueError):
        ops.row_stats(x.float())


def test_linear_strided_input_and_repeat(ops):
    """x may be a column slice of a wider buffer (row stride > k); repeated launches reuse barriers/TMEM cleanly."""
    buf = synth.tensor(32, "lin.buf", (512, 3 * 128)).bfloat16().cuda()
    w = synth.tensor(32, "lin.w2", (256, 128), 128 ** -0.5).bfloat16().cuda()
    for i in range(3):
        xs = buf[:, i * 128:(i + 1) * 128]
        y = ops.linear(xs, w)
        ref = torch.nn.functional.linear(xs.float().cpu(), w.float().cpu())
        assert relerr(y, ref) <= 2e-3


FUSED_CASES = [
    # c, f, d, b   (heads = 8)
    (320, 16, 32, 2), (320, 16, 24, 1),      # config-2 width: hd 40, S = 8; d = 24 leaves the pair's second CTA partly empty
    (320, 8, 40, 1), (320, 32, 12, 2),       # S = 16 / S = 4
    (320, 24, 13, 1),                        # S = 5: 120 of 128 tile rows, ragged site tiles
    (64, 16, 20, 2), (128, 5, 9, 1),         # narrow variants (hd 8 / 16), odd frame count
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("c,f,d,b", FUSED_CASES)
def test_temporal_attention_fused(ops, c, f, d, b, dtype):
    """ca_temporal_attn_fused == x + to_out(attn(LN(x) + pe)): LayerNorm (motion_module.py:214) + VersatileAttention.forward
    (:272-329) + AttentionProcessor (attention_processor.py:186-272) + residual (:219), against the oracle."""
    from oracle import ref_ops as R
    heads = 8
    T = b * f * d
    x = synth.tensor(41, f"fu.x.{c}.{f}.{d}", (T, c)).to(dtype)
    w = {n: synth.tensor(41, f"fu.{n}.{c}", (c, c), c ** -0.5).to(dtype) for n in ("wq", "wk", "wv", "wo")}
    bo = synth.tensor(41, "fu.bo", (c,), 0.1)
    gamma = 1 + synth.tensor(41, "fu.g", (c,), 0.1)
    beta = synth.tensor(41, "fu.b", (c,), 0.1)
    pe = R.positional_encoding(32, c)
    # oracle: rows are (b f) d c
    xf = x.float().reshape(b * f, d, c)
    n = torch.nn.functional.layer_norm(xf, (c,), gamma, beta, 1e-5)
    ref = R.versatile_attention(n, f, pe, w["wq"].float(), w["wk"].float(), w["wv"].float(), w["wo"].float(), bo, heads) + xf
    perm = ops.pack_qkv_per_head(w["wq"].cuda(), w["wk"].cuda(), w["wv"].cuda(), heads)
    y = ops.temporal_attention_fused(x.cuda(), gamma.cuda(), beta.cuda(), pe.cuda(), perm, w["wo"].cuda(), bo.cuda(), batch=b, frames=f,
                                     sites=d, heads=heads)
    # four chained roundings to the 16-bit storage type (LN output, q|k|v, attention output, result): 2x the per-op bound
    assert relerr(y, ref.reshape(T, c)) <= 2 * REL[dtype], relerr(y, ref.reshape(T, c))
    # in place, and deterministic
    x2 = x.cuda().clone()
    y2 = ops.temporal_attention_fused(x2, gamma.cuda(), beta.cuda(), pe.cuda(), perm, w["wo"].cuda(), bo.cuda(), batch=b, frames=f,
                                      sites=d, heads=heads, out=x2)
    assert torch.equal(y2, y)


def test_temporal_attention_fused_rejects_unbuilt_width(ops):
    x = torch.zeros(16 * 4, 640, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(640, 640, device="cuda", dtype=torch.bfloat16)
    v = torch.zeros(640, device="cuda")
    with pytest.raises(ValueError):
        ops.temporal_attention_fused(x, v, v, None, ops.pack_qkv_per_head(w, w, w, 8), w, v, batch=1, frames=16, sites=4, heads=8)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,c,h,w,size", [(4, 64, 8, 8, None), (3, 320, 7, 12, None), (2, 1280, 7, 4, (14, 7)), (2, 640, 14, 7, (27, 14)),
                                          (1, 8, 5, 3, (9, 7)), (32, 640, 32, 32, None)])
def test_upsample_nearest_is_torch_interpolate(ops, n, c, h, w, size, dtype):
    """Upsample3D.forward's F.interpolate(mode="nearest") (resnet.py:63-69), by 2 and to the explicit sizes of an odd pyramid
    (forward_upsample_size: config 4's 54 -> 27 -> 14 -> 7): a copy, so bit-exact against torch."""
    x = synth.tensor(51, f"up.{n}.{c}.{h}.{w}", (n, c, h, w)).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    y = ops.upsample_nearest(x, size)
    ref = (torch.nn.functional.interpolate(x, size=size, mode="nearest") if size is not None
           else torch.nn.functional.interpolate(x, scale_factor=2.0, mode="nearest"))
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(y, ref)
    with pytest.raises(ValueError):
        ops.upsample_nearest(x.contiguous(), size)          # NCHW-contiguous: not the native layout


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,ca,cb,h,w", [(4, 64, 64, 8, 8), (3, 320, 640, 7, 12), (2, 1280, 1280, 7, 4), (1, 8, 24, 5, 3), (32, 320, 320, 64, 64)])
def test_concat_channels_is_torch_cat(ops, n, ca, cb, h, w, dtype):
    """The skip concat in front of every up-block resnet (unet_blocks.py:636, :742): bit-exact against torch.cat."""
    a = synth.tensor(52, f"cat.a.{n}.{ca}.{h}", (n, ca, h, w)).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    b = synth.tensor(52, f"cat.b.{n}.{cb}.{h}", (n, cb, h, w)).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    y = ops.concat_channels(a, b)
    assert y.is_contiguous(memory_format=torch.channels_last) and torch.equal(y, torch.cat([a, b], dim=1))
    with pytest.raises(ValueError):
        ops.concat_channels(a, b[:, :, :-1])


SPATIAL_ATTN_CASES = [
    # frames, sites, heads, head_dim, input scale
    (2, 256, 2, 40, 1.0), (1, 300, 3, 40, 1.0), (2, 128, 1, 64, 1.0), (1, 1000, 2, 32, 2.5), (1, 77, 2, 40, 1.0),
    (2, 1024, 8, 40, 1.5), (1, 648, 4, 40, 3.0), (1, 4096, 2, 40, 1.0),
    (1, 5184, 1, 40, 1.0), (1, 1296, 2, 40, 2.0), (3, 130, 2, 16, 1.0),      # config 4's 96 x 54 latents (and their 48 x 27 level)
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("frames,sites,heads,hd,mag", SPATIAL_ATTN_CASES)
def test_spatial_attention_core(ops, frames, sites, heads, hd, mag, dtype):
    """ca_spatial_attn_core == softmax(q k^T / sqrt(hd)) v per (frame, head) over the sites (attention.py:268-271 through the
    processor's attention arithmetic, attention_processor.py:56-62), q / k / v read as column slices of one packed [T, 3C]
    buffer; ragged site counts (query and key tails), score ranges wide enough that the lazily raised maximum is exercised."""
    c = heads * hd
    T = frames * sites
    qkv = (synth.tensor(61, f"sa.{frames}.{sites}.{heads}.{hd}", (T, 3 * c)) * mag).to(dtype)
    q, k, v = (qkv[:, i * c:(i + 1) * c].float().reshape(frames, sites, heads, hd).transpose(1, 2) for i in range(3))
    att = torch.softmax(q @ k.transpose(-1, -2) * hd ** -0.5, dim=-1)
    ref = (att @ v).transpose(1, 2).reshape(T, c)
    g = qkv.cuda()
    o = ops.spatial_attention_core(g[:, :c], g[:, c:2 * c], g[:, 2 * c:], frames=frames, sites=sites, heads=heads)
    assert relerr(o, ref) <= 2 * REL[dtype]          # P is rounded to the storage type before P V, like every flash kernel


@pytest.mark.parametrize("lat_dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("model_dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape,cfg,t", [((1, 4, 16, 64, 64), True, 951), ((1, 4, 3, 7, 5), True, 51), ((1, 4, 8, 12, 7), False, 501),
                                         ((1, 4, 2, 6, 6), True, 1)])
def test_cfg_ddim_step(ops, shape, cfg, t, model_dtype, lat_dtype):
    """ca_cfg_ddim_step == `.to(latents_dtype)` + `u + g (c - u)` (controlanimation_pipeline.py:841, 845-846) + the oracle's
    eta-0 DDIM update (:849; diffusers DDIMScheduler, third party -> parity unpinned) on the same rounded inputs; ragged
    element counts, the last step (prev alpha = 1), in place."""
    from controlanimate_b200.pipeline import DDIMScheduler
    g = 7.5
    rows = 2 if cfg else 1
    out = synth.tensor(71, f"dd.o.{shape}.{t}", (rows,) + shape[1:]).to(model_dtype)
    lat = synth.tensor(71, f"dd.x.{shape}.{t}", shape).to(lat_dtype)
    eps = out.to(lat_dtype).float()
    if cfg:
        eps = eps[:1] + g * (eps[1:] - eps[:1])
    ref = R.ddim_step(eps, t, lat.float(), R.ddim_alphas_cumprod(), 20)
    sched = DDIMScheduler()
    sched.set_timesteps(20)
    noise = torch.empty_like(lat, device="cuda")
    x = lat.cuda()
    y = ops.cfg_ddim_step(out.cuda(), x, g if cfg else None, sched.coefficients(t), noise_out=noise)
    assert y.dtype == lat_dtype and y.shape == lat.shape and torch.equal(x.cpu(), lat)     # input untouched
    assert relerr(y, ref) <= REL[lat_dtype]
    assert relerr(noise, eps) <= REL[lat_dtype]
    plain = float((y.float().cpu() - ref).abs().max() / ref.abs().max())
    print(f"cfg_ddim_step {shape} cfg={cfg} t={t} {model_dtype}->{lat_dtype}: plain relative error {plain:.2e}")
    z = ops.cfg_ddim_step(out.cuda(), x, g if cfg else None, sched.coefficients(t), out=x)  # in place
    assert z.data_ptr() == x.data_ptr() and torch.equal(z, y)
    with pytest.raises(ValueError):
        ops.cfg_ddim_step(out.cuda()[:1], x, g, sched.coefficients(t))                      # CFG needs two rows
    with pytest.raises(ValueError):
        ops.cfg_ddim_step(out.cuda(), x, g if cfg else None, sched.coefficients(t), noise_out=x)   # a second output, not an alias
    with pytest.raises(ValueError):
        ops.cfg_ddim_step(out.cuda(), x.transpose(3, 4), g if cfg else None, sched.coefficients(t))
