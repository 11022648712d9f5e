"""GPU property tests at BASELINE.json's FULL sizes (config 2: b=2 CFG halves, 16 frames, latent 64x64, C=320/640/1280).

The CPU oracle needs minutes at these sizes, so the kernels are held to size-independent properties of the operators
instead (the small-size oracle parity lives in test_gpu_kernels.py / test_gpu_model.py):

* GroupNorm / LayerNorm without affine: every statistics domain of the output has mean 0 and variance 1; the affine
  part, the time-embedding shift and SiLU are then pinned by linearity / invariance identities.
* attention (temporal and cross): constant V rows pass through unchanged (softmax rows sum to one), K = 0 gives the
  plain average of V, permuting the keys together with the values leaves the output unchanged.
* residual merge: linear in the scales, additive over the nets, and a permutation of the reference's rearrange.
"""
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu

B, F, LAT = 2, 16, 64
LEVELS = [(320, 64), (640, 32), (1280, 16), (1280, 8)]


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from controlanimate_b200 import _lib, ops as O
    assert _lib.load(build_if_missing=False).ca_device_sm() == 100
    return O


def native(x):  # [b,c,f,h,w] logical, BFHWC memory
    return x.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)


@pytest.mark.parametrize("c,s", LEVELS)
@pytest.mark.parametrize("layout", ["native", "ncfhw"])
def test_groupnorm_full_size_properties(ops, c, s, layout):
    g = torch.Generator(device="cuda").manual_seed(c + s)
    x = (torch.randn(B, c, F, s, s, device="cuda", generator=g) * 1.7 + 0.6).bfloat16()
    xin = native(x) if layout == "native" else x.contiguous()
    ones, zeros = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    y = ops.groupnorm_silu(xin, ones, zeros, 32, 1e-5, per_frame=True, silu=False).float()
    # every (batch, frame, group) domain of the normalised tensor: mean 0, variance 1 (bf16 storage of y: ~2^-9 relative)
    d = y.reshape(B, 32, c // 32, F, s * s).permute(0, 3, 1, 2, 4).reshape(B * F * 32, -1)
    assert float(d.mean(1).abs().max()) < 4e-3
    assert float((d.var(1, unbiased=False) - 1).abs().max()) < 1e-2
    # affine: GN(x; gamma, beta) = gamma * GN(x; 1, 0) + beta
    gamma = 1 + 0.2 * torch.randn(c, device="cuda", generator=g)
    beta = 0.3 * torch.randn(c, device="cuda", generator=g)
    ya = ops.groupnorm_silu(xin, gamma, beta, 32, 1e-5, per_frame=True, silu=False).float()
    ref = y * gamma.view(1, -1, 1, 1, 1) + beta.view(1, -1, 1, 1, 1)
    assert float((ya - ref).abs().max()) < 0.06                      # two bf16 roundings of O(4) values
    # SiLU is applied pointwise on top
    ys = ops.groupnorm_silu(xin, gamma, beta, 32, 1e-5, per_frame=True, silu=True).float()
    assert float((ys - torch.nn.functional.silu(ya)).abs().max()) < 0.06
    # a per-(batch, channel) shift that is constant inside every group leaves the normalised tensor unchanged
    shift = torch.randn(B, 32, device="cuda", generator=g).repeat_interleave(c // 32, dim=1).contiguous()
    yt = ops.groupnorm_silu(xin, ones, zeros, 32, 1e-5, per_frame=True, silu=False, temb=shift).float()
    assert float((yt - y).abs().max()) < 0.06
    # bit-reproducible
    assert torch.equal(ops.groupnorm_silu(xin, gamma, beta, 32, 1e-5, per_frame=True, silu=True).float(), ys)


@pytest.mark.parametrize("c,s", LEVELS)
def test_layernorm_full_size_properties(ops, c, s):
    g = torch.Generator(device="cuda").manual_seed(7 * c + s)
    T = B * F * s * s
    x = (torch.randn(T, c, device="cuda", generator=g) * 2.0 - 0.5).bfloat16()
    ones, zeros = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    y = ops.layernorm_pe(x, ones, zeros, 1e-5).float()
    assert float(y.mean(1).abs().max()) < 4e-3
    assert float((y.var(1, unbiased=False) - 1).abs().max()) < 1.5e-2
    # the positional encoding is added per frame AFTER the affine: rows of frame k differ from the plain result by pe[k]
    pe = torch.randn(32, c, device="cuda", generator=g)
    yp = ops.layernorm_pe(x, ones, zeros, 1e-5, pe=pe, frames=F, sites=s * s).float()
    frame = (torch.arange(T, device="cuda") // (s * s)) % F
    assert float((yp - (y + pe[frame])).abs().max()) < 0.06


@pytest.mark.parametrize("c,s", LEVELS)
def test_temporal_attention_full_size_properties(ops, c, s):
    g = torch.Generator(device="cuda").manual_seed(11 * c + s)
    d = s * s
    T = B * F * d
    q = torch.randn(T, c, device="cuda", generator=g).bfloat16()
    k = torch.randn(T, c, device="cuda", generator=g).bfloat16()
    # V constant along the frame axis at every site -> the output is that constant (softmax rows sum to one)
    v_site = torch.randn(B, 1, d, c, device="cuda", generator=g).bfloat16()
    v = v_site.expand(B, F, d, c).reshape(T, c).contiguous()
    o = ops.temporal_attention_core(q, k, v, batch=B, frames=F, sites=d, heads=8)
    assert float((o.float() - v.float()).abs().max()) < 0.03
    # K = 0 -> uniform weights: the mean of V over the frames of the same site
    v2 = torch.randn(T, c, device="cuda", generator=g).bfloat16()
    o2 = ops.temporal_attention_core(q, torch.zeros_like(k), v2, batch=B, frames=F, sites=d, heads=8).float()
    mean = v2.float().reshape(B, F, d, c).mean(1, keepdim=True).expand(B, F, d, c).reshape(T, c)
    assert float((o2 - mean).abs().max()) < 0.03
    # permuting keys and values together along the frame axis leaves the output unchanged
    perm = torch.randperm(F, device="cuda", generator=g)
    kp = k.reshape(B, F, d, c)[:, perm].reshape(T, c).contiguous()
    vp = v2.reshape(B, F, d, c)[:, perm].reshape(T, c).contiguous()
    o3 = ops.temporal_attention_core(q, k, v2, batch=B, frames=F, sites=d, heads=8).float()
    o4 = ops.temporal_attention_core(q, kp, vp, batch=B, frames=F, sites=d, heads=8).float()
    assert float((o3 - o4).abs().max()) < 0.03


@pytest.mark.parametrize("c,s", LEVELS)
def test_cross_attention_full_size_properties(ops, c, s):
    g = torch.Generator(device="cuda").manual_seed(13 * c + s)
    d, frames, L = s * s, B * F, 77
    q = torch.randn(frames * d, c, device="cuda", generator=g).bfloat16()
    kv = torch.randn(B, L, 2 * c, device="cuda", generator=g).bfloat16()
    # K = 0 -> every site gets the mean of its prompt's V rows
    kz = torch.zeros(B, L, c, device="cuda", dtype=torch.bfloat16)
    o = ops.cross_attention_core(q, kz, kv[:, :, c:], frames=frames, sites=d, heads=8).float()
    mean = kv[:, :, c:].float().mean(1)                                  # [B, c]
    ref = mean.repeat_interleave(F, dim=0).unsqueeze(1).expand(frames, d, c).reshape(frames * d, c)
    assert float((o - ref).abs().max()) < 0.03
    # permuting the prompt tokens (keys with values) leaves the output unchanged; the explicit frame -> prompt map agrees
    perm = torch.randperm(L, device="cuda", generator=g)
    o1 = ops.cross_attention_core(q, kv[:, :, :c], kv[:, :, c:], frames=frames, sites=d, heads=8).float()
    kvp = kv[:, perm].contiguous()
    o2 = ops.cross_attention_core(q, kvp[:, :, :c], kvp[:, :, c:], frames=frames, sites=d, heads=8).float()
    assert float((o1 - o2).abs().max()) < 0.03
    cmap = (torch.arange(frames, device="cuda") // F).to(torch.int32)
    o3 = ops.cross_attention_core(q, kv[:, :, :c], kv[:, :, c:], frames=frames, sites=d, heads=8, ctx_of_frame=cmap).float()
    assert torch.equal(o1, o3)


def test_residual_merge_full_size_properties(ops):
    from controlanimate_b200 import _lib as L
    shapes = synth.residual_shapes((320, 640, 1280, 1280))
    g = torch.Generator(device="cuda").manual_seed(5)

    def make():
        return [torch.randn(B * F, ch, LAT // div, LAT // div, device="cuda", generator=g).bfloat16() for ch, div in shapes]

    a, b2 = make(), make()

    def merge(nets, scales):
        dst = [torch.empty(B, t.shape[1], F, t.shape[2], t.shape[3], device="cuda", dtype=torch.bfloat16) for t in nets[0]]
        ops.residual_merge(nets, [[sc] * len(dst) for sc in scales], dst, frames=F, add_into_dst=False, layout=L.CA_LAYOUT_NCFHW)
        return dst

    one = merge([a], [1.0])
    for t, m in zip(a, one):   # scale 1, one net: exactly the reference's '(b f) c h w -> b c f h w' rearrange
        assert torch.equal(m, t.reshape(B, F, *t.shape[1:]).permute(0, 2, 1, 3, 4))
    half = merge([a], [0.5])
    for m1, mh in zip(one, half):   # linear in the scale (0.5 is exact in bf16)
        assert torch.equal(mh.float(), m1.float() * 0.5)
    both = merge([a, b2], [1.0, 0.5])
    only_b = merge([b2], [0.5])
    for mab, ma, mb in zip(both, one, only_b):   # additive over the nets (fp32 accumulate, one rounding)
        assert float((mab.float() - (ma.float() + mb.float())).abs().max()) <= 0.04
