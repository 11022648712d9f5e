"""CPU oracle: plain-torch fp32 restatement of the ControlAnimate denoising hot path, op by op.

TEST INFRASTRUCTURE.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; nothing under `controlanimate_b200/` does.

Every function cites the reference lines (relative to /root/reference) it restates.  Parity
status: the functions that restate code living IN the reference (GroupNorm+SiLU, ResnetBlock3D,
motion module, attention processor, UNet residual add, whole UNet3D in ref_unet3d.py) are PINNED
against outputs of the reference's own Python, imported unmodified through `oracle/diffusers_shim`
by `oracle/gen_golden.py` (fixtures: tests/golden/*.npz, checked in tests/test_oracle_golden.py).
The ControlNet scale-and-sum (`controlnet_scale_and_sum`) restates diffusers 0.23.0
(`ControlNetModel.forward` tail / `MultiControlNetModel.forward`), which is a third-party
dependency absent from /root/reference and from this image: that one function is PARITY UNPINNED
 and anchored only on the reference's call site
`modules/controlresiduals_pipeline.py:294-316`.

All functions are layout-explicit: tensors are [b, c, f, h, w] ("ncfhw") exactly as the reference
passes them unless stated otherwise, and all arithmetic is fp32 (or the dtype given).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# When True, attention_core dispatches to F.scaled_dot_product_attention — what the reference's own
# `AttnProcessor2_0` does (modules/attention_processor.py:247-256).  Only the GPU-eager yardstick of bench.py flips it
# (the explicit softmax below would materialise a 4096 x 4096 score matrix per head and frame); parity tests keep False.
USE_SDPA = False


# ----------------------------------------------------------------------------------------------
# kernel (2): GroupNorm + SiLU (+ time-embedding add)
# ----------------------------------------------------------------------------------------------
def groupnorm(x: Tensor, gamma: Tensor, beta: Tensor, groups: int, eps: float, per_frame: bool = True) -> Tensor:
    """GroupNorm over a video tensor x[b,c,f,h,w].

    per_frame=True  : `InflatedGroupNorm` (resnet.py:23-31) — frames folded into batch, so the
                      statistics of group g are taken over (c/groups, h, w) per (b, f).
    per_frame=False : plain `nn.GroupNorm` on the 5-D tensor (resnet.py:150-151, v1 configs) —
                      statistics over (c/groups, f, h, w) per b.
    Biased variance, y = (x-mean)/sqrt(var+eps)*gamma_c + beta_c (torch semantics).
    """
    b, c, f, h, w = x.shape
    cpg = c // groups
    if per_frame:
        xg = x.reshape(b, groups, cpg, f, h * w).permute(0, 3, 1, 2, 4).reshape(b, f, groups, cpg * h * w)
        mean = xg.mean(-1)
        var = ((xg - mean[..., None]) ** 2).mean(-1)
        mean = mean.permute(0, 2, 1)[:, :, None, :, None, None]  # b g 1 f 1 1
        var = var.permute(0, 2, 1)[:, :, None, :, None, None]
    else:
        xg = x.reshape(b, groups, cpg * f * h * w)
        mean = xg.mean(-1)
        var = ((xg - mean[..., None]) ** 2).mean(-1)
        mean = mean[:, :, None, None, None, None]
        var = var[:, :, None, None, None, None]
    y = (x.reshape(b, groups, cpg, f, h, w) - mean) * torch.rsqrt(var + eps)
    y = y.reshape(b, c, f, h, w)
    return y * gamma[None, :, None, None, None] + beta[None, :, None, None, None]


def groupnorm_silu(x: Tensor, gamma: Tensor, beta: Tensor, groups: int = 32, eps: float = 1e-5,
                   per_frame: bool = True, temb: Optional[Tensor] = None, silu: bool = True) -> Tensor:
    """y = SiLU(GroupNorm(x + temb[:, :, None, None, None])).

    Restates `ResnetBlock3D.forward` resnet.py:191-192 (norm1+nonlinearity), :199-208
    (`hidden_states + temb` then norm2 + nonlinearity) and unet.py:614-615 (conv_norm_out+conv_act).
    `temb` is the already-projected [b, c] embedding (resnet.py:196-197).
    """
    if temb is not None:
        x = x + temb[:, :, None, None, None]
    y = groupnorm(x, gamma, beta, groups, eps, per_frame)
    return y * torch.sigmoid(y) if silu else y


# ----------------------------------------------------------------------------------------------
# kernel (1): temporal attention
# ----------------------------------------------------------------------------------------------
def positional_encoding(max_len: int, dim: int) -> Tensor:
    """Sinusoidal buffer pe[1, max_len, dim] — `PositionalEncoding.__init__` motion_module.py:236-244."""
    pos = torch.arange(max_len, dtype=torch.float32)[:, None]
    div = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * (-math.log(10000.0) / dim))
    pe = torch.zeros(1, max_len, dim)
    pe[0, :, 0::2] = torch.sin(pos * div)
    pe[0, :, 1::2] = torch.cos(pos * div)
    return pe


def attention_core(q: Tensor, k: Tensor, v: Tensor, heads: int, scale: Optional[float] = None) -> Tensor:
    """softmax(q kᵀ · scale) v per head; q,k,v [B, S, heads*hd] -> [B, S, heads*hd].

    Restates the math path modules/attention_processor.py:56-62 (head split, baddbmm·scale,
    softmax over keys, bmm, head merge); identical to the SDPA path :247-256 in exact arithmetic.
    """
    B, S, C = q.shape
    hd = C // heads
    scale = hd ** -0.5 if scale is None else scale
    qh = q.reshape(B, S, heads, hd).transpose(1, 2)
    kh = k.reshape(B, k.shape[1], heads, hd).transpose(1, 2)
    vh = v.reshape(B, v.shape[1], heads, hd).transpose(1, 2)
    if USE_SDPA:
        o = F.scaled_dot_product_attention(qh, kh, vh, scale=scale)
    else:
        p = torch.softmax(torch.matmul(qh, kh.transpose(-1, -2)) * scale, dim=-1)
        o = torch.matmul(p, vh)
    return o.transpose(1, 2).reshape(B, S, C)


def attention_processor(x: Tensor, wq: Tensor, wk: Tensor, wv: Tensor, wo: Tensor, bo: Tensor, heads: int,
                        context: Optional[Tensor] = None) -> Tensor:
    """The AttentionProcessor contract (B1): to_out(attn(to_q(x), to_k(ctx), to_v(ctx))).

    modules/attention_processor.py:46-66 / :232-262 with the inert branches removed
    (no spatial_norm/group_norm/norm_cross, residual_connection False, rescale_output_factor 1).
    """
    ctx = x if context is None else context
    q = F.linear(x, wq)
    k = F.linear(ctx, wk)
    v = F.linear(ctx, wv)
    return F.linear(attention_core(q, k, v, heads), wo, bo)


def ip_attention_processor(x: Tensor, ctx: Tensor, wq: Tensor, wk: Tensor, wv: Tensor, wo: Tensor, bo: Tensor, wk_ip: Tensor,
                           wv_ip: Tensor, heads: int, num_tokens: int = 4, scale: float = 1.0) -> Tensor:
    """`IPAttnProcessor2_0.__call__` modules/attention_processor.py:367-492 (IP-Adapter): the last `num_tokens` context
    rows are image tokens with their own K/V projections; the two attention outputs are summed with `scale` (:476)."""
    end = ctx.shape[1] - num_tokens                                                                  # :433-437
    text, img = ctx[:, :end], ctx[:, end:]
    q = F.linear(x, wq)
    o = attention_core(q, F.linear(text, wk), F.linear(text, wv), heads)                             # :442-459
    o_ip = attention_core(q, F.linear(img, wk_ip), F.linear(img, wv_ip), heads)                      # :462-474
    return F.linear(o + scale * o_ip, wo, bo)                                                        # :476-481


def versatile_attention(h: Tensor, video_length: int, pe: Optional[Tensor], wq, wk, wv, wo, bo, heads: int) -> Tensor:
    """`VersatileAttention.forward` motion_module.py:272-329 for attention_mode="Temporal", self-attn.

    h [(b f), d, c] -> "(b d) f c" (:285) -> + pe[:, :f] (:287-288) -> processor (:321) -> back (:327).
    The duplicate Q/K/V projections at :299-311 are dead code (results unused) and are omitted.
    """
    bf, d, c = h.shape
    f = video_length
    b = bf // f
    x = h.reshape(b, f, d, c).permute(0, 2, 1, 3).reshape(b * d, f, c)
    if pe is not None:
        x = x + pe[:, :f].to(x.dtype)
    o = attention_processor(x, wq, wk, wv, wo, bo, heads)
    return o.reshape(b, d, f, c).permute(0, 2, 1, 3).reshape(bf, d, c)


def geglu_feedforward(x: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor) -> Tensor:
    """diffusers FeedForward(activation_fn="geglu"): Linear(c->8c); a,g = chunk; a*gelu_erf(g); Linear(4c->c).

    Same arithmetic as the reference's local copy attention.py:303-357 (GEGLU = diffusers
    activations.GEGLU, exact-erf GELU).
    """
    a, g = F.linear(x, w1, b1).chunk(2, dim=-1)
    return F.linear(a * F.gelu(g), w2, b2)


def motion_module(x: Tensor, sd: Dict[str, Tensor], prefix: str = "", heads: int = 8, groups: int = 32) -> Tensor:
    """`VanillaTemporalModule.forward` (motion_module.py:79-84) -> `TemporalTransformer3DModel.forward`
    (:136-160) -> `TemporalTransformerBlock.forward` (:212-224), one transformer block, two
    Temporal_Self attention blocks (inference-v2.yaml:14-22).  x [b,c,f,h,w] -> [b,c,f,h,w].

    `sd` uses the reference state_dict keys below `prefix`.
    """
    p = prefix + "temporal_transformer."
    b, c, f, hh, ww = x.shape
    g = groupnorm(x, sd[p + "norm.weight"], sd[p + "norm.bias"], groups, 1e-6, per_frame=True)      # :139-144
    h = g.permute(0, 2, 3, 4, 1).reshape(b * f, hh * ww, c)                                         # :146
    h = F.linear(h, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])                               # :147
    t = p + "transformer_blocks.0."
    for i in (0, 1):                                                                                # :213-219
        n = F.layer_norm(h, (c,), sd[t + f"norms.{i}.weight"], sd[t + f"norms.{i}.bias"], 1e-5)
        a = t + f"attention_blocks.{i}."
        pe = sd.get(a + "pos_encoder.pe")
        h = versatile_attention(n, f, pe, sd[a + "to_q.weight"], sd[a + "to_k.weight"], sd[a + "to_v.weight"],
                                sd[a + "to_out.0.weight"], sd[a + "to_out.0.bias"], heads) + h
    n = F.layer_norm(h, (c,), sd[t + "ff_norm.weight"], sd[t + "ff_norm.bias"], 1e-5)               # :221
    h = geglu_feedforward(n, sd[t + "ff.net.0.proj.weight"], sd[t + "ff.net.0.proj.bias"],
                          sd[t + "ff.net.2.weight"], sd[t + "ff.net.2.bias"]) + h
    h = F.linear(h, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])                             # :155
    y = h.reshape(b, f, hh, ww, c).permute(0, 4, 1, 2, 3)                                           # :156-159
    return y + x


# ----------------------------------------------------------------------------------------------
# ResnetBlock3D (host of kernel (2))
# ----------------------------------------------------------------------------------------------
def conv2d_per_frame(x: Tensor, w: Tensor, bias: Optional[Tensor], stride: int = 1, padding: int = 1) -> Tensor:
    """`InflatedConv3d.forward` resnet.py:12-20: a 2-D conv applied to every frame."""
    b, c, f, h, ww = x.shape
    y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, ww), w, bias, stride=stride, padding=padding)
    return y.reshape(b, f, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def resnet_block3d(x: Tensor, temb: Optional[Tensor], sd: Dict[str, Tensor], prefix: str = "", groups: int = 32,
                   eps: float = 1e-5, per_frame: bool = True, output_scale_factor: float = 1.0) -> Tensor:
    """`ResnetBlock3D.forward` resnet.py:188-218 (time_embedding_norm="default", swish)."""
    h = groupnorm_silu(x, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], groups, eps, per_frame)
    h = conv2d_per_frame(h, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"])
    t = None
    if temb is not None:
        t = F.linear(F.silu(temb), sd[prefix + "time_emb_proj.weight"], sd[prefix + "time_emb_proj.bias"])
    h = groupnorm_silu(h, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], groups, eps, per_frame, temb=t)
    h = conv2d_per_frame(h, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"])
    if prefix + "conv_shortcut.weight" in sd:
        x = conv2d_per_frame(x, sd[prefix + "conv_shortcut.weight"], sd[prefix + "conv_shortcut.bias"], padding=0)
    return (x + h) / output_scale_factor


# ----------------------------------------------------------------------------------------------
# kernel (3): Multi-ControlNet residual scale / sum / layout / add
# ----------------------------------------------------------------------------------------------
def controlnet_scale_and_sum(per_net: Sequence[Sequence[Tensor]], scales: Sequence[float], guess_mode: bool = False
                             ) -> List[Tensor]:
    """Σ_k scale_k · r_{k,i} for the 13 residuals (12 down + mid, in that order), each [(b f), c, h, w].

    PARITY UNPINNED (third-party: diffusers==0.23.0, env.yml:120).  Published semantics restated:
    `ControlNetModel.forward` multiplies every residual by `conditioning_scale`; in guess mode
    (without global pooling) by `logspace(-1, 0, 13)[i] * conditioning_scale` instead, the mid
    residual taking the last factor; `MultiControlNetModel.forward` initialises the accumulators
    with the first net and adds the others element-wise.  Call site:
    modules/controlresiduals_pipeline.py:294-302.
    """
    n_res = len(per_net[0])
    level = torch.logspace(-1, 0, n_res) if guess_mode else torch.ones(n_res)
    out: List[Tensor] = []
    for i in range(n_res):
        acc = None
        for k, res in enumerate(per_net):
            term = res[i] * (float(level[i]) * float(scales[k]))
            acc = term if acc is None else acc + term
        out.append(acc)
    return out


def residuals_to_video_layout(res: Sequence[Tensor], frame_count: int) -> List[Tensor]:
    """'(b f) c h w -> b c f h w' for each residual — controlresiduals_pipeline.py:308-312."""
    out = []
    for r in res:
        bf, c, h, w = r.shape
        out.append(r.reshape(bf // frame_count, frame_count, c, h, w).permute(0, 2, 1, 3, 4))
    return out


def merge_controlnet_residuals(per_net: Sequence[Sequence[Tensor]], scales: Sequence[float], frame_count: int,
                               guess_mode: bool = False) -> Tuple[Tuple[Tensor, ...], Tensor]:
    """The B3 producer contract: per-net raw residual lists -> (down_block_additional_residuals[12],
    mid_block_additional_residual) in [b, c, f, h, w] — controlresiduals_pipeline.py:294-316."""
    merged = residuals_to_video_layout(controlnet_scale_and_sum(per_net, scales, guess_mode), frame_count)
    return tuple(merged[:-1]), merged[-1]


def add_residuals_to_skips(skips: Sequence[Tensor], mid: Tensor, down_res: Sequence[Tensor], mid_res: Tensor
                           ) -> Tuple[List[Tensor], Tensor]:
    """skip_i + res_i (12×) and sample + mid_res — unet.py:567-576, 584-585 (batch broadcast allowed)."""
    return [s + r for s, r in zip(skips, down_res)], mid + mid_res


# ----------------------------------------------------------------------------------------------
# loop-level glue (A1): DDIM + CFG, restated from SURVEY Appendix A.4 (diffusers DDIMScheduler,
# PARITY UNPINNED — third-party) and controlanimation_pipeline.py:844-849.
# ----------------------------------------------------------------------------------------------
def ddim_alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> Tensor:
    betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)  # "linear", inference-v2.yaml:24-27
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps(num_inference_steps: int, num_train_timesteps: int = 1000, steps_offset: int = 1) -> List[int]:
    ratio = num_train_timesteps // num_inference_steps
    return [int(round(k * ratio)) + steps_offset for k in range(num_inference_steps)][::-1]


def cfg_combine(noise_pred: Tensor, guidance_scale: float) -> Tensor:
    """controlanimation_pipeline.py:845-846."""
    u, c = noise_pred.chunk(2)
    return u + guidance_scale * (c - u)


def ddim_step(noise: Tensor, t: int, sample: Tensor, alphas_cumprod: Tensor, num_inference_steps: int,
              num_train_timesteps: int = 1000) -> Tensor:
    """η=0 DDIM update (epsilon prediction, clip_sample False)."""
    prev_t = t - num_train_timesteps // num_inference_steps
    a_t = alphas_cumprod[t]
    a_prev = alphas_cumprod[prev_t] if prev_t >= 0 else torch.tensor(1.0)
    x0 = (sample - (1 - a_t).sqrt() * noise) / a_t.sqrt()
    return a_prev.sqrt() * x0 + (1 - a_prev).sqrt() * noise
