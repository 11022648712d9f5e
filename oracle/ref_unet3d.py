"""CPU oracle: whole `UNet3DConditionModel.forward` and the SD1.5 `ControlNetModel.forward`,
restated functionally over a state_dict in plain torch fp32.  TEST INFRASTRUCTURE (see ref_ops.py).

UNet3D follows /root/reference/animatediff/models/unet.py:458-621 with the block forwards of
unet_blocks.py (:273-280 mid, :384-423 / :495-523 down, :623-669 / :737-762 up), the per-frame
spatial transformer of attention.py:120-167 + :254-300, `ResnetBlock3D` (ref_ops.resnet_block3d)
and the motion module (ref_ops.motion_module).  PINNED against the reference's own Python by
tests/golden/unet3d_tiny.npz.

ControlNet restates diffusers==0.23.0 `ControlNetModel` (third-party, absent from
/root/reference): PARITY UNPINNED; structure from SURVEY.md Appendix A.3, call site
modules/controlresiduals_pipeline.py:294-302.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import ref_ops as R

Tensor = torch.Tensor


def timestep_embedding(timesteps: Tensor, dim: int, flip_sin_to_cos: bool = True, freq_shift: float = 0.0) -> Tensor:
    """diffusers `Timesteps`: [cos, sin] halves when flip_sin_to_cos."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / (half - freq_shift))
    e = timesteps[:, None].float() * freqs[None]
    s, c = torch.sin(e), torch.cos(e)
    return torch.cat([c, s], -1) if flip_sin_to_cos else torch.cat([s, c], -1)


def time_embedding(sd: Dict[str, Tensor], t_emb: Tensor, cond: Optional[Tensor] = None, prefix="time_embedding.") -> Tensor:
    t_emb = t_emb.to(sd[prefix + "linear_1.weight"].dtype)      # unet.py:532 `t_emb.to(dtype=self.dtype)` (identity in fp32)
    if cond is not None:
        t_emb = t_emb + F.linear(cond, sd[prefix + "cond_proj.weight"])
    h = F.linear(t_emb, sd[prefix + "linear_1.weight"], sd[prefix + "linear_1.bias"])
    return F.linear(F.silu(h), sd[prefix + "linear_2.weight"], sd[prefix + "linear_2.bias"])


def spatial_transformer(x: Tensor, ctx: Tensor, sd: Dict[str, Tensor], prefix: str, heads: int, groups: int = 32,
                        ip: Optional[dict] = None) -> Tensor:
    """`Transformer3DModel.forward` attention.py:120-167 (use_linear_projection False) with one
    `BasicTransformerBlock` (:254-300: self-attn, cross-attn, GEGLU FF; no temporal attention).

    x [b,c,f,h,w]; ctx [b, n, cross_dim] is repeated over frames (:125).
    ip = dict(sd=..., num_tokens=4, scale=1.0): attn2 runs the IP-Adapter processor (modules/attention_processor.py:367-492,
    installed on every cross-attention by modules/ip_adapter.py:95-126); its weights are keyed
    `<prefix>transformer_blocks.0.attn2.processor.to_{k,v}_ip.weight` in ip["sd"].
    """
    b, c, f, hh, ww = x.shape
    g = R.groupnorm(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], groups, 1e-6, per_frame=True)
    h = R.conv2d_per_frame(g, sd[prefix + "proj_in.weight"], sd[prefix + "proj_in.bias"], padding=0)
    h = h.permute(0, 2, 3, 4, 1).reshape(b * f, hh * ww, c)
    ctx_f = ctx[:, None].expand(b, f, *ctx.shape[1:]).reshape(b * f, *ctx.shape[1:])
    t = prefix + "transformer_blocks.0."

    def attn(name, xin, context):
        a = t + name + "."
        return R.attention_processor(xin, sd[a + "to_q.weight"], sd[a + "to_k.weight"], sd[a + "to_v.weight"],
                                     sd[a + "to_out.0.weight"], sd[a + "to_out.0.bias"], heads, context)

    n = F.layer_norm(h, (c,), sd[t + "norm1.weight"], sd[t + "norm1.bias"], 1e-5)
    h = attn("attn1", n, None) + h
    n = F.layer_norm(h, (c,), sd[t + "norm2.weight"], sd[t + "norm2.bias"], 1e-5)
    if ip is None:
        h = attn("attn2", n, ctx_f) + h
    else:
        a = t + "attn2."
        h = R.ip_attention_processor(n, ctx_f, sd[a + "to_q.weight"], sd[a + "to_k.weight"], sd[a + "to_v.weight"],
                                     sd[a + "to_out.0.weight"], sd[a + "to_out.0.bias"], ip["sd"][a + "processor.to_k_ip.weight"],
                                     ip["sd"][a + "processor.to_v_ip.weight"], heads, ip["num_tokens"], ip["scale"]) + h
    n = F.layer_norm(h, (c,), sd[t + "norm3.weight"], sd[t + "norm3.bias"], 1e-5)
    h = R.geglu_feedforward(n, sd[t + "ff.net.0.proj.weight"], sd[t + "ff.net.0.proj.bias"],
                            sd[t + "ff.net.2.weight"], sd[t + "ff.net.2.bias"]) + h
    h = h.reshape(b, f, hh, ww, c).permute(0, 4, 1, 2, 3)
    h = R.conv2d_per_frame(h, sd[prefix + "proj_out.weight"], sd[prefix + "proj_out.bias"], padding=0)
    return h + x


def _upsample_nearest(x: Tensor, size: Optional[Tuple[int, int]]) -> Tensor:
    """`Upsample3D.forward` resnet.py:63-66: nearest ×2 on (h, w), or forced output size."""
    b, c, f, h, w = x.shape
    x4 = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    y = F.interpolate(x4, scale_factor=2.0, mode="nearest") if size is None else F.interpolate(x4, size=size, mode="nearest")
    return y.reshape(b, f, c, *y.shape[-2:]).permute(0, 2, 1, 3, 4)


def unet3d_forward(sd: Dict[str, Tensor], cfg: dict, sample: Tensor, timestep, encoder_hidden_states: Tensor,
                   down_block_additional_residuals: Optional[Sequence[Tensor]] = None,
                   mid_block_additional_residual: Optional[Tensor] = None,
                   timestep_cond: Optional[Tensor] = None, ip: Optional[dict] = None) -> Tensor:
    """`UNet3DConditionModel.forward` unet.py:458-621.  sample [b,4,f,h,w] -> [b,4,f,h,w].  `ip`: see spatial_transformer."""
    boc = tuple(cfg["block_out_channels"])
    lpb = cfg["layers_per_block"]
    groups, eps = cfg["norm_num_groups"], cfg["norm_eps"]
    heads = cfg["attention_head_dim"]
    per_frame = bool(cfg.get("use_inflated_groupnorm", False))
    mm_heads = cfg["motion_module_kwargs"]["num_attention_heads"]
    use_mm = cfg.get("use_motion_module", False)
    mm_res = tuple(cfg.get("motion_module_resolutions", (1, 2, 4, 8)))
    b = sample.shape[0]

    upf = 2 ** (len(boc) - 1)
    forward_upsample_size = any(s % upf != 0 for s in sample.shape[-2:])                     # :491-499

    ts = torch.as_tensor(timestep, device=sample.device).reshape(-1).expand(b)               # :510-524
    emb = time_embedding(sd, timestep_embedding(ts, boc[0]), timestep_cond)                  # :526-534

    x = R.conv2d_per_frame(sample, sd["conv_in.weight"], sd["conv_in.bias"])                 # :547
    skips = [x]

    def resnet(x, prefix):
        return R.resnet_block3d(x, emb, sd, prefix, groups, eps, per_frame)

    def motion(x, prefix, enabled):
        return R.motion_module(x, sd, prefix, mm_heads) if enabled else x

    for i, btype in enumerate(cfg["down_block_types"]):                                      # :551-562
        p = f"down_blocks.{i}."
        mm = use_mm and (2 ** i in mm_res) and not cfg.get("motion_module_decoder_only", False)
        for j in range(lpb):
            x = resnet(x, p + f"resnets.{j}.")
            if btype == "CrossAttnDownBlock3D":
                x = spatial_transformer(x, encoder_hidden_states, sd, p + f"attentions.{j}.", heads, groups, ip)
            x = motion(x, p + f"motion_modules.{j}.", mm)
            skips.append(x)
        if i != len(boc) - 1:
            x = R.conv2d_per_frame(x, sd[p + "downsamplers.0.conv.weight"], sd[p + "downsamplers.0.conv.bias"], stride=2)
            skips.append(x)

    if down_block_additional_residuals is not None:                                          # :567-576
        skips = [s + r for s, r in zip(skips, down_block_additional_residuals)]

    x = resnet(x, "mid_block.resnets.0.")                                                    # unet_blocks.py:273-280
    x = spatial_transformer(x, encoder_hidden_states, sd, "mid_block.attentions.0.", heads, groups, ip)
    x = motion(x, "mid_block.motion_modules.0.", use_mm and cfg.get("motion_module_mid_block", False))
    x = resnet(x, "mid_block.resnets.1.")
    if mid_block_additional_residual is not None:                                            # :584-585
        x = x + mid_block_additional_residual

    for i, btype in enumerate(cfg["up_block_types"]):                                        # :588-611
        p = f"up_blocks.{i}."
        final = i == len(boc) - 1
        mm = use_mm and (2 ** (3 - i) in mm_res)
        res = skips[-(lpb + 1):]
        skips = skips[:-(lpb + 1)]
        up_size = tuple(skips[-1].shape[-2:]) if (not final and forward_upsample_size) else None
        for j in range(lpb + 1):
            x = torch.cat([x, res.pop()], dim=1)
            x = resnet(x, p + f"resnets.{j}.")
            if btype == "CrossAttnUpBlock3D":
                x = spatial_transformer(x, encoder_hidden_states, sd, p + f"attentions.{j}.", heads, groups, ip)
            x = motion(x, p + f"motion_modules.{j}.", mm)
        if not final:
            x = _upsample_nearest(x, up_size)
            x = R.conv2d_per_frame(x, sd[p + "upsamplers.0.conv.weight"], sd[p + "upsamplers.0.conv.bias"])

    x = R.groupnorm_silu(x, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], groups, eps, per_frame)  # :614-615
    return R.conv2d_per_frame(x, sd["conv_out.weight"], sd["conv_out.bias"])                 # :616


# ----------------------------------------------------------------------------------------------
# ControlNet (diffusers 0.23.0 ControlNetModel, SD1.5 configuration) — PARITY UNPINNED
# ----------------------------------------------------------------------------------------------
COND_EMBED_CHANNELS = (16, 32, 96, 256)


def controlnet_forward(sd: Dict[str, Tensor], cfg: dict, sample: Tensor, timestep, encoder_hidden_states: Tensor,
                       controlnet_cond: Tensor) -> List[Tensor]:
    """Raw (unscaled) ControlNet outputs: 12 down residuals + mid, each [n, c, h, w].

    sample [n,4,h,w] (n = b·f frames), encoder_hidden_states [n,L,768], controlnet_cond [n,3,8h,8w].
    The conditioning scale / guess-mode factors are applied by ref_ops.controlnet_scale_and_sum.
    """
    boc = tuple(cfg["block_out_channels"])
    lpb = cfg["layers_per_block"]
    groups, eps, heads = cfg["norm_num_groups"], cfg["norm_eps"], cfg["attention_head_dim"]
    n = sample.shape[0]
    ts = torch.as_tensor(timestep, device=sample.device).reshape(-1).expand(n)
    emb = time_embedding(sd, timestep_embedding(ts, boc[0]))

    def v(x):  # [n,c,h,w] -> [n,c,1,h,w] so the per-frame 3-D helpers apply with f = 1
        return x[:, :, None]

    x = v(sample)
    x = R.conv2d_per_frame(x, sd["conv_in.weight"], sd["conv_in.bias"])
    c = v(controlnet_cond)
    ce = "controlnet_cond_embedding."
    c = F.silu(R.conv2d_per_frame(c, sd[ce + "conv_in.weight"], sd[ce + "conv_in.bias"]))
    for k in range(2 * (len(COND_EMBED_CHANNELS) - 1)):
        c = F.silu(R.conv2d_per_frame(c, sd[ce + f"blocks.{k}.weight"], sd[ce + f"blocks.{k}.bias"], stride=1 + (k % 2)))
    c = R.conv2d_per_frame(c, sd[ce + "conv_out.weight"], sd[ce + "conv_out.bias"])
    x = x + c

    skips = [x]
    for i in range(len(boc)):
        p = f"down_blocks.{i}."
        for j in range(lpb):
            x = R.resnet_block3d(x, emb, sd, p + f"resnets.{j}.", groups, eps, True)
            if i != len(boc) - 1:
                x = spatial_transformer(x, encoder_hidden_states, sd, p + f"attentions.{j}.", heads, groups)
            skips.append(x)
        if i != len(boc) - 1:
            x = R.conv2d_per_frame(x, sd[p + "downsamplers.0.conv.weight"], sd[p + "downsamplers.0.conv.bias"], stride=2)
            skips.append(x)
    x = R.resnet_block3d(x, emb, sd, "mid_block.resnets.0.", groups, eps, True)
    x = spatial_transformer(x, encoder_hidden_states, sd, "mid_block.attentions.0.", heads, groups)
    x = R.resnet_block3d(x, emb, sd, "mid_block.resnets.1.", groups, eps, True)

    out = []
    for i, s in enumerate(skips):
        out.append(R.conv2d_per_frame(s, sd[f"controlnet_down_blocks.{i}.weight"], sd[f"controlnet_down_blocks.{i}.bias"],
                                      padding=0)[:, :, 0])
    out.append(R.conv2d_per_frame(x, sd["controlnet_mid_block.weight"], sd["controlnet_mid_block.bias"], padding=0)[:, :, 0])
    return out


# ----------------------------------------------------------------------------------------------
# state_dict shape tables (so oracle and product can be filled by oracle.synth without a module)
# ----------------------------------------------------------------------------------------------
def _resnet_shapes(p, cin, cout, temb):
    s = {p + "norm1.weight": (cin,), p + "norm1.bias": (cin,), p + "conv1.weight": (cout, cin, 3, 3), p + "conv1.bias": (cout,),
         p + "time_emb_proj.weight": (cout, temb), p + "time_emb_proj.bias": (cout,),
         p + "norm2.weight": (cout,), p + "norm2.bias": (cout,), p + "conv2.weight": (cout, cout, 3, 3), p + "conv2.bias": (cout,)}
    if cin != cout:
        s[p + "conv_shortcut.weight"] = (cout, cin, 1, 1)
        s[p + "conv_shortcut.bias"] = (cout,)
    return s


def _attn_shapes(a, c, ctx):
    return {a + "to_q.weight": (c, c), a + "to_k.weight": (c, ctx), a + "to_v.weight": (c, ctx),
            a + "to_out.0.weight": (c, c), a + "to_out.0.bias": (c,)}


def _transformer_shapes(p, c, cross):
    s = {p + "norm.weight": (c,), p + "norm.bias": (c,), p + "proj_in.weight": (c, c, 1, 1), p + "proj_in.bias": (c,),
         p + "proj_out.weight": (c, c, 1, 1), p + "proj_out.bias": (c,)}
    t = p + "transformer_blocks.0."
    for k in (1, 2, 3):
        s[t + f"norm{k}.weight"] = (c,)
        s[t + f"norm{k}.bias"] = (c,)
    s.update(_attn_shapes(t + "attn1.", c, c))
    s.update(_attn_shapes(t + "attn2.", c, cross))
    s.update({t + "ff.net.0.proj.weight": (8 * c, c), t + "ff.net.0.proj.bias": (8 * c,),
              t + "ff.net.2.weight": (c, 4 * c), t + "ff.net.2.bias": (c,)})
    return s


def motion_module_shapes(p, c, max_len=32):
    p = p + "temporal_transformer."
    s = {p + "norm.weight": (c,), p + "norm.bias": (c,), p + "proj_in.weight": (c, c), p + "proj_in.bias": (c,),
         p + "proj_out.weight": (c, c), p + "proj_out.bias": (c,)}
    t = p + "transformer_blocks.0."
    for i in (0, 1):
        s.update(_attn_shapes(t + f"attention_blocks.{i}.", c, c))
        s[t + f"attention_blocks.{i}.pos_encoder.pe"] = (1, max_len, c)
        s[t + f"norms.{i}.weight"] = (c,)
        s[t + f"norms.{i}.bias"] = (c,)
    s.update({t + "ff.net.0.proj.weight": (8 * c, c), t + "ff.net.0.proj.bias": (8 * c,),
              t + "ff.net.2.weight": (c, 4 * c), t + "ff.net.2.bias": (c,),
              t + "ff_norm.weight": (c,), t + "ff_norm.bias": (c,)})
    return s


def controlnet_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    boc = tuple(cfg["block_out_channels"])
    lpb, cross = cfg["layers_per_block"], cfg["cross_attention_dim"]
    temb = boc[0] * 4
    s = {"conv_in.weight": (boc[0], 4, 3, 3), "conv_in.bias": (boc[0],),
         "time_embedding.linear_1.weight": (temb, boc[0]), "time_embedding.linear_1.bias": (temb,),
         "time_embedding.linear_2.weight": (temb, temb), "time_embedding.linear_2.bias": (temb,)}
    ce = "controlnet_cond_embedding."
    ch = COND_EMBED_CHANNELS
    s[ce + "conv_in.weight"], s[ce + "conv_in.bias"] = (ch[0], 3, 3, 3), (ch[0],)
    for i in range(len(ch) - 1):
        s[ce + f"blocks.{2 * i}.weight"], s[ce + f"blocks.{2 * i}.bias"] = (ch[i], ch[i], 3, 3), (ch[i],)
        s[ce + f"blocks.{2 * i + 1}.weight"], s[ce + f"blocks.{2 * i + 1}.bias"] = (ch[i + 1], ch[i], 3, 3), (ch[i + 1],)
    s[ce + "conv_out.weight"], s[ce + "conv_out.bias"] = (boc[0], ch[-1], 3, 3), (boc[0],)
    cin = boc[0]
    k = 0
    s[f"controlnet_down_blocks.{k}.weight"], s[f"controlnet_down_blocks.{k}.bias"] = (cin, cin, 1, 1), (cin,)
    for i, c in enumerate(boc):
        p = f"down_blocks.{i}."
        for j in range(lpb):
            s.update(_resnet_shapes(p + f"resnets.{j}.", cin, c, temb))
            cin = c
            if i != len(boc) - 1:
                s.update(_transformer_shapes(p + f"attentions.{j}.", c, cross))
            k += 1
            s[f"controlnet_down_blocks.{k}.weight"], s[f"controlnet_down_blocks.{k}.bias"] = (c, c, 1, 1), (c,)
        if i != len(boc) - 1:
            s[p + "downsamplers.0.conv.weight"], s[p + "downsamplers.0.conv.bias"] = (c, c, 3, 3), (c,)
            k += 1
            s[f"controlnet_down_blocks.{k}.weight"], s[f"controlnet_down_blocks.{k}.bias"] = (c, c, 1, 1), (c,)
    c = boc[-1]
    s.update(_resnet_shapes("mid_block.resnets.0.", c, c, temb))
    s.update(_transformer_shapes("mid_block.attentions.0.", c, cross))
    s.update(_resnet_shapes("mid_block.resnets.1.", c, c, temb))
    s["controlnet_mid_block.weight"], s["controlnet_mid_block.bias"] = (c, c, 1, 1), (c,)
    return s


def ip_adapter_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    """`to_k_ip` / `to_v_ip` of every cross-attention processor (modules/ip_adapter.py:95-126), keyed like `attn_processors`."""
    s = {}
    for k in unet3d_shapes(cfg):
        if k.endswith("attn2.to_k.weight"):
            c = unet3d_shapes(cfg)[k][0]
            base = k[:-len("to_k.weight")] + "processor."
            s[base + "to_k_ip.weight"] = (c, cfg["cross_attention_dim"])
            s[base + "to_v_ip.weight"] = (c, cfg["cross_attention_dim"])
    return s


def unet3d_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    """Every parameter/buffer the oracle UNet3D reads, keyed like the reference state_dict
    (the reference additionally carries unused `to_q/to_k/to_v/to_out` on each
    BasicTransformerBlock because of the inheritance quirk at attention.py:170,187)."""
    boc = tuple(cfg["block_out_channels"])
    lpb, cross = cfg["layers_per_block"], cfg["cross_attention_dim"]
    temb = boc[0] * 4
    max_len = cfg["motion_module_kwargs"]["temporal_position_encoding_max_len"]
    use_mm = cfg.get("use_motion_module", False)
    s = {"conv_in.weight": (boc[0], cfg["in_channels"], 3, 3), "conv_in.bias": (boc[0],),
         "time_embedding.linear_1.weight": (temb, boc[0]), "time_embedding.linear_1.bias": (temb,),
         "time_embedding.linear_2.weight": (temb, temb), "time_embedding.linear_2.bias": (temb,)}
    if cfg.get("time_cond_proj_dim"):
        s["time_embedding.cond_proj.weight"] = (boc[0], cfg["time_cond_proj_dim"])
    cin = boc[0]
    skip_ch = [boc[0]]
    for i, (c, bt) in enumerate(zip(boc, cfg["down_block_types"])):
        p = f"down_blocks.{i}."
        for j in range(lpb):
            s.update(_resnet_shapes(p + f"resnets.{j}.", cin, c, temb))
            cin = c
            if bt == "CrossAttnDownBlock3D":
                s.update(_transformer_shapes(p + f"attentions.{j}.", c, cross))
            if use_mm:
                s.update(motion_module_shapes(p + f"motion_modules.{j}.", c, max_len))
            skip_ch.append(c)
        if i != len(boc) - 1:
            s[p + "downsamplers.0.conv.weight"], s[p + "downsamplers.0.conv.bias"] = (c, c, 3, 3), (c,)
            skip_ch.append(c)
    c = boc[-1]
    s.update(_resnet_shapes("mid_block.resnets.0.", c, c, temb))
    s.update(_transformer_shapes("mid_block.attentions.0.", c, cross))
    if use_mm and cfg.get("motion_module_mid_block", False):
        s.update(motion_module_shapes("mid_block.motion_modules.0.", c, max_len))
    s.update(_resnet_shapes("mid_block.resnets.1.", c, c, temb))
    rev = list(reversed(boc))
    prev = rev[0]
    for i, bt in enumerate(cfg["up_block_types"]):
        p = f"up_blocks.{i}."
        c = rev[i]
        for j in range(lpb + 1):
            s.update(_resnet_shapes(p + f"resnets.{j}.", prev + skip_ch.pop(), c, temb))
            prev = c
            if bt == "CrossAttnUpBlock3D":
                s.update(_transformer_shapes(p + f"attentions.{j}.", c, cross))
            if use_mm:
                s.update(motion_module_shapes(p + f"motion_modules.{j}.", c, max_len))
        if i != len(boc) - 1:
            s[p + "upsamplers.0.conv.weight"], s[p + "upsamplers.0.conv.bias"] = (c, c, 3, 3), (c,)
    s["conv_norm_out.weight"], s["conv_norm_out.bias"] = (boc[0],), (boc[0],)
    s["conv_out.weight"], s["conv_out.bias"] = (cfg["out_channels"], boc[0], 3, 3), (cfg["out_channels"],)
    return s


def synth_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int) -> Dict[str, Tensor]:
    """Fill a shape table with oracle.synth values (PE buffers analytic)."""
    from . import synth
    out = {}
    for k, shp in shapes.items():
        if k.endswith(".pe"):
            out[k] = R.positional_encoding(shp[1], shp[2])
        else:
            out[k] = synth.synth_value(seed, k, torch.empty(shp, device="meta"))
    return out
