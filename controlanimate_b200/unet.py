"""UNet3DConditionModel and ControlNetModel hosts for the B200 kernels.

`UNet3DConditionModel` mirrors reference animatediff/models/unet.py:50-621 (ctor kwargs, state_dict keys,
`forward(sample, timestep, encoder_hidden_states, ..., down_block_additional_residuals,
mid_block_additional_residual, timestep_cond)`, `attn_processors`/`set_attn_processor`) with blocks that mirror
unet_blocks.py.  Internally every activation is a channels_last [(b f), C, h, w] tensor (see layers.py).

`ControlNetModel` restates the SD1.5 diffusers ControlNetModel (third-party) on
the same blocks; it returns RAW residuals, the conditioning scales are applied by kernel (3).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import layers as Ly
from . import ops
from .residuals import ResidualSet


@dataclass
class UNet3DConditionOutput:
    sample: torch.Tensor


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos=True, freq_shift=0):
        super().__init__()
        self.num_channels, self.flip, self.shift = num_channels, flip_sin_to_cos, freq_shift

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / (half - self.shift)
        e = timesteps[:, None].float() * torch.exp(exponent)[None]
        s, c = torch.sin(e), torch.cos(e)
        return torch.cat([c, s], -1) if self.flip else torch.cat([s, c], -1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim, cond_proj_dim=None):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.cond_proj = nn.Linear(cond_proj_dim, in_channels, bias=False) if cond_proj_dim else None
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample, condition=None):
        if condition is not None:
            sample = sample + self.cond_proj(condition)
        return self.linear_2(F.silu(self.linear_1(sample)))


class _Sampler(nn.Module):
    """Downsample3D / Upsample3D container (key `conv`), resnet.py:34-108."""

    def __init__(self, channels, stride):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=stride, padding=1)


def _mm(channels, kwargs, enabled):
    return Ly.B200MotionModule(in_channels=channels, **kwargs) if enabled else None


class _Block(nn.Module):
    """One Down/Mid/Up block: resnet -> [spatial transformer] -> [motion module] per layer."""

    def __init__(self, resnets, attentions, motion_modules, down=None, up=None):
        super().__init__()
        self.resnets = nn.ModuleList(resnets)
        if attentions is not None:
            self.attentions = nn.ModuleList(attentions)
        self.has_cross_attention = attentions is not None
        self.motion_modules = nn.ModuleList(motion_modules) if motion_modules is not None else None
        if down is not None:
            self.downsamplers = nn.ModuleList([down])
        if up is not None:
            self.upsamplers = nn.ModuleList([up])

    def layer(self, j, x, temb, frames, ctx, ctx_map, shift=None):
        x = self.resnets[j].native(x, temb, frames, shift)
        if self.has_cross_attention:
            x = self.attentions[j].native(x, frames, ctx, ctx_map)
        if self.motion_modules is not None and self.motion_modules[j] is not None:
            x = self.motion_modules[j].native(x, frames)
        return x


class UNet3DConditionModel(nn.Module):
    def __init__(self, sample_size=None, in_channels=4, out_channels=4, flip_sin_to_cos=True, freq_shift=0,
                 down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
                 mid_block_type="UNetMidBlock3DCrossAttn",
                 up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
                 block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, norm_num_groups=32, norm_eps=1e-5,
                 cross_attention_dim=1280, attention_head_dim=8, use_inflated_groupnorm=False, time_cond_proj_dim=None,
                 use_motion_module=False, motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=False,
                 motion_module_decoder_only=False, motion_module_type=None, motion_module_kwargs=None,
                 unet_use_cross_frame_attention=None, unet_use_temporal_attention=None, mid_block_scale_factor=1, **unused):
        super().__init__()
        if unet_use_cross_frame_attention or unet_use_temporal_attention:
            raise ValueError("cross-frame / in-block temporal attention are not part of the shipped configs")
        if use_motion_module and motion_module_type != "Vanilla":
            raise ValueError("motion_module_type must be 'Vanilla' (motion_module.py:44-47)")
        mmk = dict(motion_module_kwargs or {})
        self.config = dict(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                           block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                           cross_attention_dim=cross_attention_dim, attention_head_dim=attention_head_dim,
                           norm_num_groups=norm_num_groups, norm_eps=norm_eps, use_inflated_groupnorm=use_inflated_groupnorm,
                           time_cond_proj_dim=time_cond_proj_dim, center_input_sample=False)
        self.sample_size = sample_size
        boc = tuple(block_out_channels)
        temb_dim = boc[0] * 4
        heads = attention_head_dim
        self.per_frame = bool(use_inflated_groupnorm)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_proj = Timesteps(boc[0], flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(boc[0], temb_dim, time_cond_proj_dim)

        def resnet(cin, cout):
            return Ly.B200ResnetBlock3D(in_channels=cin, out_channels=cout, temb_channels=temb_dim, groups=norm_num_groups,
                                        eps=norm_eps, use_inflated_groupnorm=use_inflated_groupnorm,
                                        output_scale_factor=1.0)

        def xf(c):
            return Ly.SpatialTransformer3D(heads, c, cross_attention_dim, norm_num_groups)

        self.down_blocks = nn.ModuleList()
        cout = boc[0]
        for i, bt in enumerate(down_block_types):
            cin, cout = cout, boc[i]
            mm_on = use_motion_module and (2 ** i in motion_module_resolutions) and not motion_module_decoder_only
            rs = [resnet(cin if j == 0 else cout, cout) for j in range(layers_per_block)]
            at = [xf(cout) for _ in range(layers_per_block)] if bt == "CrossAttnDownBlock3D" else None
            mm = [_mm(cout, mmk, mm_on) for _ in range(layers_per_block)]
            self.down_blocks.append(_Block(rs, at, mm, down=_Sampler(cout, 2) if i != len(boc) - 1 else None))

        c = boc[-1]
        self.mid_block = _Block([resnet(c, c), resnet(c, c)], [xf(c)], [_mm(c, mmk, use_motion_module and motion_module_mid_block)])

        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        cout = rev[0]
        self.num_upsamplers = 0
        for i, bt in enumerate(up_block_types):
            prev, cout = cout, rev[i]
            cin = rev[min(i + 1, len(boc) - 1)]
            n_layers = layers_per_block + 1
            mm_on = use_motion_module and (2 ** (3 - i) in motion_module_resolutions)
            rs = []
            for j in range(n_layers):
                skip_c = cin if j == n_layers - 1 else cout
                rs.append(resnet((prev if j == 0 else cout) + skip_c, cout))
            at = [xf(cout) for _ in range(n_layers)] if bt == "CrossAttnUpBlock3D" else None
            mm = [_mm(cout, mmk, mm_on) for _ in range(n_layers)]
            final = i == len(boc) - 1
            if not final:
                self.num_upsamplers += 1
            self.up_blocks.append(_Block(rs, at, mm, up=None if final else _Sampler(cout, 1)))

        self.conv_norm_out = Ly.InflatedGroupNormParams(norm_num_groups, boc[0], eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self._temb_bank = Ly.TembBank()

    def _resnets_in_order(self):
        rs = [r for blk in self.down_blocks for r in blk.resnets] + list(self.mid_block.resnets)
        return rs + [r for blk in self.up_blocks for r in blk.resnets]

    # ---- processor protocol (unet.py:320-382) ------------------------------------------------------------
    @property
    def attn_processors(self) -> Dict[str, object]:
        out = {}
        for name, m in self.named_modules():
            if hasattr(m, "get_processor"):
                out[f"{name}.processor"] = m.get_processor(return_deprecated_lora=True)
        return out

    def set_attn_processor(self, processor, _remove_lora=False):
        mods = {f"{n}.processor": m for n, m in self.named_modules() if hasattr(m, "set_processor")}
        if isinstance(processor, dict):
            if len(processor) != len(mods):
                raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does not match"
                                 f" the number of attention layers: {len(mods)}.")
            for k, m in mods.items():
                m.set_processor(processor[k])
        else:
            for m in mods.values():
                m.set_processor(processor)

    def load_reference_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Load a reference UNet3DConditionModel state_dict.  The reference additionally carries unused
        `transformer_blocks.N.to_{q,k,v,out}` parameters (BasicTransformerBlock subclasses Attention, attention.py:170,187);
        those are dropped."""
        import re
        stray = re.compile(r"transformer_blocks\.\d+\.(to_q|to_k|to_v|to_out)\.")
        own = self.state_dict()
        sd = {k: v for k, v in sd.items() if not (stray.search(k) and "attention_blocks" not in k and k not in own)}
        return self.load_state_dict(sd, strict=True)

    # ---- forward (unet.py:458-621) -------------------------------------------------------------------------
    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, attention_mask=None, return_dict=True,
                down_block_additional_residuals=None, mid_block_additional_residual=None, timestep_cond=None,
                cross_attention_kwargs=None):
        if attention_mask is not None or class_labels is not None:
            raise ValueError("attention_mask / class_labels are not used on this path")
        if not sample.is_cuda:
            raise ValueError("controlanimate_b200 runs on CUDA only (no CPU fallback)")
        dtype = self.conv_in.weight.dtype
        b, _, frames, hh, ww = sample.shape
        upf = 2 ** self.num_upsamplers
        forward_upsample_size = any(s % upf != 0 for s in (hh, ww))                       # unet.py:491-499

        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.int64, device=sample.device)
        timesteps = timestep.reshape(-1).to(sample.device).expand(b)
        emb = self.time_embedding(self.time_proj(timesteps).to(dtype), timestep_cond)      # :526-534 (computed once)

        shifts = iter(self._temb_bank.shifts(self._resnets_in_order(), emb))               # resnet.py:196-200, batched

        ctx = encoder_hidden_states.to(dtype)
        x = Ly.frames4(Ly.to_native(sample.to(dtype)))
        x = Ly.conv_bias(self.conv_in, x)                                                  # :547
        skips = [x]
        for blk in self.down_blocks:                                                       # :551-562
            for j in range(len(blk.resnets)):
                x = blk.layer(j, x, emb, frames, ctx, None, next(shifts))
                skips.append(x)
            if hasattr(blk, "downsamplers"):
                x = Ly.conv_bias(blk.downsamplers[0].conv, x)
                skips.append(x)

        res = ResidualSet.coerce(down_block_additional_residuals, mid_block_additional_residual, frames)
        if res is not None:                                                                # :567-576 (kernel 3, in place)
            # the reference builds NEW skip tensors; the last skip aliases the mid-block input, which must stay un-added
            skips[-1] = skips[-1].clone(memory_format=torch.preserve_format)
            res.add_into(skips, None)

        mb = self.mid_block                                                                # unet_blocks.py:273-280
        x = mb.resnets[0].native(x, emb, frames, next(shifts))
        x = mb.attentions[0].native(x, frames, ctx, None)
        if mb.motion_modules[0] is not None:
            x = mb.motion_modules[0].native(x, frames)
        x = mb.resnets[1].native(x, emb, frames, next(shifts))
        if res is not None:                                                                # :584-585
            x = Ly._cl(x)
            res.add_into(None, x)

        for i, blk in enumerate(self.up_blocks):                                           # :588-611
            n = len(blk.resnets)
            rs, skips = skips[-n:], skips[:-n]
            for j in range(n):
                x = Ly.concat_channels(x, rs.pop())                                        # unet_blocks.py:636, :742
                x = blk.layer(j, x, emb, frames, ctx, None, next(shifts))
            if hasattr(blk, "upsamplers"):
                x = Ly.upsample_nearest(x, tuple(skips[-1].shape[-2:]) if forward_upsample_size else None)  # resnet.py:63-69
                x = Ly.conv_bias(blk.upsamplers[0].conv, x)

        x = Ly.group_norm(x, self.conv_norm_out, frames, per_frame=self.per_frame, silu=True)  # :614-615
        x = self.conv_out(x)                                                               # :616
        out = Ly.video5(x, frames).contiguous()  # hand the caller the reference's NCFHW layout
        if not return_dict:
            return (out,)
        return UNet3DConditionOutput(sample=out)


class ControlNetModel(nn.Module):
    """SD1.5 ControlNet (diffusers 0.23.0 `ControlNetModel`, restated; parity unpinned — third party).

    forward(sample [(b f),4,h,w], timestep, encoder_hidden_states [B_ctx, L, D], controlnet_cond [(b f),3,8h,8w],
            ctx_map) -> list of 13 RAW residuals, channels_last [(b f), c_i, h_i, w_i].
    """

    COND_CHANNELS = (16, 32, 96, 256)

    def __init__(self, in_channels=4, conditioning_channels=3, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=768, attention_head_dim=8, **unused):
        super().__init__()
        boc = tuple(block_out_channels)
        temb_dim = boc[0] * 4
        self.config = dict(block_out_channels=boc, layers_per_block=layers_per_block, cross_attention_dim=cross_attention_dim,
                           attention_head_dim=attention_head_dim, norm_num_groups=norm_num_groups, norm_eps=norm_eps)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_proj = Timesteps(boc[0], True, 0)
        self.time_embedding = TimestepEmbedding(boc[0], temb_dim)
        ce = nn.Module()
        ch = self.COND_CHANNELS
        ce.conv_in = nn.Conv2d(conditioning_channels, ch[0], 3, padding=1)
        blocks = []
        for i in range(len(ch) - 1):
            blocks += [nn.Conv2d(ch[i], ch[i], 3, padding=1), nn.Conv2d(ch[i], ch[i + 1], 3, padding=1, stride=2)]
        ce.blocks = nn.ModuleList(blocks)
        ce.conv_out = nn.Conv2d(ch[-1], boc[0], 3, padding=1)
        self.controlnet_cond_embedding = ce

        def resnet(cin, cout):
            return Ly.B200ResnetBlock3D(in_channels=cin, out_channels=cout, temb_channels=temb_dim, groups=norm_num_groups,
                                        eps=norm_eps, use_inflated_groupnorm=True)

        def xf(c):
            return Ly.SpatialTransformer3D(attention_head_dim, c, cross_attention_dim, norm_num_groups)

        self.down_blocks = nn.ModuleList()
        zero = [nn.Conv2d(boc[0], boc[0], 1)]
        cout = boc[0]
        for i in range(len(boc)):
            cin, cout = cout, boc[i]
            last = i == len(boc) - 1
            rs = [resnet(cin if j == 0 else cout, cout) for j in range(layers_per_block)]
            at = [xf(cout) for _ in range(layers_per_block)] if not last else None
            self.down_blocks.append(_Block(rs, at, None, down=None if last else _Sampler(cout, 2)))
            zero += [nn.Conv2d(cout, cout, 1) for _ in range(layers_per_block + (0 if last else 1))]
        self.controlnet_down_blocks = nn.ModuleList(zero)
        c = boc[-1]
        self.mid_block = _Block([resnet(c, c), resnet(c, c)], [xf(c)], None)
        self.controlnet_mid_block = nn.Conv2d(c, c, 1)
        self._temb_bank = Ly.TembBank()

    @staticmethod
    def _zero_conv(conv: nn.Conv2d, x4: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        n, c, h, w = x4.shape
        if out is not None:   # store straight into the caller's buffer (e.g. a symmetric-memory arena read by a peer GPU)
            if tuple(out.shape) != (n, c, h, w) or not out.permute(0, 2, 3, 1).is_contiguous() or out.dtype != x4.dtype:
                raise ValueError("ControlNet output buffers must be channels_last tensors of the residual shapes")
            out = Ly.tokens(out)
        y = ops.linear(Ly.tokens(Ly._cl(x4)), conv.weight.reshape(c, c), Ly.f32(conv.bias), out=out)   # 1x1 conv == token GEMM
        return Ly.from_tokens(y, n, h, w)

    def embed_condition(self, controlnet_cond: torch.Tensor) -> torch.Tensor:
        """`controlnet_cond_embedding(controlnet_cond)` WITHOUT conv_out's bias, channels_last [(b f), boc[0], h, w].  It depends
        on the prepared control video and the weights only — not on the latents or the timestep — so a caller that denoises the
        same window for many steps may compute it once per window and pass it to `forward(cond_embedding=...)`
        (`MultiControlNetResiduals.hoist_cond_embedding`); the reference (diffusers) re-evaluates it inside every step."""
        dtype = self.conv_in.weight.dtype
        ce = self.controlnet_cond_embedding
        c = Ly.conv_bias(ce.conv_in, Ly._cl(controlnet_cond.to(dtype)), silu=True)         # conv -> +bias -> SiLU: one epilogue
        for blk in ce.blocks:
            c = Ly.conv_bias(blk, c, silu=True)
        return Ly._cl(Ly.conv_nobias(ce.conv_out, c))

    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, ctx_map=None,
                out: Optional[Sequence[torch.Tensor]] = None, cond_embedding: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        dtype = self.conv_in.weight.dtype
        n = sample.shape[0]
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.int64, device=sample.device)
        emb = self.time_embedding(self.time_proj(timestep.reshape(-1).to(sample.device).expand(n)).to(dtype))
        ctx = encoder_hidden_states.to(dtype)
        if ctx_map is None and ctx.shape[0] != n:
            raise ValueError("encoder_hidden_states must have one row per frame, or pass ctx_map")
        resnets = [r for blk in self.down_blocks for r in blk.resnets] + list(self.mid_block.resnets)
        shifts = iter(self._temb_bank.shifts(resnets, emb))
        ce = self.controlnet_cond_embedding
        if cond_embedding is not None:
            if cond_embedding.shape[0] != n or cond_embedding.dtype != dtype:
                raise ValueError("cond_embedding must be embed_condition() of this call's control images")
            # conv_in(sample) + (hoisted) conv_out(c): both biases and the sum in one epilogue pass
            x = Ly.conv_bias(self.conv_in, Ly._cl(sample.to(dtype)), residual=cond_embedding, extra_bias=ce.conv_out.bias)
        else:
            c = Ly.conv_bias(ce.conv_in, Ly._cl(controlnet_cond.to(dtype)), silu=True)     # conv -> +bias -> SiLU: one epilogue
            for blk in ce.blocks:
                c = Ly.conv_bias(blk, c, silu=True)
            # conv_in(sample) + conv_out(c): both biases and the sum in one epilogue pass
            x = Ly.conv_bias(ce.conv_out, c, residual=Ly._cl(Ly.conv_nobias(self.conv_in, Ly._cl(sample.to(dtype)))),
                             extra_bias=self.conv_in.bias)
        frames = 1  # every frame is an independent image here: per-frame GroupNorm statistics
        # the temb rows are per frame already (n rows), so the "batch" of the video view is n
        skips = [x]
        for blk in self.down_blocks:
            for j in range(len(blk.resnets)):
                x = blk.layer(j, x, emb, frames, ctx, ctx_map, next(shifts))
                skips.append(x)
            if hasattr(blk, "downsamplers"):
                x = Ly.conv_bias(blk.downsamplers[0].conv, x)
                skips.append(x)
        mb = self.mid_block
        x = mb.resnets[0].native(x, emb, frames, next(shifts))
        x = mb.attentions[0].native(x, frames, ctx, ctx_map)
        x = mb.resnets[1].native(x, emb, frames, next(shifts))
        bufs = list(out) if out is not None else [None] * (len(skips) + 1)
        res = [self._zero_conv(z, s, o) for z, s, o in zip(self.controlnet_down_blocks, skips, bufs)]
        res.append(self._zero_conv(self.controlnet_mid_block, x, bufs[-1]))
        return res
