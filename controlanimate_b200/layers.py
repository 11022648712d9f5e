"""Host-side mirror of the reference's operator surface for the denoising hot path.

Every class keeps the reference module's name-space (state_dict keys, forward signature) and routes
the arithmetic to the sm_100a kernels behind the C ABI (controlanimate_b200.ops).  Nothing here has
a CPU or eager fallback: tensors must be CUDA bf16/f16 and the library must load.

Native activation layout: a video activation with logical shape [b, c, f, h, w] is stored in
memory order b,f,h,w,c ("BFHWC").  Its 4-D view [(b f), c, h, w] is torch `channels_last` (what
cuDNN wants for 16-bit convolutions) and its 2-D view [(b f h w), c] is the token matrix that the
transformer kernels and tcgen05 GEMMs consume — so none of the reference's einops rearrange copies
(resnet.py:16-18,27-29; motion_module.py:139,146,155-159,285,327; attention.py:124,137,153-164) exist.
Modules also accept the reference's NCFHW-contiguous tensors at their public `forward` (drop-in
boundaries B2/B4) and return the layout they were given.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import ops


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
def f32(p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """fp32 copy of a (small) parameter, cached on the tensor and refreshed when it is modified in place
    or moved: kernels read gamma/beta/bias/pe as fp32 while the module may be held in bf16 (`.half()` in the
    reference, modules/controlanimate_pipeline.py:108-110)."""
    if p is None:
        return None
    if p.dtype == torch.float32 and p.is_contiguous():
        return p.detach()
    key = (p._version, p.data_ptr(), p.device)
    cached = getattr(p, "_ca_f32", None)
    if cached is None or cached[0] != key:
        cached = (key, p.detach().float().contiguous())
        try:
            p._ca_f32 = cached
        except AttributeError:  # pragma: no cover
            pass
    return cached[1]


def to_native(x: torch.Tensor) -> torch.Tensor:
    """[b,c,f,h,w] with ANY strides -> same logical tensor with BFHWC memory order (copy only if needed).  The reference
    hands its modules whatever `rearrange` produced, e.g. the "(b f) c h w -> b c f h w" VIEW behind every InflatedConv3d
    (resnet.py:16-18), which is neither NCFHW-contiguous nor BFHWC."""
    if x.dim() != 5:
        raise ValueError(f"expected a 5-D [b,c,f,h,w] tensor, got {tuple(x.shape)}")
    p = x.permute(0, 2, 3, 4, 1)
    return x if p.is_contiguous() else p.contiguous().permute(0, 4, 1, 2, 3)


def _like_input(y5: torch.Tensor, x5: torch.Tensor) -> torch.Tensor:
    """Drop-in modules return NCFHW-contiguous tensors for NCFHW-contiguous inputs (what a caller that allocated such a tensor
    may rely on) and the native BFHWC layout otherwise: every consumer in the reference is a strided torch op / rearrange,
    and `b c f h w -> (b f) c h w` of a BFHWC tensor is a free channels_last view (no copy in front of the next cuDNN conv)."""
    return y5.contiguous() if x5.is_contiguous() else y5


def frames4(x5: torch.Tensor) -> torch.Tensor:
    """BFHWC [b,c,f,h,w] -> channels_last [(b f), c, h, w] view."""
    b, c, f, h, w = x5.shape
    return x5.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)


def video5(x4: torch.Tensor, frames: int) -> torch.Tensor:
    """channels_last [(b f), c, h, w] -> BFHWC [b,c,f,h,w] view."""
    n, c, h, w = x4.shape
    return x4.reshape(n // frames, frames, c, h, w).permute(0, 2, 1, 3, 4)


def tokens(x4: torch.Tensor) -> torch.Tensor:
    """channels_last [(b f), c, h, w] -> token matrix [(b f h w), c] view."""
    n, c, h, w = x4.shape
    return x4.permute(0, 2, 3, 1).reshape(n * h * w, c)


def from_tokens(t: torch.Tensor, n: int, h: int, w: int) -> torch.Tensor:
    return t.reshape(n, h, w, t.shape[-1]).permute(0, 3, 1, 2)


def _cl(x4: torch.Tensor) -> torch.Tensor:
    return x4 if x4.is_contiguous(memory_format=torch.channels_last) else x4.contiguous(memory_format=torch.channels_last)


def group_norm(x4: torch.Tensor, norm: nn.GroupNorm, frames: int, *, per_frame: bool = True, silu: bool = False,
               temb: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Kernel (2) on a native 4-D activation; returns a channels_last 4-D tensor."""
    y5 = ops.groupnorm_silu(video5(_cl(x4), frames), f32(norm.weight), f32(norm.bias), norm.num_groups, norm.eps,
                            per_frame=per_frame, silu=silu, temb=temb)
    return frames4(y5)


def _weight_cl(conv: nn.Conv2d) -> torch.Tensor:
    """The convolution weight in channels_last order, converted ONCE per weight version.  cuDNN picks its NHWC kernels for
    channels_last activations and torch would otherwise re-lay the (NCHW-stored) filter on every call: r01c's profile
    showed 158 such copies per denoising step (2.7 ms, 3.4 % of the step)."""
    w = conv.weight
    if w.is_contiguous(memory_format=torch.channels_last):
        return w
    cache = getattr(conv, "_ca_weight_cl", None)
    if cache is None or cache[0] is not w or cache[1] != w._version or cache[2].device != w.device or cache[2].dtype != w.dtype:
        cache = (w, w._version, w.detach().contiguous(memory_format=torch.channels_last))
        conv._ca_weight_cl = cache
    return cache[2]


def conv_nobias(conv: nn.Conv2d, x4: torch.Tensor) -> torch.Tensor:
    """cuDNN convolution WITHOUT its bias: torch would add it in a separate broadcast pass; the callers fold it into the
    next fused kernel instead (GroupNorm's per-(b,c) shift, the residual epilogue, or ops.bias_act_residual)."""
    return F.conv2d(x4, _weight_cl(conv), None, conv.stride, conv.padding, conv.dilation, conv.groups)


def conv_bias(conv: nn.Conv2d, x4: torch.Tensor, *, silu: bool = False, residual: Optional[torch.Tensor] = None,
              extra_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """conv(x) + bias [-> SiLU] [+ residual]: cuDNN convolution followed by ONE in-place epilogue pass."""
    y = _cl(conv_nobias(conv, x4))
    if conv.out_channels % 8:
        y = y + conv.bias.to(y.dtype).view(1, -1, 1, 1) if conv.bias is not None else y
        y = F.silu(y) if silu else y
        return y if residual is None else y + residual
    bias = sum_f32(conv, conv.bias, extra_bias) if extra_bias is not None else f32(conv.bias)
    return ops.bias_act_residual(y, bias, residual, silu=silu, inplace=True)


def concat_channels(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """torch.cat([a, b], dim=1) of two [(b f), c, h, w] tensors; one streaming kernel when both are channels_last 16-bit."""
    a, b = _cl(a), _cl(b)
    if a.is_cuda and a.dtype == b.dtype and a.shape[1] % 8 == 0 and b.shape[1] % 8 == 0 and a.dtype != torch.float32:
        return ops.concat_channels(a, b)
    return torch.cat([a, b], dim=1)


def upsample_nearest(x: torch.Tensor, size=None) -> torch.Tensor:
    """F.interpolate(mode="nearest") by 2 or to `size` (Upsample3D.forward, resnet.py:63-69)."""
    x = _cl(x)
    if x.is_cuda and x.shape[1] % 8 == 0 and x.dtype != torch.float32:
        return ops.upsample_nearest(x, size)
    return F.interpolate(x, size=size, mode="nearest") if size is not None else F.interpolate(x, scale_factor=2.0, mode="nearest")


def sum_f32(owner: nn.Module, *ps: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """fp32 sum of small parameter vectors (folded biases), cached ON THE OWNING MODULE until one of them changes —
    the cache dies with the module, so a recycled id()/address of a later model can never hit a stale entry."""
    ps = [p for p in ps if p is not None]
    if not ps:
        return None
    key = tuple((id(p), p._version, p.data_ptr(), p.device) for p in ps)
    hit = owner.__dict__.get("_ca_bias_sum")
    if hit is None or hit[0] != key or any(a is not b for a, b in zip(hit[2], ps)):
        total = ps[0].detach().float()
        for p in ps[1:]:
            total = total + p.detach().float()
        hit = (key, total.contiguous(), tuple(ps))   # holding the params keeps their id()s unique while cached
        owner.__dict__["_ca_bias_sum"] = hit
    return hit[1]


class TembBank:
    """All time_emb_proj layers of a model evaluated as ONE GEMM per forward (they share the input silu(temb)):
    returns, per resnet, the [b, C_out] fp32 shift GroupNorm-2 adds (projection + its bias + conv1's bias) as row-strided
    views of one [b, sum C_out] buffer.  Replaces ~4 tiny launches per resnet (resnet.py:196-200)."""

    def __init__(self):
        self._key = None

    def shifts(self, resnets, temb: torch.Tensor):
        key = tuple((r.time_emb_proj.weight._version, r.time_emb_proj.weight.data_ptr(), r.time_emb_proj.bias._version,
                     r.conv1.bias._version, r.conv1.bias.data_ptr()) for r in resnets)
        if key != self._key:
            self._w = torch.cat([r.time_emb_proj.weight.detach() for r in resnets], dim=0).contiguous()
            self._b = torch.cat([r.time_emb_proj.bias.detach().float() + r.conv1.bias.detach().float() for r in resnets])
            self._tb = torch.cat([r.time_emb_proj.bias.detach() for r in resnets])
            self._cb = torch.cat([r.conv1.bias.detach().float() for r in resnets])
            self._splits = [r.out_channels for r in resnets]
            self._key = key
        # same rounding points as the per-layer evaluation: the projection (with its bias) is produced in the model
        # dtype, then widened; conv1's bias is added in fp32
        proj = F.linear(F.silu(temb), self._w, self._tb).float() + self._cb
        return list(proj.split(self._splits, dim=1))


class _FusedWeights:
    """Concatenated projection weights (to_q|to_k|to_v -> one [3C, C] GEMM), rebuilt if a source changes."""

    def __init__(self):
        self._key = None
        self._w = None

    def get(self, *ws: torch.Tensor) -> torch.Tensor:
        key = tuple((w._version, w.data_ptr(), w.dtype) for w in ws)
        if key != self._key:
            self._w = torch.cat([w.detach() for w in ws], dim=0).contiguous()
            self._key = key
        return self._w


class _LnFold:
    """A LayerNorm folded into the projection(s) behind it (ops.fold_layernorm): gain-scaled weights, their column sums
    and the shift table, rebuilt when any source tensor changes."""

    def __init__(self):
        self._key = None
        self._val = None

    def get(self, ws: Sequence[torch.Tensor], norm: nn.LayerNorm, bias: Optional[torch.Tensor] = None,
            pe: Optional[torch.Tensor] = None):
        srcs = list(ws) + [norm.weight, norm.bias] + [t for t in (bias, pe) if t is not None]
        key = tuple((t._version, t.data_ptr(), t.dtype, t.device) for t in srcs)
        if key != self._key:
            w = ws[0].detach() if len(ws) == 1 else torch.cat([w.detach() for w in ws], dim=0)
            self._val = ops.fold_layernorm(w, norm.weight, norm.bias, bias=bias, pe=pe)
            self._key = key
        return self._val


def ln_foldable(h: torch.Tensor, sites: int = 32) -> bool:
    """CA_LN_FOLD=1: LayerNorm -> projection pairs run as row statistics + one GEMM whose epilogue normalises (ca_row_stats /
    ca_linear_ln) instead of ca_layernorm_pe + ca_linear.  Parity-green, but measured a wash (profiles/r02_notes.md §2: the
    step gains 0.5 ms of 72 while the projection GEMMs, now carrying the normalisation, drop from 0.69 to 0.66 of the
    tensor roofline), so the two-launch form stays the default."""
    return _LN_FOLD and h.dtype in (torch.bfloat16, torch.float16) and h.shape[-1] % 32 == 0 and h.shape[-1] <= 1280 \
        and sites % 32 == 0


import os as _os

_LN_FOLD = _os.environ.get("CA_LN_FOLD", "0") == "1"
# CA_OWN_FMHA: the h*w x h*w self-attention core of the spatial transformers on ca_spatial_attn_core instead of torch SDPA
_OWN_FMHA = _os.environ.get("CA_OWN_FMHA", "0") == "1"

# CA_FUSED_TEMPORAL=1 routes the temporal-attention blocks of the 320-wide level through the one-launch kernel
# (ca_temporal_attn_fused).  It is parity-green but, as measured in profiles/r02_fused_notes.md, still slower than the
# four-launch block (LayerNorm+PE, QKV GEMM, attention core, out-proj GEMM: 330 vs 250 us at 64x64), so it is opt-in.
_FUSED_TEMPORAL = _os.environ.get("CA_FUSED_TEMPORAL", "0") == "1"

# --------------------------------------------------------------------------------------------------
# B1: AttentionProcessor for VersatileAttention (temporal self-attention)
# --------------------------------------------------------------------------------------------------
class B200TemporalAttnProcessor(nn.Module):
    """Drop-in AttentionProcessor (reference modules/attention_processor.py:186-272 is the semantic spec).

    `processor(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None)` with
    hidden_states [(b d), f, C] already LayerNorm'd and PE-added by VersatileAttention.forward
    (motion_module.py:285-288,321).  Computes to_out(softmax(q k^T scale) v) with: one fused tcgen05
    GEMM for to_q|to_k|to_v, the TMA/mma temporal-attention core reading the "(b d) f c" row order in
    place, and a tcgen05 GEMM with fused bias for to_out[0].  nn.Module so it can live in the
    ModuleList IP-Adapter builds (modules/ip_adapter.py:184).
    """

    def __init__(self, hidden_size=None, cross_attention_dim=None):
        super().__init__()
        self._qkv = _FusedWeights()

    @staticmethod
    def _is_self(hidden_states, encoder_hidden_states) -> bool:
        """VersatileAttention.forward (motion_module.py:309,321) never passes None: for self-attention it hands the
        processor `encoder_hidden_states = hidden_states`.  Same object, or the same memory viewed the same way, is
        self-attention; anything else is a genuine context this processor does not implement."""
        e = encoder_hidden_states
        return e is None or e is hidden_states or (
            e.data_ptr() == hidden_states.data_ptr() and e.shape == hidden_states.shape
            and e.stride() == hidden_states.stride() and e.dtype == hidden_states.dtype)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        if not self._is_self(hidden_states, encoder_hidden_states) or attention_mask is not None:
            raise ValueError("B200TemporalAttnProcessor handles temporal SELF-attention without a mask only")
        for name in ("spatial_norm", "group_norm", "norm_cross"):
            if getattr(attn, name, None) is not None:
                raise ValueError(f"attn.{name} is not supported on the temporal path")
        if getattr(attn, "residual_connection", False) or getattr(attn, "rescale_output_factor", 1.0) != 1.0:
            raise ValueError("residual_connection / rescale_output_factor are not supported on the temporal path")
        if hidden_states.dim() != 3:
            raise ValueError("expected hidden_states [(b d), f, C]")
        if attn.to_q.bias is not None:
            raise ValueError("temporal attention projections carry no bias (attention_bias=False)")
        bd, f, c = hidden_states.shape
        x = hidden_states.contiguous().reshape(bd * f, c)
        wqkv = self._qkv.get(attn.to_q.weight, attn.to_k.weight, attn.to_v.weight)
        qkv = ops.linear(x, wqkv)
        o = ops.temporal_attention_core(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], batch=1, frames=f, sites=bd,
                                        heads=attn.heads, scale=attn.scale, seq_major=True)
        out = ops.linear(o, attn.to_out[0].weight, f32(attn.to_out[0].bias))
        return out.reshape(bd, f, c)  # to_out[1] is Dropout(0) in eval


# --------------------------------------------------------------------------------------------------
# B2: motion module
# --------------------------------------------------------------------------------------------------
class _PositionalEncoding(nn.Module):
    """Holds the `pe` buffer of reference PositionalEncoding (motion_module.py:227-245)."""

    def __init__(self, d_model: int, max_len: int):
        super().__init__()
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(1, max_len, d_model)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)

    def _apply(self, fn, recurse=True):
        # follow device moves but keep the table in fp32 (the kernel adds it in fp32; `.half()`/`.bfloat16()` on the
        # model must not quantise it)
        pe = self.pe
        super()._apply(fn, recurse)
        dev = self.pe.device
        if dev.type == "meta":
            return self
        if pe.is_meta:  # materialised from the meta device (utils.build_on_device): the table is analytic
            from .utils import analytic_pe
            self.pe = analytic_pe(pe.shape[1], pe.shape[2], device=dev)
        else:
            self.pe = pe.to(device=dev, dtype=torch.float32)
        return self


class TemporalAttention(nn.Module):
    """Parameter container mirroring VersatileAttention (motion_module.py:248-270) incl. the processor protocol."""

    def __init__(self, dim: int, heads: int, max_len: Optional[int]):
        super().__init__()
        self.heads = heads
        self.scale = (dim // heads) ** -0.5
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(dim, dim, bias=False)
        self.to_v = nn.Linear(dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
        self.pos_encoder = _PositionalEncoding(dim, max_len) if max_len else None
        self.attention_mode = "Temporal"
        self.is_cross_attention = False
        self.group_norm = self.spatial_norm = self.norm_cross = None
        self.residual_connection, self.rescale_output_factor = False, 1.0
        self.processor = B200TemporalAttnProcessor()
        self._qkv = _FusedWeights()

    # processors that are plain softmax(q k^T) v self-attention: arithmetically what B200TemporalAttnProcessor computes
    _PLAIN = ("AttnProcessor", "AttnProcessor2_0", "XFormersAttnProcessor")

    def set_processor(self, processor, _remove_lora=False):
        """`set_ip_adapter` (modules/ip_adapter.py:95-126) and `enable_xformers_memory_efficient_attention`
        (controlanimate_pipeline.py:112) overwrite EVERY processor of the UNet, the temporal ones with a plain
        AttnProcessor2_0 / xformers processor.  Those compute exactly what the B200 processor computes, so the B200
        processor survives; any other (foreign) processor is honoured through the AttentionProcessor protocol."""
        if type(processor).__name__ in self._PLAIN and isinstance(getattr(self, "processor", None), B200TemporalAttnProcessor):
            return
        if isinstance(getattr(self, "processor", None), nn.Module) and not isinstance(processor, nn.Module):
            self._modules.pop("processor", None)
        self.processor = processor

    def get_processor(self, return_deprecated_lora: bool = False):
        return self.processor

    def fused_ok(self, h: torch.Tensor, f: int) -> bool:
        """The one-launch block (ca_temporal_attn_fused) applies: a built width, our own processor, 16-bit activations."""
        c = h.shape[-1]
        return (_FUSED_TEMPORAL and c in ops.FUSED_TEMPORAL_WIDTHS and isinstance(self.processor, B200TemporalAttnProcessor)
                and h.dtype in (torch.bfloat16, torch.float16) and f <= 32 and (c // self.heads) % 8 == 0 and h.is_contiguous())

    def fused(self, h: torch.Tensor, norm: nn.LayerNorm, b: int, f: int, d: int) -> torch.Tensor:
        """h + to_out(attn(LN(h) + pe)) in one launch (motion_module.py:213-219 for one attention block)."""
        key = tuple((w._version, w.data_ptr(), w.dtype) for w in (self.to_q.weight, self.to_k.weight, self.to_v.weight))
        if getattr(self, "_perm_key", None) != key:
            self._perm = ops.pack_qkv_per_head(self.to_q.weight, self.to_k.weight, self.to_v.weight, self.heads)
            self._perm_key = key
        pe = self.pos_encoder.pe if self.pos_encoder is not None else None
        return ops.temporal_attention_fused(h, f32(norm.weight), f32(norm.bias), f32(pe), self._perm, self.to_out[0].weight,
                                            f32(self.to_out[0].bias), batch=b, frames=f, sites=d, heads=self.heads, eps=norm.eps,
                                            scale=self.scale)

    def native(self, n_tok: torch.Tensor, residual: torch.Tensor, b: int, f: int, d: int) -> torch.Tensor:
        """to_out(attn(qkv(n_tok))) + residual on token-major rows (n_tok already LayerNorm'd + PE-added)."""
        c = n_tok.shape[-1]
        if not isinstance(self.processor, B200TemporalAttnProcessor):
            # a foreign processor was installed (e.g. ip_adapter.py:95-126 overwrites every processor):
            # honour the AttentionProcessor protocol in the reference's "(b d) f c" order.
            x = n_tok.reshape(b, f, d, c).permute(0, 2, 1, 3).reshape(b * d, f, c)
            o = self.processor(self, x, encoder_hidden_states=None, attention_mask=None)
            return o.reshape(b, d, f, c).permute(0, 2, 1, 3).reshape(b * f * d, c) + residual
        wqkv = self._qkv.get(self.to_q.weight, self.to_k.weight, self.to_v.weight)
        qkv = ops.linear(n_tok, wqkv)
        return self._attend(qkv, residual, b, f, d)

    def _attend(self, qkv: torch.Tensor, residual: torch.Tensor, b: int, f: int, d: int) -> torch.Tensor:
        c = residual.shape[-1]
        o = ops.temporal_attention_core(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], batch=b, frames=f, sites=d,
                                        heads=self.heads, scale=self.scale)
        return ops.linear(o, self.to_out[0].weight, f32(self.to_out[0].bias), residual=residual)

    def native_ln(self, h: torch.Tensor, norm: nn.LayerNorm, b: int, f: int, d: int) -> torch.Tensor:
        """h + to_out(attn(qkv(LayerNorm(h) + pe))) from the raw rows h: the LayerNorm (+PE) is applied by the QKV GEMM's
        epilogue (motion_module.py:214-215, 285-288 folded into :321's projections)."""
        pe = self.pos_encoder.pe if self.pos_encoder is not None else None
        if pe is not None and pe.shape[-2] < f:
            raise ValueError(f"video has {f} frames but the positional encoding only {pe.shape[-2]} (motion_module.py:236)")
        if "_ln" not in self.__dict__:
            self.__dict__["_ln"] = _LnFold()
        w_gain, colsum, shift = self.__dict__["_ln"].get((self.to_q.weight, self.to_k.weight, self.to_v.weight), norm, pe=pe)
        qkv = ops.linear_ln(h, ops.row_stats(h, norm.eps), w_gain, colsum, shift, frames=f, sites=d)
        return self._attend(qkv, h, b, f, d)


class _GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)


class FeedForward(nn.Module):
    """diffusers FeedForward(geglu) container: net.0.proj [8C, C], net.2 [C, 4C]."""

    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([_GEGLU(dim, 4 * dim), nn.Dropout(0.0), nn.Linear(4 * dim, dim)])

    def native(self, n_tok: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
        u = ops.linear(n_tok, self.net[0].proj.weight, f32(self.net[0].proj.bias), geglu=True)
        return ops.linear(u, self.net[2].weight, f32(self.net[2].bias), residual=residual)

    def native_ln(self, h: torch.Tensor, norm: nn.LayerNorm) -> torch.Tensor:
        """h + net.2(GEGLU(net.0.proj(LayerNorm(h)))) from the raw rows: the LayerNorm is applied in the GEGLU GEMM's epilogue."""
        if "_ln" not in self.__dict__:
            self.__dict__["_ln"] = _LnFold()
        w_gain, colsum, shift = self.__dict__["_ln"].get((self.net[0].proj.weight,), norm, bias=self.net[0].proj.bias)
        u = ops.linear_ln(h, ops.row_stats(h, norm.eps), w_gain, colsum, shift, geglu=True)
        return ops.linear(u, self.net[2].weight, f32(self.net[2].bias), residual=h)


class _TemporalTransformerBlock(nn.Module):
    def __init__(self, dim, heads, n_attn, max_len):
        super().__init__()
        self.attention_blocks = nn.ModuleList([TemporalAttention(dim, heads, max_len) for _ in range(n_attn)])
        self.norms = nn.ModuleList([nn.LayerNorm(dim) for _ in range(n_attn)])
        self.ff = FeedForward(dim)
        self.ff_norm = nn.LayerNorm(dim)

    def native(self, h: torch.Tensor, b: int, f: int, d: int) -> torch.Tensor:
        for attn, norm in zip(self.attention_blocks, self.norms):            # motion_module.py:213-219
            if attn.fused_ok(h, f):                                          # kernel (1) fused: one launch per attention block
                h = attn.fused(h, norm, b, f, d)
                continue
            if ln_foldable(h, d) and isinstance(attn.processor, B200TemporalAttnProcessor):
                h = attn.native_ln(h, norm, b, f, d)
                continue
            pe = attn.pos_encoder.pe if attn.pos_encoder is not None else None
            n = ops.layernorm_pe(h, f32(norm.weight), f32(norm.bias), norm.eps, pe=f32(pe), frames=f, sites=d)
            h = attn.native(n, h, b, f, d)
        if ln_foldable(h):
            return self.ff.native_ln(h, self.ff_norm)                        # :221
        n = ops.layernorm_pe(h, f32(self.ff_norm.weight), f32(self.ff_norm.bias), self.ff_norm.eps)
        return self.ff.native(n, h)                                          # :221


class _TemporalTransformer3D(nn.Module):
    def __init__(self, channels, heads, num_layers, n_attn, max_len, norm_num_groups=32):
        super().__init__()
        self.norm = nn.GroupNorm(norm_num_groups, channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(channels, channels)
        self.transformer_blocks = nn.ModuleList(
            [_TemporalTransformerBlock(channels, heads, n_attn, max_len) for _ in range(num_layers)])
        self.proj_out = nn.Linear(channels, channels)


class B200MotionModule(nn.Module):
    """Drop-in for VanillaTemporalModule (reference motion_module.py:50-84): same ctor kwargs, same
    state_dict keys (so animatediff/utils/util.py:117-120 `load_weights` works), same forward signature.

    forward(input_tensor [b,C,f,h,w], temb, encoder_hidden_states, attention_mask=None, anchor_frame_idx=None)
    """

    def __init__(self, in_channels, num_attention_heads=8, num_transformer_block=2,
                 attention_block_types=("Temporal_Self", "Temporal_Self"), cross_frame_attention_mode=None,
                 temporal_position_encoding=False, temporal_position_encoding_max_len=24,
                 temporal_attention_dim_div=1, zero_initialize=True):
        super().__init__()
        if any(t != "Temporal_Self" for t in attention_block_types):
            raise ValueError("only Temporal_Self attention blocks exist in the shipped motion modules")
        if temporal_attention_dim_div != 1:
            raise ValueError("temporal_attention_dim_div != 1 is not supported")
        self.temporal_transformer = _TemporalTransformer3D(
            in_channels, num_attention_heads, num_transformer_block, len(attention_block_types),
            temporal_position_encoding_max_len if temporal_position_encoding else None)
        if zero_initialize:
            nn.init.zeros_(self.temporal_transformer.proj_out.weight)
            nn.init.zeros_(self.temporal_transformer.proj_out.bias)

    def native(self, x4: torch.Tensor, frames: int) -> torch.Tensor:
        """x4: channels_last [(b f), C, h, w] -> same."""
        tt = self.temporal_transformer
        n, c, h, w = x4.shape
        b, d = n // frames, h * w
        x4 = _cl(x4)
        x_tok = tokens(x4)
        g = tokens(group_norm(x4, tt.norm, frames, per_frame=True, silu=False))           # motion_module.py:144
        hdn = ops.linear(g, tt.proj_in.weight, f32(tt.proj_in.bias))                       # :147
        for blk in tt.transformer_blocks:
            hdn = blk.native(hdn, b, frames, d)
        y = ops.linear(hdn, tt.proj_out.weight, f32(tt.proj_out.bias), residual=x_tok)     # :155-158
        return from_tokens(y, n, h, w)

    def forward(self, input_tensor, temb=None, encoder_hidden_states=None, attention_mask=None, anchor_frame_idx=None):
        if input_tensor.dim() != 5:
            raise ValueError(f"Expected hidden_states to have ndim=5, but got ndim={input_tensor.dim()}.")
        x5 = to_native(input_tensor)
        y5 = video5(self.native(frames4(x5), x5.shape[2]), x5.shape[2])
        return _like_input(y5, input_tensor)


# --------------------------------------------------------------------------------------------------
# B4: ResnetBlock3D
# --------------------------------------------------------------------------------------------------
class InflatedGroupNormParams(nn.GroupNorm):
    """Parameter holder (keys weight/bias) for InflatedGroupNorm / nn.GroupNorm (resnet.py:23-31)."""

    def forward(self, x):  # pragma: no cover - arithmetic lives in kernel (2)
        raise RuntimeError("GroupNorm runs inside the fused ca_groupnorm_silu kernel")


class B200ResnetBlock3D(nn.Module):
    """Drop-in for ResnetBlock3D (reference resnet.py:111-218): norm1+SiLU and (+temb) norm2+SiLU are one
    fused kernel each; the convolutions stay cuDNN (channels_last, no rearrange copies)."""

    def __init__(self, *, in_channels, out_channels=None, temb_channels=512, groups=32, groups_out=None, eps=1e-6,
                 output_scale_factor=1.0, use_in_shortcut=None, use_inflated_groupnorm=True, **unused):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.output_scale_factor = output_scale_factor
        self.per_frame = bool(use_inflated_groupnorm)
        self.norm1 = InflatedGroupNormParams(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = InflatedGroupNormParams(groups_out or groups, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        use_in_shortcut = in_channels != out_channels if use_in_shortcut is None else use_in_shortcut
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if use_in_shortcut else None

    def temb_shift(self, temb: Optional[torch.Tensor], batch: int) -> Optional[torch.Tensor]:
        """[b, C] fp32 shift GroupNorm-2 adds before its statistics: time_emb_proj(silu(temb)) (resnet.py:196-200) plus
        conv1's bias, which is constant over (f, h, w) exactly like the time embedding and therefore folds into it."""
        t = None
        if temb is not None and self.time_emb_proj is not None:
            t = self.time_emb_proj(F.silu(temb)).float()
        if self.conv1.bias is not None:
            b1 = f32(self.conv1.bias)
            t = b1.unsqueeze(0).expand(batch, -1).contiguous() if t is None else t + b1
        return t

    def native(self, x4: torch.Tensor, temb: Optional[torch.Tensor], frames: int,
               shift: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x4 channels_last [(b f), C, h, w].  `shift` = precomputed temb_shift (UNet batches all of them in one GEMM)."""
        x4 = _cl(x4)
        h = group_norm(x4, self.norm1, frames, per_frame=self.per_frame, silu=True)        # resnet.py:191-192
        h = conv_nobias(self.conv1, h)                                                     # :194 (bias -> shift)
        t = self.temb_shift(temb, x4.shape[0] // frames) if shift is None else shift
        h = group_norm(h, self.norm2, frames, per_frame=self.per_frame, silu=True, temb=t)  # :199-208 fused
        h = _cl(conv_nobias(self.conv2, h))                                                # :211 (bias -> epilogue)
        inv = 1.0 / self.output_scale_factor
        if self.conv_shortcut is not None and inv == 1.0 and self.in_channels % 8 == 0:
            # 1x1 shortcut == token GEMM whose epilogue adds both biases and conv2's output   :213-216
            n, _, hh, ww = x4.shape
            w2 = self.conv_shortcut.weight.reshape(self.out_channels, self.in_channels)
            y = ops.linear(tokens(x4), w2, sum_f32(self, self.conv_shortcut.bias, self.conv2.bias), residual=tokens(h))
            return from_tokens(y, n, hh, ww)
        if self.conv_shortcut is not None:
            x4 = conv_bias(self.conv_shortcut, x4)
        return ops.bias_act_residual(h, f32(self.conv2.bias), x4, scale=inv, inplace=True)

    def forward(self, input_tensor, temb):
        x5 = to_native(input_tensor)
        y5 = video5(self.native(frames4(x5), temb, x5.shape[2]), x5.shape[2])
        return _like_input(y5, input_tensor)


# --------------------------------------------------------------------------------------------------
# Spatial transformer (reference attention.py:52-167, 170-300).  Row N2 of SURVEY §8f: the h*w-token
# attention core is still a library call (torch SDPA); norms and every projection run on our kernels.
# --------------------------------------------------------------------------------------------------
_CTX_I32 = {}


def _ctx_i32(ctx_map: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """frame -> prompt index as the int32 device array the C ABI takes; converted once per index tensor."""
    if ctx_map is None:
        return None
    key = (ctx_map.data_ptr(), ctx_map.numel(), ctx_map._version, str(ctx_map.device))
    hit = _CTX_I32.get(key)
    if hit is None:
        if len(_CTX_I32) > 64:
            _CTX_I32.clear()
        hit = (ctx_map, ctx_map.to(torch.int32).contiguous())   # keeps ctx_map alive so the data_ptr key stays unique
        _CTX_I32[key] = hit
    return hit[1]


class B200IPAttnProcessor(nn.Module):
    """IP-Adapter dual-KV cross-attention (reference modules/attention_processor.py:367-492, `IPAttnProcessor2_0`): the
    last `num_tokens` rows of encoder_hidden_states are image tokens with their own K/V projections,
        out = to_out( softmax(q k_text^T) v_text + scale * softmax(q k_ip^T) v_ip ).
    Same ctor arguments and parameter names (`to_k_ip.weight`, `to_v_ip.weight`) as the reference class, so the loader of
    modules/ip_adapter.py:136-185 fills it; both softmaxes run on the cross-attention kernel (K/V resident in smem).
    Works as a drop-in AttentionProcessor (`__call__` protocol) and is recognised by the native spatial transformer."""

    def __init__(self, hidden_size, cross_attention_dim=None, scale=1.0, num_tokens=4):
        super().__init__()
        self.hidden_size, self.cross_attention_dim, self.scale, self.num_tokens = hidden_size, cross_attention_dim, scale, num_tokens
        self.to_k_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.to_v_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self._kv = _FusedWeights()
        self._kv_ip = _FusedWeights()

    @staticmethod
    def attend(proc, attn, q_tok, ctx, n_frames, d, ctx_map=None):
        """q_tok [(n_frames d), C] token-major, ctx [B_ctx, L, D] -> attention output [(n_frames d), C] before to_out.
        `proc` is any module with to_k_ip / to_v_ip / scale / num_tokens (this class or the reference's)."""
        c = q_tok.shape[-1]
        end = ctx.shape[1] - int(proc.num_tokens)
        if end <= 0:
            raise ValueError("encoder_hidden_states holds no text tokens in front of the image tokens")
        fuse = getattr(proc, "_kv", None) or _FusedWeights()
        fuse_ip = getattr(proc, "_kv_ip", None) or _FusedWeights()
        kv = ops.linear(ctx[:, :end].contiguous(), fuse.get(attn.to_k.weight, attn.to_v.weight))
        kv_ip = ops.linear(ctx[:, end:].contiguous(), fuse_ip.get(proc.to_k_ip.weight, proc.to_v_ip.weight))
        cmap = _ctx_i32(ctx_map)
        o = ops.cross_attention_core(q_tok, kv[:, :, :c], kv[:, :, c:], frames=n_frames, sites=d, heads=attn.heads,
                                     ctx_of_frame=cmap, scale=attn.scale)
        o_ip = ops.cross_attention_core(q_tok, kv_ip[:, :, :c], kv_ip[:, :, c:], frames=n_frames, sites=d, heads=attn.heads,
                                        ctx_of_frame=cmap, scale=attn.scale)
        return o.add_(o_ip, alpha=float(proc.scale))

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        if attention_mask is not None or encoder_hidden_states is None or hidden_states.dim() != 3:
            raise ValueError("B200IPAttnProcessor handles cross-attention on [(b f), d, C] tokens without a mask")
        for name in ("spatial_norm", "group_norm", "norm_cross"):
            if getattr(attn, name, None) is not None:
                raise ValueError(f"attn.{name} is not supported")
        n, d, c = hidden_states.shape
        x = hidden_states.contiguous().reshape(n * d, c)
        q_tok = ops.linear(x, attn.to_q.weight)
        o = self.attend(self, attn, q_tok, encoder_hidden_states.to(x.dtype), n, d)
        out = ops.linear(o, attn.to_out[0].weight, f32(attn.to_out[0].bias))
        return out.reshape(n, d, c)


class _SpatialAttention(nn.Module):
    def __init__(self, dim, heads, cross_dim=None):
        super().__init__()
        self.heads = heads
        self.scale = (dim // heads) ** -0.5
        cd = dim if cross_dim is None else cross_dim
        self.is_cross = cross_dim is not None
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(cd, dim, bias=False)
        self.to_v = nn.Linear(cd, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
        self.processor = None
        self._fused = _FusedWeights()

    def set_processor(self, processor, _remove_lora=False):
        if isinstance(getattr(self, "processor", None), nn.Module) and not isinstance(processor, nn.Module):
            self._modules.pop("processor", None)
        self.processor = processor

    def get_processor(self, return_deprecated_lora: bool = False):
        return self.processor

    def _project(self, n_tok, ws, norm):
        """The input projection(s) of this attention: from LayerNorm'd tokens (norm None), or from the raw rows with the
        LayerNorm `norm` folded into the GEMM (attention.py:271-286's norm1 / norm2 + to_q[/to_k/to_v])."""
        if norm is None:
            return ops.linear(n_tok, ws[0] if len(ws) == 1 else self._fused.get(*ws))
        if "_ln" not in self.__dict__:
            self.__dict__["_ln"] = _LnFold()
        w_gain, colsum, shift = self.__dict__["_ln"].get(ws, norm)
        return ops.linear_ln(n_tok, ops.row_stats(n_tok, norm.eps), w_gain, colsum, shift)

    def native(self, n_tok, residual, n_frames, d, ctx=None, ctx_map=None, norm=None):
        """`norm` given: n_tok are the RAW rows and the LayerNorm rides in the projection's epilogue."""
        c = n_tok.shape[-1]
        hd = c // self.heads
        if not self.is_cross:
            qkv = self._project(n_tok, (self.to_q.weight, self.to_k.weight, self.to_v.weight), norm)
            if _OWN_FMHA and hd % 8 == 0 and 16 <= hd <= 64:
                # own tcgen05 flash attention over the h*w sites (ca_spatial_attn_core), q / k / v read in place from the
                # packed projection output
                o = ops.spatial_attention_core(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], frames=n_frames, sites=d,
                                               heads=self.heads, scale=self.scale)
                return ops.linear(o, self.to_out[0].weight, f32(self.to_out[0].bias), residual=residual)
            qkv = qkv.reshape(n_frames, d, 3, self.heads, hd)
            q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
        else:
            q_tok = self._project(n_tok, (self.to_q.weight,), norm)                          # [T, C] token-major
            if hasattr(self.processor, "to_k_ip"):                                           # IP-Adapter dual-KV (config 4)
                o = B200IPAttnProcessor.attend(self.processor, self, q_tok, ctx, n_frames, d, ctx_map)
                return ops.linear(o, self.to_out[0].weight, f32(self.to_out[0].bias), residual=residual)
            kv = ops.linear(ctx, self._fused.get(self.to_k.weight, self.to_v.weight))        # [B_ctx, L, 2C], once per prompt
            if hd in (40, 80, 160) and kv.shape[1] <= 96 and (ctx_map is not None or n_frames % kv.shape[0] == 0):
                # own kernel (row N2): K/V of a (prompt, head) stay in smem, Q streams in and O out once
                o = ops.cross_attention_core(q_tok, kv[:, :, :c], kv[:, :, c:], frames=n_frames, sites=d, heads=self.heads,
                                             ctx_of_frame=_ctx_i32(ctx_map), scale=self.scale)
                return ops.linear(o, self.to_out[0].weight, f32(self.to_out[0].bias), residual=residual)
            q = q_tok.reshape(n_frames, d, self.heads, hd).transpose(1, 2)
            kv = kv.reshape(ctx.shape[0], ctx.shape[1], 2, self.heads, hd)
            kv = kv[ctx_map] if ctx_map is not None else kv
            k, v = kv[:, :, 0].transpose(1, 2), kv[:, :, 1].transpose(1, 2)
            if k.shape[0] != n_frames:
                k = k.repeat_interleave(n_frames // k.shape[0], dim=0)
                v = v.repeat_interleave(n_frames // v.shape[0], dim=0)
        o = F.scaled_dot_product_attention(q, k, v, scale=self.scale)                       # library (N2)
        o = o.transpose(1, 2).reshape(n_frames * d, c)
        return ops.linear(o, self.to_out[0].weight, f32(self.to_out[0].bias), residual=residual)


class _BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, cross_dim):
        super().__init__()
        self.attn1 = _SpatialAttention(dim, heads)
        self.norm1 = nn.LayerNorm(dim)
        self.attn2 = _SpatialAttention(dim, heads, cross_dim)
        self.norm2 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.processor = None  # the reference block subclasses Attention (attention.py:170) and so is enumerated too

    def set_processor(self, processor, _remove_lora=False):
        self.processor = processor

    def get_processor(self, return_deprecated_lora: bool = False):
        return self.processor

    def native(self, h, n_frames, d, ctx, ctx_map):
        if ln_foldable(h):
            h = self.attn1.native(h, h, n_frames, d, norm=self.norm1)                       # attention.py:268-271
            h = self.attn2.native(h, h, n_frames, d, ctx, ctx_map, norm=self.norm2)         # :273-286
            return self.ff.native_ln(h, self.norm3)                                         # :289
        ln = lambda x, m: ops.layernorm_pe(x, f32(m.weight), f32(m.bias), m.eps)  # noqa: E731
        h = self.attn1.native(ln(h, self.norm1), h, n_frames, d)                            # attention.py:268-271
        h = self.attn2.native(ln(h, self.norm2), h, n_frames, d, ctx, ctx_map)              # :273-286
        return self.ff.native(ln(h, self.norm3), h)                                         # :289


class SpatialTransformer3D(nn.Module):
    """Transformer3DModel mirror (use_linear_projection False: proj_in/out are 1x1 convs == token GEMMs)."""

    def __init__(self, heads, in_channels, cross_attention_dim, norm_num_groups=32):
        super().__init__()
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, in_channels, 1)
        self.transformer_blocks = nn.ModuleList([_BasicTransformerBlock(in_channels, heads, cross_attention_dim)])
        self.proj_out = nn.Conv2d(in_channels, in_channels, 1)

    def native(self, x4, frames, ctx, ctx_map=None):
        n, c, h, w = x4.shape
        x4 = _cl(x4)
        x_tok = tokens(x4)
        g = tokens(group_norm(x4, self.norm, frames, per_frame=True, silu=False))           # attention.py:131
        hdn = ops.linear(g, self.proj_in.weight.reshape(c, c), f32(self.proj_in.bias))       # :133
        for blk in self.transformer_blocks:
            hdn = blk.native(hdn, n, h * w, ctx, ctx_map)
        y = ops.linear(hdn, self.proj_out.weight.reshape(c, c), f32(self.proj_out.bias), residual=x_tok)  # :157-162
        return from_tokens(y, n, h, w)
