"""Tensor-level wrappers over the C ABI: torch is used for device memory and streams only.

Every function validates what the kernels assume, passes raw `data_ptr()`s and the current CUDA
stream, and raises (ValueError for contract violations, RuntimeError for CUDA/library failures).
There is no CPU or eager fallback: a non-CUDA tensor is an error.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import profiler as P

_DTYPES = {torch.bfloat16: L.CA_BF16, torch.float16: L.CA_F16, torch.float32: L.CA_F32}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise ValueError(f"unsupported dtype {t.dtype}") from None


def _cuda(*ts: Optional[torch.Tensor]):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("controlanimate_b200 kernels need CUDA tensors (there is no CPU fallback)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32(t: Optional[torch.Tensor], n: Optional[int] = None) -> Optional[torch.Tensor]:
    if t is None:
        return None
    t = t.detach()
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_ws_cache = {}


def _workspace(nbytes: int, device) -> Optional[torch.Tensor]:
    if nbytes == 0:
        return None
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def video_layout(x: torch.Tensor) -> int:
    """Classify a [b,c,f,h,w] tensor: NCFHW-contiguous or BFHWC (native) strides."""
    if x.dim() != 5:
        raise ValueError(f"expected a 5-D [b,c,f,h,w] tensor, got {tuple(x.shape)}")
    if x.is_contiguous():
        return L.CA_LAYOUT_NCFHW
    if x.permute(0, 2, 3, 4, 1).is_contiguous():
        return L.CA_LAYOUT_BFHWC
    raise ValueError("video tensor must be NCFHW-contiguous or BFHWC (b,f,h,w,c memory order)")


def groupnorm_silu(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float, *,
                   per_frame: bool = True, silu: bool = True, temb: Optional[torch.Tensor] = None,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = SiLU(GroupNorm(x + temb[:, :, None, None, None])) for x [b,c,f,h,w] in either layout (kernel 2)."""
    _cuda(x, gamma, beta, temb)
    layout = video_layout(x)
    b, c, f, h, w = x.shape
    y = torch.empty_like(x) if out is None else out  # empty_like preserves the (dense) strides
    if y.shape != x.shape or y.stride() != x.stride() or y.dtype != x.dtype:
        raise ValueError("out must match x in shape, strides and dtype")
    g32, b32 = _f32(gamma), _f32(beta)
    t32 = temb
    if t32 is not None and (t32.dtype != torch.float32 or t32.dim() != 2 or t32.stride(1) != 1):
        t32 = t32.detach().float().contiguous()   # row-strided fp32 [b, c] views are passed through as they are
    if g32.numel() != c or b32.numel() != c or (t32 is not None and tuple(t32.shape) != (b, c)):
        raise ValueError("gamma/beta must be [c] and temb [b, c]")
    lib = L.load()
    nws = lib.ca_groupnorm_workspace_bytes(b, c, f, h, w, groups, int(per_frame), layout, _dt(x))
    ws = _workspace(nws, x.device)
    with P.span("groupnorm_silu", 1, 2.0 * x.numel() * x.element_size()):
        L.check(lib.ca_groupnorm_silu(x.data_ptr(), y.data_ptr(), g32.data_ptr(), b32.data_ptr(), _ptr(t32),
                                      0 if t32 is None else t32.stride(0), b, c, f, h, w,
                                      groups, float(eps), int(per_frame), int(silu), layout, _dt(x), _ptr(ws),
                                      0 if ws is None else ws.numel(), _stream()), "ca_groupnorm_silu")
    return y


def layernorm_pe(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, *,
                 pe: Optional[torch.Tensor] = None, frames: int = 1, sites: int = 1,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm over the last dim of token-major x [..., c] (+ pe[frame(t)] with t = (b*f+frame)*sites+site)."""
    _cuda(x, gamma, beta, pe)
    if not x.is_contiguous():
        raise ValueError("layernorm_pe: x must be contiguous token-major")
    c = x.shape[-1]
    rows = x.numel() // c
    y = torch.empty_like(x) if out is None else out
    g32, b32, pe32 = _f32(gamma), _f32(beta), _f32(pe)
    if pe32 is not None:
        pe32 = pe32.reshape(-1, c)
        if pe32.shape[0] < frames:
            raise ValueError(f"video has {frames} frames but the positional encoding only {pe32.shape[0]} (motion_module.py:236)")
        if rows % (frames * sites) != 0:
            raise ValueError("rows must be b*frames*sites")
    with P.span("layernorm_pe", 1, 2.0 * x.numel() * x.element_size()):
        L.check(L.load().ca_layernorm_pe(x.data_ptr(), y.data_ptr(), g32.data_ptr(), b32.data_ptr(), _ptr(pe32), rows, c,
                                         frames, sites, float(eps), _dt(x), _stream()), "ca_layernorm_pe")
    return y


def temporal_attention_core(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, batch: int, frames: int, sites: int,
                            heads: int, scale: Optional[float] = None, seq_major: bool = False,
                            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T scale) v over the frame axis for q/k/v [(b f d), C] row matrices (kernel 1a).

    Rows are token-major (t = (b*f+frame)*d+site) or, with seq_major, the reference's "(b d) f c" order.
    q/k/v may be column slices of one packed [T, 3C] buffer (row stride 3C).
    """
    _cuda(q, k, v)
    T, Cq = q.shape
    if T != batch * frames * sites or k.shape != q.shape or v.shape != q.shape:
        raise ValueError("q/k/v must be [(b*f*d), C] with identical shapes")
    if Cq % heads:
        raise ValueError("C must be divisible by heads")
    for t in (q, k, v):
        if t.stride(1) != 1:
            raise ValueError("q/k/v rows must be dense")
    hd = Cq // heads
    o = torch.empty((T, Cq), dtype=q.dtype, device=q.device) if out is None else out
    if o.stride(1) != 1 or o.shape != q.shape:
        raise ValueError("bad out tensor")
    scale = hd ** -0.5 if scale is None else scale
    with P.span("temporal_attn_core", 1, 4.0 * T * Cq * q.element_size(), 4.0 * frames * Cq * T):
        L.check(L.load().ca_temporal_attn_core(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), batch, frames, sites,
                                               heads, hd, q.stride(0), k.stride(0), v.stride(0), o.stride(0), int(seq_major),
                                               float(scale), _dt(q), _stream()), "ca_temporal_attn_core")
    return o


FUSED_TEMPORAL_WIDTHS = (64, 128, 320)   # built variants of ca_temporal_attn_fused


def pack_qkv_per_head(wq: torch.Tensor, wk: torch.Tensor, wv: torch.Tensor, heads: int) -> torch.Tensor:
    """[heads * nq, C] weight for ca_temporal_attn_fused: per head the hd rows of to_q, to_k, to_v, zero-padded to nq rows."""
    c = wq.shape[0]
    hd = c // heads
    nq = (3 * hd + 15) // 16 * 16
    out = torch.zeros((heads, nq, wq.shape[1]), dtype=wq.dtype, device=wq.device)
    for i, w in enumerate((wq, wk, wv)):
        out[:, i * hd:(i + 1) * hd] = w.detach().reshape(heads, hd, -1)
    return out.reshape(heads * nq, -1).contiguous()


def temporal_attention_fused(x: torch.Tensor, ln_gamma: torch.Tensor, ln_beta: torch.Tensor, pe: Optional[torch.Tensor],
                             wqkv_perm: torch.Tensor, wo: torch.Tensor, bo: torch.Tensor, *, batch: int, frames: int, sites: int,
                             heads: int, eps: float = 1e-5, scale: Optional[float] = None,
                             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x + to_out(attn(LN(x) + pe)) for token-major x [(b f d), C] in one launch (kernel 1, fused).
    wqkv_perm = pack_qkv_per_head(to_q.weight, to_k.weight, to_v.weight, heads)."""
    _cuda(x, ln_gamma, ln_beta, pe, wqkv_perm, wo, bo)
    if x.dim() != 2 or not x.is_contiguous() or x.shape[0] != batch * frames * sites:
        raise ValueError("x must be a contiguous token matrix [(b f d), C]")
    c = x.shape[1]
    if x.dtype not in (torch.bfloat16, torch.float16) or wqkv_perm.dtype != x.dtype or wo.dtype != x.dtype:
        raise ValueError("temporal_attention_fused: x, wqkv_perm and wo must share bf16 or f16")
    hd = c // heads
    nq = (3 * hd + 15) // 16 * 16
    if tuple(wqkv_perm.shape) != (heads * nq, c) or tuple(wo.shape) != (c, c) or not wqkv_perm.is_contiguous() or not wo.is_contiguous():
        raise ValueError("temporal_attention_fused: bad weight shapes")
    g32, b32, bo32, pe32 = _f32(ln_gamma), _f32(ln_beta), _f32(bo), _f32(pe)
    if pe32 is not None:
        pe32 = pe32.reshape(-1, c)
        if pe32.shape[0] < frames:
            raise ValueError(f"video has {frames} frames but the positional encoding only {pe32.shape[0]} (motion_module.py:236)")
    y = torch.empty_like(x) if out is None else out
    scale = hd ** -0.5 if scale is None else scale
    T = x.shape[0]
    with P.span("temporal_attn_fused", 1, 2.0 * x.numel() * x.element_size() + 8.0 * c * c, T * (8.0 * c * c + 4.0 * frames * c)):
        L.check(L.load().ca_temporal_attn_fused(x.data_ptr(), y.data_ptr(), g32.data_ptr(), b32.data_ptr(), _ptr(pe32),
                                                wqkv_perm.data_ptr(), wo.data_ptr(), bo32.data_ptr(), batch, frames, sites, c, heads,
                                                float(eps), float(scale), _dt(x), _stream()), "ca_temporal_attn_fused")
    return y


def cross_attention_core(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, frames: int, sites: int, heads: int,
                         ctx_of_frame: Optional[torch.Tensor] = None, scale: Optional[float] = None,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T scale) v of every latent site against the prompt tokens (row N2, cross-attention of the spatial transformer).

    q [(frames*sites), C] token-major; k, v [n_ctx, L, C] (may be column slices of one fused [n_ctx, L, 2C] projection);
    ctx_of_frame int32 [frames] or None (frame n -> prompt n // (frames // n_ctx)).  Raises ValueError for head sizes / prompt
    lengths outside the built variants (the caller decides what to do then)."""
    _cuda(q, k, v)
    T, Cq = q.shape
    if T != frames * sites or k.dim() != 3 or k.shape != v.shape or k.shape[2] != Cq or Cq % heads:
        raise ValueError("q must be [(frames*sites), C] and k/v [n_ctx, L, C]")
    if q.stride(1) != 1 or k.stride(2) != 1 or v.stride(2) != 1:
        raise ValueError("q/k/v rows must be dense")
    if ctx_of_frame is not None and (ctx_of_frame.dtype != torch.int32 or ctx_of_frame.numel() != frames or not ctx_of_frame.is_cuda):
        raise ValueError("ctx_of_frame must be a CUDA int32 tensor with one entry per frame")
    hd = Cq // heads
    o = torch.empty((T, Cq), dtype=q.dtype, device=q.device) if out is None else out
    if o.stride(1) != 1 or o.shape != q.shape:
        raise ValueError("bad out tensor")
    scale = hd ** -0.5 if scale is None else scale
    n_ctx, L_kv = k.shape[0], k.shape[1]
    with P.span("cross_attn_core", 1, 2.0 * T * Cq * q.element_size(), 4.0 * L_kv * Cq * T):
        L.check(L.load().ca_cross_attn_core(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), frames, sites, heads, hd, n_ctx,
                                            L_kv, q.stride(0), k.stride(1), v.stride(1), o.stride(0), k.stride(0), v.stride(0),
                                            _ptr(ctx_of_frame), float(scale), _dt(q), _stream()), "ca_cross_attn_core")
    return o


def residual_merge(per_net: Sequence[Sequence[torch.Tensor]], scales: Sequence[Sequence[float]], dst: Sequence[torch.Tensor], *,
                   frames: int, add_into_dst: bool, layout: int) -> None:
    """dst_i (=|+=) Σ_k scales[k][i] * per_net[k][i] in one launch (kernel 3).

    layout NCFHW: per_net[k][i] is [(b f), c, h, w] contiguous, dst_i [b, c, f, h, w] contiguous.
    layout BFHWC: per_net[k][i] is [(b f), c, h, w] channels_last, dst_i [b,c,f,h,w] with BFHWC strides
                  (or a 4-D channels_last [(b f), c, h, w] tensor).
    """
    n_nets, n_res = len(per_net), len(dst)
    if n_nets < 1 or n_nets > L.CA_MAX_NETS or n_res > L.CA_MAX_RESIDUALS:
        raise ValueError("1..8 ControlNets and at most 16 residuals are supported")
    if any(len(r) != n_res for r in per_net) or len(scales) != n_nets or any(len(s) != n_res for s in scales):
        raise ValueError("per_net / scales / dst disagree on the number of residuals")
    _cuda(*dst, *[t for r in per_net for t in r])
    res_ptrs = (C.c_void_p * (n_nets * n_res))()
    sc = (C.c_float * (n_nets * n_res))()
    dst_ptrs = (C.c_void_p * n_res)()
    chw = (C.c_int * (3 * n_res))()
    b_res = b_dst = None
    dtype = dst[0].dtype
    for i, d in enumerate(dst):
        if d.dtype != dtype:
            raise ValueError("all tensors must share one dtype")
        if d.dim() == 5:
            bb, c, f, h, w = d.shape
            ok = d.is_contiguous() if layout == L.CA_LAYOUT_NCFHW else d.permute(0, 2, 3, 4, 1).is_contiguous()
        else:
            n, c, h, w = d.shape
            bb, f = n // frames, frames
            ok = layout == L.CA_LAYOUT_BFHWC and d.permute(0, 2, 3, 1).is_contiguous()
        if not ok or f != frames:
            raise ValueError(f"dst[{i}] has the wrong layout/frames for layout={layout}")
        b_dst = bb if b_dst is None else b_dst
        if bb != b_dst:
            raise ValueError("dst batch sizes differ")
        dst_ptrs[i] = d.data_ptr()
        chw[3 * i], chw[3 * i + 1], chw[3 * i + 2] = c, h, w
        for k in range(n_nets):
            r = per_net[k][i]
            if r.dtype != dtype or r.dim() != 4 or tuple(r.shape[1:]) != (c, h, w) or r.shape[0] % frames:
                raise ValueError(f"residual [{k}][{i}] must be [(b f), {c}, {h}, {w}] {dtype}, got {tuple(r.shape)} {r.dtype}")
            ok = r.is_contiguous() if layout == L.CA_LAYOUT_NCFHW else r.permute(0, 2, 3, 1).is_contiguous()
            if not ok:
                raise ValueError(f"residual [{k}][{i}] has the wrong memory layout")
            br = r.shape[0] // frames
            b_res = br if b_res is None else b_res
            if br != b_res:
                raise ValueError("residual batch sizes differ")
            res_ptrs[k * n_res + i] = r.data_ptr()
            sc[k * n_res + i] = float(scales[k][i])
    nbytes = sum((n_nets * (d.numel() // b_dst) * b_res + (2 if add_into_dst else 1) * d.numel()) * d.element_size() for d in dst)
    with P.span("residual_merge", 1, float(nbytes)):
        L.check(L.load().ca_residual_merge(res_ptrs, sc, dst_ptrs, chw, n_nets, n_res, b_res, b_dst, frames, int(add_into_dst),
                                           layout, _DTYPES[dtype], _stream()), "ca_residual_merge")


def _rows_view(t: torch.Tensor, what: str) -> Tuple[int, int]:
    """(rows, c) of a channels_last 4-D [(b f), c, h, w] activation or a dense [rows, c] token matrix."""
    if t.dim() == 4:
        if not t.permute(0, 2, 3, 1).is_contiguous():
            raise ValueError(f"{what} must be channels_last")
        return t.shape[0] * t.shape[2] * t.shape[3], t.shape[1]
    if t.dim() == 2 and t.is_contiguous():
        return t.shape[0], t.shape[1]
    raise ValueError(f"{what} must be a channels_last 4-D activation or a contiguous [rows, c] matrix")


def bias_act_residual(x: torch.Tensor, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, *,
                      silu: bool = False, scale: float = 1.0, inplace: bool = False) -> torch.Tensor:
    """y = (act(x + bias[c]) + residual) * scale on channels-last rows (convolution epilogue, ca_bias_act_residual)."""
    _cuda(x, bias, residual)
    rows, c = _rows_view(x, "x")
    if residual is not None and (residual.shape != x.shape or residual.dtype != x.dtype or _rows_view(residual, "residual") != (rows, c)):
        raise ValueError("residual must match x in shape, dtype and layout")
    b32 = _f32(bias)
    if b32 is not None and b32.numel() != c:
        raise ValueError("bias must be [c]")
    y = x if inplace else torch.empty_like(x)
    nbytes = x.numel() * x.element_size() * (3 if residual is not None else 2)
    with P.span("bias_act_residual", 1, float(nbytes)):
        L.check(L.load().ca_bias_act_residual(x.data_ptr(), _ptr(b32), _ptr(residual), y.data_ptr(), rows, c, float(scale),
                                              int(silu), _dt(x), _stream()), "ca_bias_act_residual")
    return y


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, residual: Optional[torch.Tensor] = None,
           geglu: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x @ w^T + bias [-> a*gelu(g)] [+ residual] on tcgen05 tensor cores (ca_linear).

    x [..., k] (dense rows, arbitrary row stride for 2-D views), w [n, k] contiguous (nn.Linear layout), bias [n].
    """
    _cuda(x, w, bias, residual)
    if x.dtype != w.dtype or x.dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("linear: x and w must both be bf16 or f16")
    k = x.shape[-1]
    n = w.shape[0]
    if w.dim() != 2 or w.shape[1] != k or not w.is_contiguous():
        raise ValueError("linear: w must be a contiguous [n, k] matrix")
    lead = x.shape[:-1]
    x2 = x if x.dim() == 2 else x.reshape(-1, k)
    if x2.stride(1) != 1:
        raise ValueError("linear: x rows must be dense")
    m = x2.shape[0]
    n_out = n // 2 if geglu else n
    y = torch.empty((m, n_out), dtype=x.dtype, device=x.device) if out is None else out.reshape(m, n_out)
    r2 = None
    if residual is not None:
        r2 = residual.reshape(m, n_out)
        if r2.stride(1) != 1 or r2.dtype != x.dtype:
            raise ValueError("linear: bad residual")
    b32 = _f32(bias)
    if b32 is not None and b32.numel() != n:
        raise ValueError("linear: bias must be [n]")
    es = x.element_size()
    nbytes = (m * k + n * k + m * n_out * (2 if r2 is not None else 1)) * es
    with P.span("linear_tcgen05", 1, float(nbytes), 2.0 * m * n * k):
        L.check(L.load().ca_linear(x2.data_ptr(), w.data_ptr(), _ptr(b32), _ptr(r2), y.data_ptr(), m, n, k, x2.stride(0),
                                   0 if r2 is None else r2.stride(0), y.stride(0), L.CA_EPI_GEGLU if geglu else L.CA_EPI_NONE,
                                   _dt(x), _stream()), "ca_linear")
    return y.reshape(*lead, n_out)


def row_stats(x: torch.Tensor, eps: float = 1e-5, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(mean, rstd) of every row of the token matrix x [T, c] as fp32 pairs [T, 2] (ca_row_stats): the statistics half of a
    LayerNorm whose normalisation is applied by `linear_ln`."""
    _cuda(x)
    if x.dim() != 2 or x.stride(1) != 1 or x.dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("row_stats: x must be a bf16/f16 token matrix with dense rows")
    T, c = x.shape
    st = torch.empty((T, 2), dtype=torch.float32, device=x.device) if out is None else out
    with P.span("row_stats", 1, float(T * c * x.element_size() + 8 * T)):
        L.check(L.load().ca_row_stats(x.data_ptr(), st.data_ptr(), T, c, x.stride(0), float(eps), _dt(x), _stream()), "ca_row_stats")
    return st


def fold_layernorm(w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias: Optional[torch.Tensor] = None,
                   pe: Optional[torch.Tensor] = None):
    """Fold LayerNorm(gamma, beta) (+ a positional-encoding table pe [R, k]) into the projection w [n, k] (+ bias) behind it:
    returns (w_gain [n, k] in w.dtype, colsum [n] fp32, shift [R or 1, n] fp32) for `linear_ln` — see ca_linear_ln."""
    wf = w.detach().float()
    w_gain = (wf * gamma.detach().float()[None, :]).to(w.dtype).contiguous()
    colsum = w_gain.float().sum(dim=1).contiguous()
    rows = beta.detach().float()[None, :]
    if pe is not None:
        rows = rows + pe.detach().float().reshape(-1, w.shape[1])
    shift = rows @ wf.t()
    if bias is not None:
        shift = shift + bias.detach().float()[None, :]
    return w_gain, colsum, shift.contiguous()


def linear_ln(x: torch.Tensor, stats: torch.Tensor, w_gain: torch.Tensor, colsum: torch.Tensor, shift: torch.Tensor, *,
              frames: int = 1, sites: int = 1, geglu: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Linear(LayerNorm(x) + pe[frame]) [-> a*gelu(g)] from the RAW rows x [T, k]: the tcgen05 GEMM runs on x with the
    gain-folded weights and its epilogue applies the rows' (mean, rstd) and the per-frame shift (ca_linear_ln)."""
    _cuda(x, stats, w_gain, colsum, shift)
    if x.dim() != 2 or x.stride(1) != 1 or x.dtype != w_gain.dtype or x.dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("linear_ln: x must be a bf16/f16 token matrix of the weights' dtype")
    m, k = x.shape
    n = w_gain.shape[0]
    if tuple(w_gain.shape) != (n, k) or not w_gain.is_contiguous():
        raise ValueError("linear_ln: w_gain must be a contiguous [n, k] matrix")
    if tuple(stats.shape) != (m, 2) or stats.dtype != torch.float32 or not stats.is_contiguous():
        raise ValueError("linear_ln: stats must be row_stats(x)")
    if colsum.dtype != torch.float32 or colsum.numel() != n or shift.dtype != torch.float32 or shift.dim() != 2 \
            or shift.shape[1] != n or not shift.is_contiguous() or not colsum.is_contiguous():
        raise ValueError("linear_ln: colsum [n] / shift [R, n] must be contiguous fp32 (fold_layernorm)")
    n_out = n // 2 if geglu else n
    y = torch.empty((m, n_out), dtype=x.dtype, device=x.device) if out is None else out.reshape(m, n_out)
    nbytes = (m * k + n * k + m * n_out) * x.element_size() + 8 * m
    with P.span("linear_tcgen05", 1, float(nbytes), 2.0 * m * n * k):
        L.check(L.load().ca_linear_ln(x.data_ptr(), w_gain.data_ptr(), colsum.data_ptr(), shift.data_ptr(), shift.shape[0],
                                      frames, sites, stats.data_ptr(), y.data_ptr(), m, n, k, x.stride(0), y.stride(0),
                                      L.CA_EPI_GEGLU if geglu else L.CA_EPI_NONE, _dt(x), _stream()), "ca_linear_ln")
    return y


def _is_cl(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)


def upsample_nearest(x: torch.Tensor, size: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """F.interpolate(x, scale_factor=2.0 | size=size, mode="nearest") for a channels_last [n, c, h, w] tensor
    (Upsample3D.forward, resnet.py:63-69) as one streaming copy (ca_upsample_nearest)."""
    _cuda(x)
    if not _is_cl(x):
        raise ValueError("upsample_nearest: x must be a channels_last [n, c, h, w] tensor")
    n, c, h, w = x.shape
    oh, ow = (2 * h, 2 * w) if size is None else (int(size[0]), int(size[1]))
    y = torch.empty((n, c, oh, ow), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    with P.span("copy_ops", 1, float((x.numel() + y.numel()) * x.element_size())):
        L.check(L.load().ca_upsample_nearest(x.data_ptr(), y.data_ptr(), n, c, h, w, oh, ow, 1 if size is None else 0, _dt(x),
                                             _stream()), "ca_upsample_nearest")
    return y


def concat_channels(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """torch.cat([a, b], dim=1) for channels_last [n, c, h, w] tensors (the skip concat of the up blocks,
    unet_blocks.py:636, :742) as one streaming copy (ca_concat_channels)."""
    _cuda(a, b)
    if not (_is_cl(a) and _is_cl(b)) or a.dtype != b.dtype or a.shape[0] != b.shape[0] or a.shape[2:] != b.shape[2:]:
        raise ValueError("concat_channels: a and b must be channels_last [n, c, h, w] tensors of one dtype and spatial size")
    n, ca, h, w = a.shape
    cb = b.shape[1]
    y = torch.empty((n, ca + cb, h, w), dtype=a.dtype, device=a.device, memory_format=torch.channels_last)
    with P.span("copy_ops", 1, float(2 * y.numel() * y.element_size())):
        L.check(L.load().ca_concat_channels(a.data_ptr(), b.data_ptr(), y.data_ptr(), n * h * w, ca, cb, _dt(a), _stream()),
                "ca_concat_channels")
    return y


def cfg_ddim_step(model_out: torch.Tensor, latents: torch.Tensor, guidance_scale: Optional[float],
                  coefficients: Sequence[float], *, out: Optional[torch.Tensor] = None,
                  noise_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Guidance combine + DDIM update of one denoising step in one launch (ca_cfg_ddim_step; reference
    controlanimation_pipeline.py:841-849).  model_out: the UNet output, [2, ...] = [uncond, cond] rows when guidance_scale
    is given, else [1, ...]; latents [1, ...] (any of bf16 / f16 / f32); coefficients = DDIMScheduler.coefficients(t) =
    (sqrt a_t, sqrt(1 - a_t), sqrt a_prev, sqrt(1 - a_prev)).  Returns the new latents (`out`, which may be `latents`)."""
    _cuda(model_out, latents, out, noise_out)
    cfg = guidance_scale is not None
    n = latents.numel()
    if model_out.numel() != (2 if cfg else 1) * n or model_out.shape[1:] != latents.shape[1:]:
        raise ValueError(f"cfg_ddim_step: model_out {tuple(model_out.shape)} does not match latents {tuple(latents.shape)} "
                         f"({'two CFG rows' if cfg else 'one row'} expected)")
    if not (model_out.is_contiguous() and latents.is_contiguous()):
        raise ValueError("cfg_ddim_step: model_out and latents must be contiguous")
    if out is None:
        out = torch.empty_like(latents)
    for o in (out, noise_out):
        if o is not None and (o.shape != latents.shape or o.dtype != latents.dtype or not o.is_contiguous()):
            raise ValueError("cfg_ddim_step: out / noise_out must look like latents")
    sa, s1a, sp, s1p = (float(c) for c in coefficients)
    with P.span("loop_update", 1, float(model_out.numel() * model_out.element_size() + 2 * n * latents.element_size())):
        L.check(L.load().ca_cfg_ddim_step(model_out.data_ptr(), latents.data_ptr(), out.data_ptr(), _ptr(noise_out), n,
                                          int(cfg), float(guidance_scale) if cfg else 0.0, sa, s1a, sp, s1p, _dt(model_out),
                                          _dt(latents), _stream()), "ca_cfg_ddim_step")
    return out


def spatial_attention_core(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, frames: int, sites: int, heads: int,
                           scale: Optional[float] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T scale) v over the sites of each frame, per head (ca_spatial_attn_core).  q / k / v: [frames * sites, C]
    row matrices (dense rows, any row stride: e.g. column slices of the packed [T, 3C] projection output)."""
    _cuda(q, k, v)
    T, c = q.shape
    if T != frames * sites or k.shape != q.shape or v.shape != q.shape or c % heads:
        raise ValueError("spatial_attention_core: q, k, v must be [frames * sites, heads * head_dim]")
    if q.dtype not in (torch.bfloat16, torch.float16) or k.dtype != q.dtype or v.dtype != q.dtype:
        raise ValueError("spatial_attention_core: bf16 or f16 only")
    if q.stride(1) != 1 or k.stride(1) != 1 or v.stride(1) != 1:
        raise ValueError("spatial_attention_core: rows must be dense")
    hd = c // heads
    o = torch.empty((T, c), dtype=q.dtype, device=q.device) if out is None else out
    scale = hd ** -0.5 if scale is None else scale
    with P.span("spatial_attn_core", 1, 4.0 * T * c * q.element_size(), 4.0 * frames * heads * float(sites) * sites * hd):
        L.check(L.load().ca_spatial_attn_core(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), frames, sites, heads, hd,
                                              q.stride(0), k.stride(0), v.stride(0), o.stride(0), float(scale), _dt(q), _stream()),
                "ca_spatial_attn_core")
    return o
