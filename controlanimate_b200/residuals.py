"""Boundary B3: the ControlNet residual contract.

Two ways to honour `down_block_additional_residuals` (12) / `mid_block_additional_residual`:

* `merge_controlnet_residuals(...)` — contract-preserving producer: reads the N per-net RAW residual lists and
  writes ONE merged set directly in the reference's [b, c, f, h, w] layout (absorbs the per-net conditioning
  scale, the guess-mode log-spaced factors, the sum over nets and the 13 '(b f) c h w -> b c f h w' rearranges of
  modules/controlresiduals_pipeline.py:294-312): (N+1)·E·s bytes of HBM traffic.
* `ResidualSet` — lazy set (per-net raw tensors + scales) that `UNet3DConditionModel.forward` folds straight into
  its skip tensors in place, `skip_i += Σ_k s_k r_{k,i}` (also replaces unet.py:567-576, 584-585): (N+2)·E·s bytes,
  nothing materialised.  Plain tuples of merged tensors are still accepted by the UNet.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import ops

N_RESIDUALS = 13


def guess_mode_factors(n: int = N_RESIDUALS) -> List[float]:
    """diffusers ControlNetModel guess-mode scaling: logspace(-1, 0, n), mid residual takes the last factor."""
    return [float(v) for v in torch.logspace(-1, 0, n)]


def _scales(cond_scale: Sequence[float], n_res: int, guess_mode: bool) -> List[List[float]]:
    level = guess_mode_factors(n_res) if guess_mode else [1.0] * n_res
    return [[level[i] * float(s) for i in range(n_res)] for s in cond_scale]


class ResidualSet:
    """Per-net raw residuals [(b f), c_i, h_i, w_i] (12 down + mid, in UNet skip order) + per-net scales."""

    def __init__(self, per_net: Sequence[Sequence[torch.Tensor]], cond_scale: Sequence[float], frames: int,
                 guess_mode: bool = False):
        if len(per_net) != len(cond_scale) or len(per_net) == 0:
            raise ValueError("one conditioning scale per ControlNet is required")
        self.per_net = [list(r) for r in per_net]
        self.n_res = len(self.per_net[0])
        self.frames = frames
        self.scales = _scales(cond_scale, self.n_res, guess_mode)
        # native nets emit channels_last tensors; diffusers-style nets emit NCHW-contiguous ones
        t = self.per_net[0][1] if self.n_res > 1 else self.per_net[0][0]
        self.layout = L.CA_LAYOUT_BFHWC if t.is_contiguous(memory_format=torch.channels_last) else L.CA_LAYOUT_NCFHW
        self._pending = None

    def produced_on(self, events, keep=()) -> "ResidualSet":
        """The per-net tensors are still being written on other streams (MultiControlNetResiduals.overlap): `events[k]` is
        recorded behind net k's last kernel; the first consumer makes ITS stream wait for them.  `keep` holds the producers'
        inputs alive until then (they were allocated on the consumer's stream)."""
        self._pending = (list(events), keep)
        return self

    def _join(self) -> None:
        if self._pending is not None:
            cur = torch.cuda.current_stream()
            for ev in self._pending[0]:
                cur.wait_event(ev)
            self._pending = None

    @staticmethod
    def coerce(down, mid, frames) -> Optional["ResidualSet"]:
        """Accept what the reference loop passes (controlanimation_pipeline.py:836-841): None, a ResidualSet in the
        `down` slot, or plain merged [b,c,f,h,w] tuples."""
        if down is None and mid is None:
            return None
        if isinstance(down, ResidualSet):
            return down
        return _MergedResiduals(list(down), mid, frames)

    def add_into(self, skips: Optional[Sequence[torch.Tensor]], mid: Optional[torch.Tensor]) -> None:
        """skips: 12 channels_last [(b f), c, h, w] tensors (in place); mid: the mid-block output (in place)."""
        self._join()
        if self.layout == L.CA_LAYOUT_NCFHW:
            # diffusers-style NCHW residuals against native (BFHWC) skips: merge once into the reference layout
            # (one kernel, contract-preserving) and add that; the single-pass path needs native residuals.
            if getattr(self, "_merged", None) is None:
                self._merged = _MergedResiduals(*self._materialize_lists(), self.frames)
            return self._merged.add_into(skips, mid)
        if skips is not None:
            k = len(skips)
            ops.residual_merge([r[:k] for r in self.per_net], [s[:k] for s in self.scales], list(skips),
                               frames=self.frames, add_into_dst=True, layout=self.layout)
        if mid is not None:
            ops.residual_merge([r[-1:] for r in self.per_net], [s[-1:] for s in self.scales], [mid],
                               frames=self.frames, add_into_dst=True, layout=self.layout)

    def added_to(self, skip: torch.Tensor, index: int) -> torch.Tensor:
        """`skip + Σ_k s_k r_{k,index}` as a NEW tensor in the skip's layout: what the reference's UNet computes at
        unet.py:572 / :585 when it adds entry `index` of the residual tuple (the B3 drop-in route: see LazyResidual)."""
        self._join()
        if skip.is_contiguous() or skip.permute(0, 2, 3, 4, 1).is_contiguous():
            out = skip.clone(memory_format=torch.preserve_format)
        else:                                   # a rearrange view of the reference: densify in the native layout
            out = skip.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
        layout = ops.video_layout(out)
        srcs = [[r[index]] for r in self.per_net]
        want_cl = layout == L.CA_LAYOUT_BFHWC
        srcs = [[t if t.is_contiguous(memory_format=torch.channels_last) == want_cl and (want_cl or t.is_contiguous())
                 else t.contiguous(memory_format=torch.channels_last if want_cl else torch.contiguous_format) for t in row] for row in srcs]
        ops.residual_merge(srcs, [[s[index]] for s in self.scales], [out], frames=self.frames, add_into_dst=True, layout=layout)
        return out

    def lazy_tuple(self):
        """(down_block_additional_residuals, mid_block_additional_residual) as LazyResidual proxies."""
        n = self.n_res
        return tuple(LazyResidual(self, i) for i in range(n - 1)), LazyResidual(self, n - 1)

    def materialize(self) -> Tuple[Tuple[torch.Tensor, ...], torch.Tensor]:
        self._join()
        return merge_controlnet_residuals(self.per_net, None, self.frames, scales=self.scales)

    def _materialize_lists(self):
        down, mid = self.materialize()
        return list(down), mid


class LazyResidual:
    """One entry of `down_block_additional_residuals` (or the mid residual) that is still a per-net RAW set.  The
    reference's UNet adds residuals with `skip + residual` (animatediff/models/unet.py:572, :585); for an operand torch does
    not know, Python falls back to `residual.__radd__(skip)`, which runs kernel (3): per-net scales, sum over the nets,
    the '(b f) c h w -> b c f h w' rearrange and the skip add in one pass — without touching the reference's forward."""

    def __init__(self, rset: "ResidualSet", index: int):
        self.rset, self.index = rset, index

    def __radd__(self, skip: torch.Tensor) -> torch.Tensor:
        return self.rset.added_to(skip, self.index)

    __add__ = __radd__

    def materialize(self) -> torch.Tensor:
        down, mid = self.rset.materialize()
        return mid if self.index == self.rset.n_res - 1 else down[self.index]


class RemoteResidualSet(ResidualSet):
    """A ResidualSet whose per-net tensors alias OTHER ranks' symmetric memory (parallel.SymmetricResiduals): kernel (3)
    reads them over NVLink while it scales, sums and adds into the local skips — the ControlNet "reduce" of SURVEY §8e
    fused into the consumer.  The first add waits (on the stream, device side) until every owner has published its
    residuals; after the last read of the step (the mid-block add, unet.py:584-585) the owners are released."""

    def __init__(self, per_net, cond_scale, frames, guess_mode, transport, owners: Sequence[int]):
        super().__init__(per_net, cond_scale, frames, guess_mode)
        self.transport = transport
        self.owners = sorted({int(o) for o in owners if int(o) != transport.rank})
        self._waited = False

    def add_into(self, skips, mid):
        if not self._waited:
            for o in self.owners:
                self.transport.wait_published(o)
            self._waited = True
        super().add_into(skips, mid)
        if mid is not None:
            for o in self.owners:
                self.transport.release(o)


class _MergedResiduals(ResidualSet):
    """Already-merged tensors in the reference layout [b, c, f, h, w] (the plain B3 contract)."""

    def __init__(self, down, mid, frames):
        self.down, self.mid, self.frames = down, mid, frames
        self._pending = None

    def add_into(self, skips, mid):
        from .layers import video5
        if skips is not None:
            for s, r in zip(skips, self.down):
                video5(s, self.frames).add_(r.to(s.dtype))     # strided elementwise add, batch broadcast as unet.py:572
        if mid is not None and self.mid is not None:
            video5(mid, self.frames).add_(self.mid.to(mid.dtype))


def merge_controlnet_residuals(per_net: Sequence[Sequence[torch.Tensor]], cond_scale: Optional[Sequence[float]], frames: int,
                               guess_mode: bool = False, scales=None) -> Tuple[Tuple[torch.Tensor, ...], torch.Tensor]:
    """Per-net raw residual lists ([(b f), c, h, w], NCHW-contiguous as diffusers ControlNets emit them) ->
    (down_block_additional_residuals[12], mid_block_additional_residual) in [b, c, f, h, w] (one kernel launch)."""
    n_res = len(per_net[0])
    scales = _scales(cond_scale, n_res, guess_mode) if scales is None else scales
    first = per_net[0]
    nchw = first[0].is_contiguous()
    layout = L.CA_LAYOUT_NCFHW if nchw else L.CA_LAYOUT_BFHWC
    dst = []
    for r in first:
        n, c, h, w = r.shape
        b = n // frames
        if nchw:
            dst.append(torch.empty((b, c, frames, h, w), dtype=r.dtype, device=r.device))
        else:
            dst.append(torch.empty((b, frames, h, w, c), dtype=r.dtype, device=r.device).permute(0, 4, 1, 2, 3))
    ops.residual_merge(per_net, scales, dst, frames=frames, add_into_dst=False, layout=layout)
    return tuple(dst[:-1]), dst[-1]
