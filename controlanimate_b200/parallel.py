"""Multi-GPU partitioning of the hot loop.  One process per GPU, torch.distributed (NCCL over
NVLink 5 / NVSwitch on the B200 box; gloo in the CPU tests).  The path shards only along three natural axes:

* frame windows with latent overlap  -> `WindowParallel`  (per-step send/recv of the overlap frames + linear blend)
* CFG cond / uncond halves           -> `CFGParallel`     (all-gather of the [1,4,f,h,w] noise prediction)
* the ControlNets of a Multi-ControlNet set -> `ControlNetParallel` (each rank runs its nets; the residual sets are
  reduced into the UNet rank's skips — by kernel (3) reading the peers' buffers through symmetric memory when
  available, else after an all-gather)

Everything else is replicated.  The reference has no distributed code; its only long-video mechanism is the
SEQUENTIAL sliding window of scripts/vid2vid.py:168-231 (pixel-space hand-off + cross-fade :225-226), which
`WindowParallel` reformulates as parallel windows blended in latent space every step.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def window_starts(n_windows: int, frames: int, overlap: int) -> List[int]:
    """Window w owns frames [w*(frames-overlap), w*(frames-overlap)+frames) (config 3: 0,12,24,36,48)."""
    return [w * (frames - overlap) for w in range(n_windows)]


def blend_ramp(overlap: int, device=None, dtype=torch.float32) -> torch.Tensor:
    """Weight of the LATER window on the shared frames, rising linearly as vid2vid.py:225-226 cross-fades."""
    return (torch.arange(1, overlap + 1, device=device, dtype=dtype) / (overlap + 1)).reshape(1, 1, overlap, 1, 1)


class WindowParallel:
    def __init__(self, rank: int, world: int, frames: int, overlap: int, group=None):
        if overlap * 2 > frames:
            raise ValueError("overlap must be at most half a window")
        self.rank, self.world, self.frames, self.overlap, self.group = rank, world, frames, overlap, group

    def exchange(self, latents: torch.Tensor) -> torch.Tensor:
        """latents [b,c,frames,h,w] of this rank's window -> same, with the `overlap` frames shared with each neighbour
        replaced by the blend of both windows (identical values on both ranks)."""
        ov = self.overlap
        if self.world == 1 or ov == 0:
            return latents
        lo, hi = self.rank - 1, self.rank + 1
        head = latents[:, :, :ov].contiguous()
        tail = latents[:, :, -ov:].contiguous()
        ops, recv_lo, recv_hi = [], None, None
        if lo >= 0:
            recv_lo = torch.empty_like(head)
            ops += [dist.P2POp(dist.isend, head, lo, self.group), dist.P2POp(dist.irecv, recv_lo, lo, self.group)]
        if hi < self.world:
            recv_hi = torch.empty_like(tail)
            ops += [dist.P2POp(dist.isend, tail, hi, self.group), dist.P2POp(dist.irecv, recv_hi, hi, self.group)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        w = blend_ramp(ov, latents.device, latents.dtype)
        out = latents.clone()
        if recv_lo is not None:   # I am the LATER window on my first `ov` frames
            out[:, :, :ov] = w * head + (1 - w) * recv_lo
        if recv_hi is not None:   # I am the EARLIER window on my last `ov` frames
            out[:, :, -ov:] = w * recv_hi + (1 - w) * tail
        return out


def blend_windows_reference(windows: Sequence[torch.Tensor], overlap: int) -> List[torch.Tensor]:
    """Single-process statement of `WindowParallel.exchange` over all windows (used by the tests)."""
    out = [w.clone() for w in windows]
    ramp = blend_ramp(overlap, windows[0].device, windows[0].dtype)
    for i in range(len(windows) - 1):
        shared = ramp * windows[i + 1][:, :, :overlap] + (1 - ramp) * windows[i][:, :, -overlap:]
        out[i][:, :, -overlap:] = shared
        out[i + 1][:, :, :overlap] = shared
    return out


class CFGParallel:
    """2-way split of classifier-free guidance: rank 0 evaluates the unconditional row, rank 1 the conditional one
    (batch rows are independent everywhere: GroupNorm statistics are per-b, attention per-b), then the two noise
    predictions are all-gathered (cfg2: 2 x 524 KB) and combined on both ranks (controlanimation_pipeline.py:845-846)."""

    def __init__(self, rank: int, world: int, group=None):
        if world != 2:
            raise ValueError("CFG parallelism is 2-way")
        self.rank, self.world, self.group = rank, world, group

    def my_rows(self, t: torch.Tensor) -> torch.Tensor:
        return t[self.rank:self.rank + 1]

    def combine(self, noise_local: torch.Tensor, guidance_scale: float) -> torch.Tensor:
        parts = [torch.empty_like(noise_local) for _ in range(2)]
        dist.all_gather(parts, noise_local.contiguous(), group=self.group)
        return parts[0] + guidance_scale * (parts[1] - parts[0])


class ControlNetParallel:
    """ControlNets sharded over ranks: rank r evaluates nets {k : k % world == r}; the per-net RAW residual sets are
    then made visible to the UNet rank.  `gather` returns all N per-net lists in net order so that kernel (3) can
    apply the per-net scales and reduce in fp32 at the consumer (exact up to summation order)."""

    def __init__(self, rank: int, world: int, n_nets: int, group=None):
        if world > n_nets:
            raise ValueError(f"ControlNet sharding needs world <= n_nets (got {world} ranks for {n_nets} nets): a rank "
                             "without a net has no residual shapes to contribute")
        self.rank, self.world, self.n_nets, self.group = rank, world, n_nets, group

    def my_nets(self) -> List[int]:
        return [k for k in range(self.n_nets) if k % self.world == self.rank]

    def gather(self, mine: Sequence[Sequence[torch.Tensor]]) -> List[List[torch.Tensor]]:
        """mine: residual lists of my nets (my_nets() order).  Returns the lists of all nets, on every rank."""
        if self.world == 1:
            return [list(r) for r in mine]
        per_rank_counts = [len([k for k in range(self.n_nets) if k % self.world == r]) for r in range(self.world)]
        rounds = max(per_rank_counts)
        out: List[Optional[List[torch.Tensor]]] = [None] * self.n_nets
        template = mine[0]
        for j in range(rounds):
            src = mine[j] if j < len(mine) else template       # ranks with fewer nets contribute a dummy set
            flat = torch.cat([t.reshape(-1) for t in src])
            bufs = [torch.empty_like(flat) for _ in range(self.world)]
            dist.all_gather(bufs, flat, group=self.group)
            for r in range(self.world):
                k = r + j * self.world
                if k < self.n_nets and j < per_rank_counts[r]:
                    parts, off = [], 0
                    for t in template:
                        parts.append(bufs[r][off:off + t.numel()].view(t.shape).contiguous(
                            memory_format=torch.channels_last if t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous() else torch.contiguous_format))
                        off += t.numel()
                    out[k] = parts
        return out  # type: ignore[return-value]
