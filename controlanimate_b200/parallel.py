"""Multi-GPU partitioning of the hot loop.  One process per GPU, torch.distributed (NCCL over
NVLink 5 / NVSwitch on the B200 box; gloo in the CPU tests).  The path shards only along three natural axes:

* frame windows with latent overlap  -> `WindowParallel`  (per-step send/recv of the overlap frames + linear blend)
* CFG cond / uncond halves           -> `CFGParallel`     (all-gather of the [1,4,f,h,w] noise prediction)
* the ControlNets of a Multi-ControlNet set -> `ControlNetParallel` (gloo / fallback transport: the raw residual sets are
  all-gathered) and `StepParallel` + `SymmetricResiduals` (NVLink transport: the ControlNet ranks' zero-convs write their
  RAW residuals into symmetric memory and kernel (3) on the UNet rank reads the peers' buffers directly — scale, sum over
  nets and skip add in ONE pass over NVLink, ordered by device-side signals, no copy and no host synchronisation)

`StepParallel` composes the CFG split with the ControlNet sharding for ONE window (strong scaling: step latency), and
`WindowParallel` stacks windows on top (weak scaling: frames/s).

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


# --------------------------------------------------------------------------------------------------
# One window over several GPUs: CFG halves x (UNet rank + ControlNet ranks)
# --------------------------------------------------------------------------------------------------
class StepParallel:
    """Rank layout of one denoising step.

    mode "cfg"            : world = 2.  Rank r evaluates CFG row r (ControlNets + UNet at b = 1).
    mode "controlnet"     : world = G.  Rank 0 runs the UNet (both CFG rows), ranks 1..G-1 the ControlNets.
    mode "cfg+controlnet" : world = 2 G.  CFG half h = rank // G; inside a half, role 0 runs the UNet, roles 1..G-1 the
                            ControlNets (net k on role 1 + k % (G - 1)).
    The UNet rank starts its down path immediately; the ControlNet ranks' residuals are only needed at the skip add
    (unet.py:567-585), so their 2 x 4.6 TFLOP overlap the UNet's first ~7 TFLOP instead of preceding all 17.8.
    """

    MODES = ("cfg", "controlnet", "cfg+controlnet")

    def __init__(self, mode: str, rank: int, world: int, n_nets: int, group=None):
        if mode not in self.MODES:
            raise ValueError(f"unknown step parallelism {mode!r}")
        self.mode, self.rank, self.world, self.n_nets, self.group = mode, rank, world, n_nets, group
        if mode == "cfg":
            if world != 2:
                raise ValueError("CFG parallelism is 2-way")
            self.halves, self.g = 2, 1
        elif mode == "controlnet":
            self.halves, self.g = 1, world
        else:
            if world % 2:
                raise ValueError("cfg+controlnet needs an even number of ranks")
            self.halves, self.g = 2, world // 2
        if self.g > 1 and (self.g - 1 > n_nets or n_nets < 1):
            raise ValueError(f"{self.g - 1} ControlNet ranks per CFG half for {n_nets} nets: a rank without a net has nothing to do")
        self.rows_per_unet = 2               # batch rows a UNet rank returns when CFG is not split (1 in the LCM branch)
        self.half = rank // self.g           # which CFG row(s) this rank works on
        self.role = rank % self.g            # 0 = UNet rank of the half

    @property
    def is_unet_rank(self) -> bool:
        return self.role == 0

    @property
    def unet_rank(self) -> int:
        """Global rank of the UNet rank of my CFG half."""
        return self.half * self.g

    def unet_ranks(self) -> List[int]:
        return [h * self.g for h in range(self.halves)]

    def rows(self, t: torch.Tensor) -> torch.Tensor:
        """CFG rows of a [2, ...] tensor this rank's half evaluates (both rows when CFG is not split)."""
        return t[self.half:self.half + 1] if self.halves == 2 else t

    def nets_of(self, role: int) -> List[int]:
        if self.g == 1:
            return list(range(self.n_nets)) if role == 0 else []
        if role == 0:
            return []
        return [k for k in range(self.n_nets) if k % (self.g - 1) == role - 1]

    def my_nets(self) -> List[int]:
        return self.nets_of(self.role)

    def owner_of(self, net: int) -> int:
        """Global rank (inside my CFG half) that evaluates ControlNet `net`."""
        if self.g == 1:
            return self.unet_rank
        return self.unet_rank + 1 + net % (self.g - 1)

    def combine_noise(self, noise_local: Optional[torch.Tensor], like: torch.Tensor, guidance_scale: float) -> torch.Tensor:
        """All ranks obtain the guided noise prediction (controlanimation_pipeline.py:845-846): the UNet ranks contribute
        their rows ([1, 4, f, h, w] each: cfg2 524 KB), the ControlNet ranks a dummy of the same shape."""
        if noise_local is not None:
            mine = noise_local.contiguous()
        else:
            mine = torch.zeros((1 if self.halves == 2 else self.rows_per_unet, *like.shape[1:]), dtype=like.dtype, device=like.device)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        rows = [parts[r] for r in self.unet_ranks()]
        if self.halves == 2:
            return rows[0] + guidance_scale * (rows[1] - rows[0])
        if rows[0].shape[0] == 1:        # LCM branch: a single row, the guidance scale went in through timestep_cond
            return rows[0]
        u, c = rows[0].chunk(2)
        return u + guidance_scale * (c - u)


class SymmetricResiduals:
    """The NVLink transport of the ControlNet sharding: one symmetric-memory arena per rank holding the RAW residuals of the
    nets that rank evaluates ([(b f), c_i, h_i, w_i] channels_last, 13 per net, back to back).  Owners write through
    `out_views(slot)` (the zero-conv GEMMs store straight into the arena); the UNet rank reads `peer_views(rank, slot)` —
    tensors that alias the PEER's memory, handed to kernel (3) like local ones (`ld.global` over NVLink).  Ordering is by
    device-side signals on the current stream: `publish()` / `wait_published()` before the merge, `release()` /
    `wait_released()` before the owners overwrite the arena in the next step."""

    def __init__(self, shapes: Sequence[Tuple[int, int, int, int]], slots: int, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.shapes = [tuple(int(v) for v in s) for s in shapes]              # (n, c, h, w) of the 13 residuals of one net
        self.per_net = sum(n * c * h * w for n, c, h, w in self.shapes)
        self.slots, self.dtype, self.device = slots, dtype, device
        group = group if group is not None else dist.group.WORLD
        self.arena = symm_mem.empty(max(1, slots) * self.per_net, dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.arena, group)
        self.rank = self.handle.rank

    def _views(self, flat: torch.Tensor, slot: int) -> List[torch.Tensor]:
        out, off = [], slot * self.per_net
        for n, c, h, w in self.shapes:
            out.append(flat[off:off + n * c * h * w].view(n, h, w, c).permute(0, 3, 1, 2))   # channels_last [(b f), c, h, w]
            off += n * c * h * w
        return out

    def out_views(self, slot: int) -> List[torch.Tensor]:
        return self._views(self.arena, slot)

    def peer_views(self, rank: int, slot: int) -> List[torch.Tensor]:
        if rank == self.rank:
            return self.out_views(slot)
        flat = self.handle.get_buffer(rank, (self.arena.numel(),), self.dtype, 0)
        return self._views(flat, slot)

    # device-side ordering (stream-ordered signal kernels on the symmetric signal pads; no host synchronisation)
    def publish(self, to_rank: int):
        self.handle.put_signal(to_rank, channel=0)

    def wait_published(self, from_rank: int):
        self.handle.wait_signal(from_rank, channel=0)

    def release(self, to_rank: int):
        self.handle.put_signal(to_rank, channel=1)

    def wait_released(self, from_rank: int):
        self.handle.wait_signal(from_rank, channel=1)


class ClipLayout:
    """Rank layout of a whole clip on one box (the 8-GPU plan of BASELINE config 3): `n_windows` UNet ranks —
    rank w owns window w and evaluates the first `local_nets` ControlNets of its own window — and `world - n_windows`
    SERVER ranks that evaluate the remaining (window, net) jobs, dealt window-major in contiguous, balanced runs (config 3 on
    8 GPUs: 5 windows x 4 nets -> 5 local jobs + 15 served ones, 5 per server).  A UNet rank ships its latents to its
    servers at the start of a step and reads their raw residuals over NVLink in kernel (3) (`ClipTransport`)."""

    def __init__(self, rank: int, world: int, n_windows: int, n_nets: int, local_nets: int = 1):
        if not (1 <= n_windows <= world):
            raise ValueError(f"{n_windows} windows need at least as many ranks (got {world})")
        if not (0 <= local_nets <= n_nets):
            raise ValueError("local_nets must be within [0, n_nets]")
        self.rank, self.world, self.n_windows, self.n_nets = rank, world, n_windows, n_nets
        self.n_servers = world - n_windows
        if self.n_servers == 0:
            local_nets = n_nets                      # nobody to serve: every window evaluates all of its own nets
        self.local_nets = local_nets
        served = [(w, k) for w in range(n_windows) for k in range(local_nets, n_nets)]
        if served and self.n_servers > len(served):
            raise ValueError(f"{self.n_servers} server ranks for {len(served)} ControlNet jobs: a server would idle")
        self.jobs = {r: [] for r in range(world)}                  # rank -> [(window, net)] in evaluation order
        for w in range(n_windows):
            self.jobs[w] = [(w, k) for k in range(local_nets)]
        for s in range(self.n_servers):                            # contiguous balanced split of the served jobs
            lo, hi = s * len(served) // self.n_servers, (s + 1) * len(served) // self.n_servers
            self.jobs[n_windows + s] = served[lo:hi]
        self._owner = {job: r for r, js in self.jobs.items() for job in js}

    @property
    def is_unet_rank(self) -> bool:
        return self.rank < self.n_windows

    @property
    def window(self) -> Optional[int]:
        return self.rank if self.is_unet_rank else None

    def unet_ranks(self) -> List[int]:
        return list(range(self.n_windows))

    def owner_of(self, window: int, net: int) -> int:
        return self._owner[(window, net)]

    def slot_of(self, window: int, net: int) -> int:
        """Arena slot of the job on its owner (position in the owner's job list)."""
        return self.jobs[self.owner_of(window, net)].index((window, net))

    def servers_of(self, window: int) -> List[int]:
        return sorted({self.owner_of(window, k) for k in range(self.local_nets, self.n_nets)})

    def windows_of(self, server: int) -> List[int]:
        """Windows a server works for, in the order it serves them."""
        out: List[int] = []
        for w, _ in self.jobs[server]:
            if w not in out:
                out.append(w)
        return out

    def max_slots(self) -> int:
        return max(len(js) for js in self.jobs.values())


class ClipTransport(SymmetricResiduals):
    """SymmetricResiduals (raw residual arenas + publish / release signals) plus the per-step latent hand-off of a
    ClipLayout: a second symmetric buffer with one latent slot per window; UNet rank w stores its window's latents into slot
    w of each of its servers (a peer write over NVLink) and raises a signal the server's stream waits on."""

    def __init__(self, shapes, slots: int, dtype, device, n_windows: int, latent_shape: Sequence[int], latent_dtype, group=None):
        super().__init__(shapes, slots, dtype, device, group)
        import torch.distributed._symmetric_memory as symm_mem
        self.latent_shape = tuple(int(v) for v in latent_shape)
        self.latent_numel = 1
        for v in self.latent_shape:
            self.latent_numel *= v
        self.latent_dtype, self.n_windows = latent_dtype, n_windows
        self.lat = symm_mem.empty(n_windows * self.latent_numel, dtype=latent_dtype, device=device)
        self.lat_handle = symm_mem.rendezvous(self.lat, group if group is not None else dist.group.WORLD)

    def send_latents(self, server: int, window: int, latents: torch.Tensor):
        peer = self.lat_handle.get_buffer(server, (self.n_windows, self.latent_numel), self.latent_dtype, 0)
        peer[window].copy_(latents.reshape(-1))
        self.lat_handle.put_signal(server, channel=0)

    def recv_latents(self, window: int, from_rank: int) -> torch.Tensor:
        self.lat_handle.wait_signal(from_rank, channel=0)
        return self.lat.view(self.n_windows, self.latent_numel)[window].view(self.latent_shape)
