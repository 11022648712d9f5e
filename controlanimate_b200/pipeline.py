"""Loop-level host code of the hot path: the per-step body of
ControlAnimationPipeline.__call__ (reference animatediff/pipelines/controlanimation_pipeline.py:793-849) and
MultiControlNetResidualsPipeline.__call__ (modules/controlresiduals_pipeline.py:278-316), on the B200 modules.

Out of scope here: prompt encoding, VAE, annotators, video I/O.  Inputs are prompt embeddings,
latents and already-prepared control images.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .residuals import RemoteResidualSet, ResidualSet, merge_controlnet_residuals
from .unet import ControlNetModel, UNet3DConditionModel


class DDIMScheduler:
    """η=0 DDIM with the reference's settings (inference-v2.yaml:24-27 linear betas; steps_offset=1 and
    clip_sample=False forced at controlanimation_pipeline.py:103-128).  diffusers semantics."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.init_noise_sigma = 1.0
        self.timesteps: List[int] = []
        self.num_inference_steps = 0

    def set_timesteps(self, n: int):
        ratio = self.num_train_timesteps // n
        self.num_inference_steps = n
        self.timesteps = [int(round(k * ratio)) + self.steps_offset for k in range(n)][::-1]
        return self.timesteps

    def coefficients(self, t: int) -> Tuple[float, float, float, float]:
        """(sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev))."""
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else 1.0
        return a_t ** 0.5, (1 - a_t) ** 0.5, a_p ** 0.5, (1 - a_p) ** 0.5

    def scale_model_input(self, x, t=None):
        return x

    def step(self, noise_pred: torch.Tensor, t: int, sample: torch.Tensor) -> torch.Tensor:
        sa, s1a, sp, s1p = self.coefficients(t)
        x0 = (sample - s1a * noise_pred) / sa
        return sp * x0 + s1p * noise_pred


def get_w_embedding(w: torch.Tensor, embedding_dim: int = 256, dtype=torch.float32) -> torch.Tensor:
    """LCM guidance-scale embedding (reference controlanimation_pipeline.py:477-498): sinusoidal features of 1000 w, fed to
    `time_embedding.cond_proj` as `timestep_cond` (unet.py:526-534)."""
    if w.dim() != 1:
        raise ValueError("w must be 1-D")
    half = embedding_dim // 2
    freq = torch.exp(torch.arange(half, dtype=dtype, device=w.device) * -(torch.log(torch.tensor(10000.0)) / (half - 1)))
    emb = (w.to(dtype) * 1000.0)[:, None] * freq[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
    return torch.nn.functional.pad(emb, (0, 1)) if embedding_dim % 2 == 1 else emb


class MultiControlNetResiduals:
    """Mirror of MultiControlNetResidualsPipeline.__call__ (modules/controlresiduals_pipeline.py:278-316) for N native
    ControlNets.  `prep_images[k]` is the prepared control video of net k as [(b f), 3, H, W] (what
    `prep_control_images` :226-273 leaves in `self.prep_images`, already duplicated for CFG)."""

    def __init__(self, controlnets: Sequence[ControlNetModel], cond_scale: Sequence[float]):
        if len(controlnets) != len(cond_scale):
            raise ValueError("one conditioning scale per ControlNet")
        self.controlnets = list(controlnets)
        self.cond_scale = [float(s) for s in cond_scale]
        self.prep_images: Optional[List[torch.Tensor]] = None
        self.lazy = True  # hand the UNet a ResidualSet (single-pass merge) instead of merged tensors
        # overlap: every ControlNet runs on its own CUDA stream, forked off the caller's stream and joined by the first
        # consumer of the ResidualSet (the skip adds after the UNet's down path, unet.py:567-576) — the nets do not depend
        # on each other or on the UNet encoder, and their 16x16 / 8x8 levels leave most of the 148 SMs idle on their own
        self.overlap = False
        self._streams: List[torch.cuda.Stream] = []
        self._pending = None
        # hoist_cond_embedding: `controlnet_cond_embedding(image)` (eight convolutions at up to full image resolution per net)
        # depends on the prepared control video only, which is fixed for all steps of a window: evaluate it once per
        # (net, images) instead of once per step as diffusers does.  Same arithmetic on the same operands; off by default so
        # that a step of this class does exactly the work of a reference step.
        self.hoist_cond_embedding = False
        self._cond_cache = {}

    def cond_embedding(self, k: int, image: torch.Tensor, into: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The hoisted conditioning embedding of net k for `image`: one cache entry per (net, image buffer), recomputed when
        the buffer's contents (version counter), shape or the embedding weights changed.  `into`: a caller-owned buffer the
        value must live in (a captured CUDA graph reads the embedding through a baked-in pointer)."""
        net = self.controlnets[k]
        # the entry HOLDS the image's storage owner: while it is cached its address cannot be recycled for another window's
        # images, so (owner identity, address, version) cannot hit stale contents
        owner = image._base if image._base is not None else image
        slot = (k, image.data_ptr(), tuple(image.shape), image.dtype)
        ident = (image._version, tuple((p.data_ptr(), p._version) for p in net.controlnet_cond_embedding.parameters()))
        ent = self._cond_cache.get(slot)
        if ent is not None and ent[0] == ident and ent[2] is owner and (into is None or ent[1] is into):
            return ent[1]
        if image.is_cuda and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("the hoisted ControlNet conditioning embedding must be evaluated before the graph capture")
        val = net.embed_condition(image)
        if into is not None:
            into.copy_(val)
            val = into
        if len(self._cond_cache) >= 64 and slot not in self._cond_cache:      # image buffers come and go with the windows
            self._cond_cache.clear()
        self._cond_cache[slot] = (ident, val, owner)
        return val

    def raw(self, control_model_input, t, controlnet_prompt_embeds, frame_count, nets=None, images=None, out=None,
            sample_offset: int = 0):
        """Raw residual lists of the nets in `nets` (default: all).  `images[k]` overrides prep_images[k] (a CFG half's
        rows), `out[j]` gives the 13 output buffers of the j-th evaluated net (symmetric memory of a sharded set),
        `sample_offset` the index of this call's first sample in the full (b f) batch (a CFG half evaluates rows
        [half * f, (half + 1) * f) and must see the prompts the reference's tiling quirk gives exactly those rows)."""
        b = control_model_input.shape[0]
        x = control_model_input.permute(0, 2, 1, 3, 4).reshape(b * frame_count, *control_model_input.shape[1:2],
                                                               *control_model_input.shape[3:])   # :287
        # :292 tiles the prompt embeddings as cat([e]*f): row n of the tiled tensor is e[n % b] (kept as-is)
        n_prompts = controlnet_prompt_embeds.shape[0]
        ctx_map = (sample_offset + torch.arange(b * frame_count, device=x.device)) % n_prompts
        nets = list(range(len(self.controlnets))) if nets is None else list(nets)
        images = self.prep_images if images is None else images
        hoisted = {k: self.cond_embedding(k, images[k]) for k in nets} if self.hoist_cond_embedding else {}
        if not (self.overlap and self.lazy and x.is_cuda):
            return [self.controlnets[k](x, t, controlnet_prompt_embeds, images[k], ctx_map=ctx_map,
                                        out=None if out is None else out[j], cond_embedding=hoisted.get(k))
                    for j, k in enumerate(nets)]
        from .layers import _ctx_i32
        _ctx_i32(ctx_map)        # shared by every net: convert on the caller's stream, BEFORE the fork (the nets only hit the cache)
        main = torch.cuda.current_stream(x.device)
        while len(self._streams) < len(nets):
            self._streams.append(torch.cuda.Stream(x.device))
        fork = torch.cuda.Event()
        fork.record(main)
        per_net, events = [], []
        for j, k in enumerate(nets):
            side = self._streams[j]
            side.wait_event(fork)
            with torch.cuda.stream(side):
                per_net.append(self.controlnets[k](x, t, controlnet_prompt_embeds, images[k], ctx_map=ctx_map,
                                                   out=None if out is None else out[j], cond_embedding=hoisted.get(k)))
                ev = torch.cuda.Event()
                ev.record(side)
                events.append(ev)
        self._pending = (events, (x, ctx_map, t, controlnet_prompt_embeds, images))
        return per_net

    def join(self, rset=None):
        """Hand the outstanding side-stream work of the last raw() to `rset` (its first consumer waits), or wait now."""
        pending, self._pending = self._pending, None
        if pending is None:
            return rset
        if rset is not None:
            return rset.produced_on(*pending)
        cur = torch.cuda.current_stream()
        for ev in pending[0]:
            cur.wait_event(ev)
        return None

    def __call__(self, control_model_input, t, controlnet_prompt_embeds, frame_count, image_embeds=None,
                 do_classifier_free_guidance=True, guess_mode=True):
        per_net = self.raw(control_model_input, t, controlnet_prompt_embeds, frame_count)
        if self.lazy:
            rs = self.join(ResidualSet(per_net, self.cond_scale, frame_count, guess_mode))
            return rs, None
        return merge_controlnet_residuals(per_net, self.cond_scale, frame_count, guess_mode)     # :294-316


class DenoisingLoop:
    """The hot loop (controlanimation_pipeline.py:790-856) for the non-LCM path: CFG-duplicated latents, ControlNets,
    UNet3D, `eps = eps_u + g (eps_c - eps_u)` (:845-846), scheduler step (:849).  No per-step empty_cache().

    use_cuda_graph: the noise prediction of a step (ControlNets -> kernel-(3) merge -> UNet3D -> CFG; ~5000 launches) is
    captured once per input signature into a CUDA graph and replayed, so the step is bound by the GPU, not by Python.
    """

    def __init__(self, unet: UNet3DConditionModel, controlnets: Optional[MultiControlNetResiduals], scheduler: DDIMScheduler,
                 guidance_scale: float = 7.5, guess_mode: bool = False, use_cuda_graph: bool = False, parallel=None,
                 use_lcm: bool = False):
        """parallel: a `parallel.StepParallel` — this window's step is split over its ranks (CFG halves and / or ControlNet
        sharding over NVLink); every rank of the group must call `step` with the same latents.
        use_lcm: the LCM branch of the loop (controlanimation_pipeline.py:770-771, 823-833): no CFG duplication (b = 1), the
        guidance scale enters through `timestep_cond = get_w_embedding(guidance_scale)`, prompt_embeds is the positive prompt
        [1, L, D].  (The LCM scheduler itself is outside the hot path; `scheduler.step` is applied to the model output.)"""
        self.unet, self.controlnets, self.scheduler = unet, controlnets, scheduler
        self.guidance_scale, self.guess_mode = guidance_scale, guess_mode
        self.use_lcm = use_lcm
        if use_lcm and parallel is not None and parallel.halves == 2:
            raise ValueError("the LCM branch has no CFG halves to split")
        self.use_cuda_graph = use_cuda_graph
        self.fused_update = True   # step(): guidance combine + DDIM update as one kernel (ca_cfg_ddim_step) instead of torch ops
        self._graphs = {}
        self.parallel = parallel
        self._transport = None
        if parallel is not None:
            parallel.rows_per_unet = 1 if use_lcm else 2
            if guess_mode or (guidance_scale <= 1.0 and not use_lcm):
                raise ValueError("step parallelism covers the CFG path without guess mode (BASELINE configs 2-4)")
            if parallel.g > 1 and controlnets is None:
                raise ValueError("ControlNet sharding needs a ControlNet set")

    @property
    def do_cfg(self):
        return self.guidance_scale > 1.0 and not self.use_lcm

    def _w_embedding(self, latents: torch.Tensor) -> Optional[torch.Tensor]:
        if not self.use_lcm:
            return None
        dim = self.unet.config.get("time_cond_proj_dim")
        if not dim:
            raise ValueError("use_lcm needs a UNet built with time_cond_proj_dim (LCM checkpoints: 256)")
        w = torch.full((latents.shape[0],), float(self.guidance_scale), device=latents.device)
        return get_w_embedding(w, dim).to(self.unet.conv_in.weight.dtype)

    def predict_noise(self, latents: torch.Tensor, t, prompt_embeds: torch.Tensor) -> torch.Tensor:
        """Guided noise prediction for latents [1,4,f,h,w]; `t` is an int or a 1-element int64 device tensor."""
        if self.parallel is not None:
            local = self._local_noise(latents, t, prompt_embeds)
            return self.parallel.combine_noise(local, latents, self.guidance_scale)
        noise = self._model_out(latents, t, prompt_embeds).to(latents.dtype)                             # :841
        if self.do_cfg:
            u, c = noise.chunk(2)
            noise = u + self.guidance_scale * (c - u)                                                    # :845-846
        return noise

    def _model_out(self, latents: torch.Tensor, t, prompt_embeds: torch.Tensor) -> torch.Tensor:
        """The UNet output of one step in the model dtype, rows [uncond, cond] with CFG (ControlNets -> kernel (3) -> UNet3D)."""
        f = latents.shape[2]
        cfg = self.do_cfg
        model_in = torch.cat([latents] * 2) if cfg else latents                                          # :797
        down = mid = None
        if self.controlnets is not None:                                                                 # :807-819
            single = (self.guess_mode or not cfg) and cfg
            down, mid = self.controlnets(model_in[-1:] if single else model_in, t,
                                         prompt_embeds[-1:] if single else prompt_embeds, f,
                                         do_classifier_free_guidance=cfg, guess_mode=self.guess_mode)
        return self.unet(model_in, t, encoder_hidden_states=prompt_embeds, down_block_additional_residuals=down,
                         mid_block_additional_residual=mid, timestep_cond=self._w_embedding(model_in)).sample   # :823-841

    def invalidate_graphs(self):
        """Drop every captured graph (call after swapping modules, e.g. install() / set_ip_adapter)."""
        self._graphs.clear()
        self._params = None

    def _weight_stamp(self):
        """Cheap identity of every weight the captured graph baked in (device pointers of the parameters and of their
        cached fp32 / fused / channels_last copies follow from these): storage pointers + in-place version counters."""
        if getattr(self, "_params", None) is None:
            mods = [self.unet] + (list(self.controlnets.controlnets) if self.controlnets is not None else [])
            self._params = [p for m in mods for p in list(m.parameters()) + list(m.buffers())]
        return hash(tuple(p.data_ptr() for p in self._params)), sum(p._version for p in self._params)

    # ---- one window over several GPUs (parallel.StepParallel) --------------------------------------------------------
    def _transport_for(self, latents: torch.Tensor):
        """Symmetric-memory arena for the sharded ControlNets' raw residuals, created (collectively) on first use."""
        sp = self.parallel
        if sp.g == 1:
            return None
        if self._transport is None:
            from .parallel import SymmetricResiduals
            from .residuals import N_RESIDUALS  # noqa: F401
            f, hh, ww = latents.shape[2], latents.shape[3], latents.shape[4]
            rows = (1 if (sp.halves == 2 or self.use_lcm) else 2) * f
            cn = self.controlnets.controlnets[0]
            boc, lpb = cn.config["block_out_channels"], cn.config["layers_per_block"]
            shapes, h, w = [(rows, boc[0], hh, ww)], hh, ww
            for i, c in enumerate(boc):
                shapes += [(rows, c, h, w)] * lpb
                if i != len(boc) - 1:
                    h, w = (h + 1) // 2, (w + 1) // 2
                    shapes.append((rows, c, h, w))
            shapes.append((rows, boc[-1], h, w))
            slots = max(len(sp.nets_of(r)) for r in range(sp.g))
            self._transport = SymmetricResiduals(shapes, slots, self.unet.conv_in.weight.dtype, latents.device, sp.group)
            if sp.is_unet_rank:      # the arenas start out free: the owners' first wait_released must not block
                for o in sorted({sp.owner_of(k) for k in range(sp.n_nets)}):
                    self._transport.release(o)
        return self._transport

    def _local_noise(self, latents: torch.Tensor, t, prompt_embeds: torch.Tensor) -> Optional[torch.Tensor]:
        """This rank's share of one noise prediction (no collective inside: capturable in a CUDA graph).  UNet ranks
        return their CFG row(s) of the UNet output, ControlNet ranks publish their residuals and return None."""
        sp, mc = self.parallel, self.controlnets
        f = latents.shape[2]
        model_in = latents if (sp.halves == 2 or self.use_lcm) else torch.cat([latents] * 2)
        prompt = sp.rows(prompt_embeds)
        # the ControlNets always see BOTH prompts: row n of the (b f) batch takes prompt n % 2 (the tiling quirk of
        # controlresiduals_pipeline.py:292 is part of the reference's arithmetic), n counted over the full CFG batch
        offset = sp.half * f if sp.halves == 2 else 0
        images = None
        if mc is not None:
            images = [im[sp.half * f:(sp.half + 1) * f] if sp.halves == 2 else im for im in mc.prep_images]
        tr = self._transport_for(latents)
        if not sp.is_unet_rank:                                   # ControlNet rank: evaluate, store into the arena, publish
            nets = sp.my_nets()
            tr.wait_released(sp.unet_rank)
            mc.raw(model_in, t, prompt_embeds, f, nets=nets, images=images, out=[tr.out_views(j) for j in range(len(nets))],
                   sample_offset=offset)
            mc.join()
            tr.publish(sp.unet_rank)
            return None
        down = None
        if mc is not None and sp.g == 1:                          # CFG split only: my row's ControlNets run here
            down = mc.join(ResidualSet(mc.raw(model_in, t, prompt_embeds, f, images=images, sample_offset=offset),
                                       mc.cond_scale, f, False))
        elif mc is not None:                                      # sharded: kernel (3) will read the owners' arenas over NVLink
            per_net, owners = [], []
            for k in range(sp.n_nets):
                owner = sp.owner_of(k)
                slot = sp.nets_of(owner - sp.unet_rank).index(k)
                per_net.append(tr.peer_views(owner, slot))
                owners.append(owner)
            down = RemoteResidualSet(per_net, mc.cond_scale, f, False, tr, owners)
        return self.unet(model_in, t, encoder_hidden_states=prompt, down_block_additional_residuals=down,
                         timestep_cond=self._w_embedding(model_in)).sample.to(latents.dtype)

    def _graphed_noise(self, latents, t: int, prompt_embeds, raw: bool = False):
        """Replay (capturing first) this step's noise prediction.  raw: return the UNet output before `.to(latents.dtype)` and
        the guidance combine (the fused update kernel of `step` consumes it)."""
        mc = self.controlnets
        images = list(mc.prep_images) if mc is not None and mc.prep_images is not None else []
        # everything the captured kernels read by value or by baked-in pointer is part of the key: shapes, the scalars of
        # the step, and the weights; the per-window control images are copied into graph-owned buffers before a replay
        key = (tuple(latents.shape), latents.dtype, tuple(prompt_embeds.shape), prompt_embeds.dtype, latents.device,
               float(self.guidance_scale), bool(self.guess_mode), tuple(mc.cond_scale) if mc is not None else (),
               (bool(mc.lazy), bool(mc.overlap), bool(mc.hoist_cond_embedding)) if mc is not None else None,
               tuple((tuple(im.shape), im.dtype) for im in images),
               self._weight_stamp(), bool(raw))
        g = self._graphs.get(key)
        if mc is not None and mc.hoist_cond_embedding and self.parallel is not None:
            raise ValueError("hoisted conditioning embeddings are not wired into the captured step-parallel paths "
                             "(use them eagerly there, or capture without them)")
        if g is None:
            if len(self._graphs) >= 4:          # stale captures pin their private memory pools
                self._graphs.clear()
            s_lat, s_prompt = latents.clone(), prompt_embeds.clone()
            s_images = [im.clone() for im in images]
            s_t = torch.full((1,), int(t), dtype=torch.int64, device=latents.device)
            if mc is not None:
                mc.prep_images = s_images
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                # with step parallelism only this rank's share is captured; the (tiny) all-gather + CFG combine stay eager
                fn = self._local_noise if self.parallel is not None else (self._model_out if raw else self.predict_noise)
                with torch.cuda.stream(side):                # warm-up: cuDNN algorithm selection, caches, workspaces
                    for _ in range(2):
                        fn(s_lat, s_t, s_prompt)
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    s_out = fn(s_lat, s_t, s_prompt)
            finally:
                if mc is not None:
                    mc.prep_images = images if images else None
            # the capture baked in the pointers of the hoisted conditioning embeddings the warm-up passes left in mc's cache:
            # the graph keeps them alive and refreshes them in place when a window's images change
            emb = [mc.cond_embedding(k, im) for k, im in enumerate(s_images)] if mc is not None and mc.hoist_cond_embedding else None
            g = self._graphs[key] = dict(graph=graph, lat=s_lat, t=s_t, prompt=s_prompt, out=s_out, images=s_images,
                                         seen=[None] * len(s_images), emb=emb)
        g["lat"].copy_(latents, non_blocking=True)
        if prompt_embeds.data_ptr() != g["prompt"].data_ptr():
            g["prompt"].copy_(prompt_embeds, non_blocking=True)
        for k, im in enumerate(images):      # the reference re-runs prep_control_images for every window (pipeline :226-273)
            seen = g["seen"][k]      # (tensor, version): holding the tensor keeps its address from being recycled unnoticed
            if seen is None or seen[0] is not im or seen[1] != im._version:
                g["images"][k].copy_(im, non_blocking=True)
                g["seen"][k] = (im, im._version)
        if g["emb"] is not None:             # (no-op unless the images above changed or another loop used mc in between)
            for k, buf in enumerate(g["emb"]):
                mc.cond_embedding(k, g["images"][k], into=buf)
        g["t"].fill_(int(t))
        g["graph"].replay()
        if self.parallel is not None:
            return self.parallel.combine_noise(g["out"], latents, self.guidance_scale)
        return g["out"]

    @torch.no_grad()
    def step(self, latents: torch.Tensor, t: int, prompt_embeds: torch.Tensor) -> torch.Tensor:
        """latents [1,4,f,h,w] (fp32 or model dtype), prompt_embeds [2,L,D] = [negative, positive] when CFG."""
        fused = (self.fused_update and latents.is_cuda and latents.is_contiguous() and type(self.scheduler) is DDIMScheduler
                 and latents.dtype in (torch.float32, torch.bfloat16, torch.float16))
        if fused and self.parallel is None:
            # N4: `.to(latents_dtype)`, the guidance combine and the scheduler step as one launch on the raw UNet output
            from . import ops
            raw = self._graphed_noise(latents, t, prompt_embeds, raw=True) if self.use_cuda_graph else \
                self._model_out(latents, t, prompt_embeds)
            return ops.cfg_ddim_step(raw, latents, self.guidance_scale if self.do_cfg else None,
                                     self.scheduler.coefficients(int(t)))                                # :841-849
        if self.use_cuda_graph:
            noise = self._graphed_noise(latents, t, prompt_embeds)
        else:
            noise = self.predict_noise(latents, t, prompt_embeds)
        if fused and noise.is_contiguous() and noise.shape == latents.shape:
            from . import ops
            return ops.cfg_ddim_step(noise, latents, None, self.scheduler.coefficients(int(t)))          # :849
        return self.scheduler.step(noise, int(t), latents)                                               # :849

    @torch.no_grad()
    def run(self, latents, prompt_embeds, num_inference_steps: int):
        for t in self.scheduler.set_timesteps(num_inference_steps):
            latents = self.step(latents, t, prompt_embeds)
        return latents


class ClipLoop:
    """One denoising step of a WHOLE clip on one box (parallel.ClipLayout): UNet rank w advances window w, server ranks
    evaluate ControlNet jobs for several windows; the raw residuals travel through symmetric memory and are scaled, summed
    and added into the skips by kernel (3) reading the servers' arenas over NVLink (parallel.ClipTransport).

    images[(w, k)]: prepared control video of net k for window w ([(b f), 3, H, W], CFG-duplicated when CFG is on) — needed
    on the rank that evaluates job (w, k).  Every rank calls `step` once per timestep: UNet ranks pass and get back their
    window's latents (after the scheduler step and the overlap blend with the neighbouring windows), server ranks pass None.
    """

    def __init__(self, unet: UNet3DConditionModel, controlnets: MultiControlNetResiduals, scheduler: DDIMScheduler, layout,
                 images: dict, latent_shape: Sequence[int], guidance_scale: float = 7.5, use_lcm: bool = False,
                 use_cuda_graph: bool = False, overlap: int = 4, unet_group=None, group=None):
        from .parallel import ClipTransport, WindowParallel
        self.unet, self.mc, self.scheduler, self.layout, self.images = unet, controlnets, scheduler, layout, images
        self.base = DenoisingLoop(unet, controlnets, scheduler, guidance_scale=guidance_scale, use_lcm=use_lcm)
        self.use_cuda_graph, self._graph = use_cuda_graph, None
        b, _, f, hh, ww = latent_shape
        self.frames = f
        rows = (2 if self.base.do_cfg else 1) * f
        cn = controlnets.controlnets[0]
        boc, lpb = cn.config["block_out_channels"], cn.config["layers_per_block"]
        shapes, h, w = [(rows, boc[0], hh, ww)], hh, ww
        for i, c in enumerate(boc):
            shapes += [(rows, c, h, w)] * lpb
            if i != len(boc) - 1:
                h, w = (h + 1) // 2, (w + 1) // 2
                shapes.append((rows, c, h, w))
        shapes.append((rows, boc[-1], h, w))
        dtype, dev = unet.conv_in.weight.dtype, unet.conv_in.weight.device
        self.transport = ClipTransport(shapes, layout.max_slots(), dtype, dev, layout.n_windows, latent_shape, torch.float32, group)
        self.windows = WindowParallel(layout.rank, layout.n_windows, f, overlap, unet_group) if layout.is_unet_rank else None
        if layout.is_unet_rank:      # the arenas start out free: the servers' first wait_released must not block
            for s in layout.servers_of(layout.window):
                self.transport.release(s)

    # ---- UNet rank ---------------------------------------------------------------------------------------------------
    def _unet_noise(self, latents: torch.Tensor, t, prompt_embeds: torch.Tensor) -> torch.Tensor:
        lo, tr, mc, base = self.layout, self.transport, self.mc, self.base
        w, f = lo.window, latents.shape[2]
        for s in lo.servers_of(w):                                   # the servers start on this step's latents right away
            tr.send_latents(s, w, latents.float())
        cfg = base.do_cfg
        model_in = torch.cat([latents] * 2) if cfg else latents
        local = [k for _, k in lo.jobs[lo.rank]]
        imgs = [self.images.get((w, k)) for k in range(lo.n_nets)]
        mine = mc.raw(model_in, t, prompt_embeds, f, nets=local, images=imgs) if local else []
        per_net, owners = [], []
        for k in range(lo.n_nets):
            if k in local:
                per_net.append(mine[local.index(k)])
            else:
                per_net.append(tr.peer_views(lo.owner_of(w, k), lo.slot_of(w, k)))
                owners.append(lo.owner_of(w, k))
        down = mc.join(RemoteResidualSet(per_net, mc.cond_scale, f, False, tr, owners))
        noise = self.unet(model_in, t, encoder_hidden_states=prompt_embeds, down_block_additional_residuals=down,
                          timestep_cond=base._w_embedding(model_in)).sample.to(latents.dtype)
        if cfg:
            u, c = noise.chunk(2)
            noise = u + base.guidance_scale * (c - u)
        return noise

    # ---- server rank -------------------------------------------------------------------------------------------------
    def _serve(self, t, prompt_embeds) -> None:
        lo, tr, mc = self.layout, self.transport, self.mc
        cfg = self.base.do_cfg
        for w in lo.windows_of(lo.rank):
            latents = tr.recv_latents(w, w)                          # stream waits for UNet rank w's signal
            tr.wait_released(w)                                      # ... and until it has consumed last step's residuals
            model_in = torch.cat([latents] * 2) if cfg else latents
            jobs = [(k, slot) for slot, (jw, k) in enumerate(lo.jobs[lo.rank]) if jw == w]
            imgs = [self.images.get((w, k)) for k in range(lo.n_nets)]
            prompt = prompt_embeds[w] if isinstance(prompt_embeds, (list, tuple, dict)) else prompt_embeds
            mc.raw(model_in, t, prompt, self.frames, nets=[k for k, _ in jobs], images=imgs,
                   out=[tr.out_views(slot) for _, slot in jobs])
            mc.join()
            tr.publish(w)

    def _capture(self, fn, *static):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                                # warm-up (cuDNN algorithm selection, caches) — a full,
            for _ in range(2):                                       # collective step on every rank, signals included
                fn(*static)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = fn(*static)
        return graph, out

    @torch.no_grad()
    def step(self, latents: Optional[torch.Tensor], t: int, prompt_embeds) -> Optional[torch.Tensor]:
        lo = self.layout
        if not self.use_cuda_graph:
            if not lo.is_unet_rank:
                self._serve(t, prompt_embeds)
                return None
            noise = self._unet_noise(latents, t, prompt_embeds)
        else:
            key = (None if latents is None else (tuple(latents.shape), latents.dtype),
                   tuple(prompt_embeds.shape) if torch.is_tensor(prompt_embeds) else None)
            if self._graph is not None and self._graph["key"] != key:
                raise ValueError("ClipLoop: the captured step was recorded for other input shapes; build a new loop (the capture "
                                 "is collective: every rank would have to re-record at the same step)")
            if self.mc.hoist_cond_embedding:
                raise ValueError("ClipLoop: hoisted conditioning embeddings are not wired into the captured clip step "
                                 "(run it eagerly, or capture without them)")
            if self._graph is None:
                dev = self.unet.conv_in.weight.device
                s_t = torch.full((1,), int(t), dtype=torch.int64, device=dev)
                if lo.is_unet_rank:
                    s_lat, s_prompt = latents.clone(), prompt_embeds.clone()
                    graph, out = self._capture(self._unet_noise, s_lat, s_t, s_prompt)
                    self._graph = dict(graph=graph, lat=s_lat, t=s_t, prompt=s_prompt, out=out, key=key)
                else:
                    graph, _ = self._capture(self._serve, s_t, prompt_embeds)
                    self._graph = dict(graph=graph, t=s_t, key=key)
                # the warm-up / capture passes exchanged signals in lock step on every rank; nothing is left pending
            g = self._graph
            g["t"].fill_(int(t))
            if lo.is_unet_rank:
                g["lat"].copy_(latents, non_blocking=True)
                if prompt_embeds.data_ptr() != g["prompt"].data_ptr():
                    g["prompt"].copy_(prompt_embeds, non_blocking=True)
            g["graph"].replay()
            if not lo.is_unet_rank:
                return None
            noise = g["out"]
        latents = self.scheduler.step(noise, int(t), latents)
        return self.windows.exchange(latents)
