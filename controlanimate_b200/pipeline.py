"""Loop-level host code of the hot path: the per-step body of
ControlAnimationPipeline.__call__ (reference animatediff/pipelines/controlanimation_pipeline.py:793-849) and
MultiControlNetResidualsPipeline.__call__ (modules/controlresiduals_pipeline.py:278-316), on the B200 modules.

Out of scope here: prompt encoding, VAE, annotators, video I/O.  Inputs are prompt embeddings,
latents and already-prepared control images.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .residuals import ResidualSet, merge_controlnet_residuals
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

    def raw(self, control_model_input, t, controlnet_prompt_embeds, frame_count, nets=None):
        b = control_model_input.shape[0]
        x = control_model_input.permute(0, 2, 1, 3, 4).reshape(b * frame_count, *control_model_input.shape[1:2],
                                                               *control_model_input.shape[3:])   # :287
        # :292 tiles the prompt embeddings as cat([e]*f): row n of the tiled tensor is e[n % b] (kept as-is)
        n_prompts = controlnet_prompt_embeds.shape[0]
        ctx_map = torch.arange(b * frame_count, device=x.device) % n_prompts
        nets = range(len(self.controlnets)) if nets is None else nets
        return [self.controlnets[k](x, t, controlnet_prompt_embeds, self.prep_images[k], ctx_map=ctx_map) for k in nets]

    def __call__(self, control_model_input, t, controlnet_prompt_embeds, frame_count, image_embeds=None,
                 do_classifier_free_guidance=True, guess_mode=True):
        per_net = self.raw(control_model_input, t, controlnet_prompt_embeds, frame_count)
        if self.lazy:
            rs = ResidualSet(per_net, self.cond_scale, frame_count, guess_mode)
            return rs, None
        return merge_controlnet_residuals(per_net, self.cond_scale, frame_count, guess_mode)     # :294-316


class DenoisingLoop:
    """The hot loop (controlanimation_pipeline.py:790-856) for the non-LCM path: CFG-duplicated latents, ControlNets,
    UNet3D, `eps = eps_u + g (eps_c - eps_u)` (:845-846), scheduler step (:849).  No per-step empty_cache().

    use_cuda_graph: the noise prediction of a step (ControlNets -> kernel-(3) merge -> UNet3D -> CFG; ~5000 launches) is
    captured once per input signature into a CUDA graph and replayed, so the step is bound by the GPU, not by Python.
    """

    def __init__(self, unet: UNet3DConditionModel, controlnets: Optional[MultiControlNetResiduals], scheduler: DDIMScheduler,
                 guidance_scale: float = 7.5, guess_mode: bool = False, use_cuda_graph: bool = False):
        self.unet, self.controlnets, self.scheduler = unet, controlnets, scheduler
        self.guidance_scale, self.guess_mode = guidance_scale, guess_mode
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}

    @property
    def do_cfg(self):
        return self.guidance_scale > 1.0

    def predict_noise(self, latents: torch.Tensor, t, prompt_embeds: torch.Tensor) -> torch.Tensor:
        """Guided noise prediction for latents [1,4,f,h,w]; `t` is an int or a 1-element int64 device tensor."""
        f = latents.shape[2]
        cfg = self.do_cfg
        model_in = torch.cat([latents] * 2) if cfg else latents                                          # :797
        down = mid = None
        if self.controlnets is not None:                                                                 # :807-819
            single = (self.guess_mode or not cfg) and cfg
            down, mid = self.controlnets(model_in[-1:] if single else model_in, t,
                                         prompt_embeds[-1:] if single else prompt_embeds, f,
                                         do_classifier_free_guidance=cfg, guess_mode=self.guess_mode)
        noise = self.unet(model_in, t, encoder_hidden_states=prompt_embeds, down_block_additional_residuals=down,
                          mid_block_additional_residual=mid).sample.to(latents.dtype)                   # :836-841
        if cfg:
            u, c = noise.chunk(2)
            noise = u + self.guidance_scale * (c - u)                                                    # :845-846
        return noise

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

    def _graphed_noise(self, latents, t: int, prompt_embeds):
        mc = self.controlnets
        images = list(mc.prep_images) if mc is not None and mc.prep_images is not None else []
        # everything the captured kernels read by value or by baked-in pointer is part of the key: shapes, the scalars of
        # the step, and the weights; the per-window control images are copied into graph-owned buffers before a replay
        key = (tuple(latents.shape), latents.dtype, tuple(prompt_embeds.shape), prompt_embeds.dtype, latents.device,
               float(self.guidance_scale), bool(self.guess_mode), tuple(mc.cond_scale) if mc is not None else (),
               bool(mc.lazy) if mc is not None else None, tuple((tuple(im.shape), im.dtype) for im in images),
               self._weight_stamp())
        g = self._graphs.get(key)
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
                with torch.cuda.stream(side):                # warm-up: cuDNN algorithm selection, caches, workspaces
                    for _ in range(2):
                        self.predict_noise(s_lat, s_t, s_prompt)
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    s_out = self.predict_noise(s_lat, s_t, s_prompt)
            finally:
                if mc is not None:
                    mc.prep_images = images if images else None
            g = self._graphs[key] = dict(graph=graph, lat=s_lat, t=s_t, prompt=s_prompt, out=s_out, images=s_images,
                                         seen=[None] * len(s_images))
        g["lat"].copy_(latents, non_blocking=True)
        if prompt_embeds.data_ptr() != g["prompt"].data_ptr():
            g["prompt"].copy_(prompt_embeds, non_blocking=True)
        for k, im in enumerate(images):      # the reference re-runs prep_control_images for every window (pipeline :226-273)
            ident = (im.data_ptr(), im._version)
            if g["seen"][k] != ident:
                g["images"][k].copy_(im, non_blocking=True)
                g["seen"][k] = ident
        g["t"].fill_(int(t))
        g["graph"].replay()
        return g["out"]

    @torch.no_grad()
    def step(self, latents: torch.Tensor, t: int, prompt_embeds: torch.Tensor) -> torch.Tensor:
        """latents [1,4,f,h,w] (fp32 or model dtype), prompt_embeds [2,L,D] = [negative, positive] when CFG."""
        if self.use_cuda_graph:
            noise = self._graphed_noise(latents, t, prompt_embeds)
        else:
            noise = self.predict_noise(latents, t, prompt_embeds)
        return self.scheduler.step(noise, int(t), latents)                                               # :849

    @torch.no_grad()
    def run(self, latents, prompt_embeds, num_inference_steps: int):
        for t in self.scheduler.set_timesteps(num_inference_steps):
            latents = self.step(latents, t, prompt_embeds)
        return latents
