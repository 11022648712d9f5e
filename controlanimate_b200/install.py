"""Install the B200 kernels into an UNMODIFIED reference model (the drop-in route).

    from controlanimate_b200.install import install
    pipe = ControlAnimatePipeline(config)          # reference facade, modules/controlanimate_pipeline.py:26-121
    install(pipe.pipeline.unet)                    # AFTER load_weights / .half() / enable_xformers... (:88-116)

* B1: every VersatileAttention gets a `B200TemporalAttnProcessor` (via the module's own `set_processor`, so
  `unet.attn_processors` keeps its 90 entries).
* B2: every `block.motion_modules[i]` (VanillaTemporalModule) is replaced by a `B200MotionModule` carrying the same
  state_dict (removes the dead Q/K/V GEMMs of motion_module.py:299-311 and the six layout copies).
* B4: every `ResnetBlock3D` is replaced by a `B200ResnetBlock3D` (fused GroupNorm+SiLU(+temb)).

`enable_xformers_memory_efficient_attention()` (controlanimate_pipeline.py:112) and `set_ip_adapter`
(ip_adapter.py:95-126) overwrite processors: call `install` after them (`verify_installed` checks).
"""
from __future__ import annotations

from typing import Dict

import torch.nn as nn

from .layers import B200IPAttnProcessor, B200MotionModule, B200ResnetBlock3D, B200TemporalAttnProcessor


def _is_versatile_attention(m: nn.Module) -> bool:
    return getattr(m, "attention_mode", None) == "Temporal" and hasattr(m, "set_processor")


def _motion_kwargs(mm: nn.Module) -> dict:
    tt = mm.temporal_transformer
    blk = tt.transformer_blocks[0]
    attn = blk.attention_blocks[0]
    pe = getattr(attn, "pos_encoder", None)
    return dict(in_channels=tt.proj_in.in_features, num_attention_heads=attn.heads,
                num_transformer_block=len(tt.transformer_blocks),
                attention_block_types=("Temporal_Self",) * len(blk.attention_blocks),
                temporal_position_encoding=pe is not None,
                temporal_position_encoding_max_len=pe.pe.shape[1] if pe is not None else 24, zero_initialize=False)


def _convert_motion_module(old: nn.Module) -> B200MotionModule:
    new = B200MotionModule(**_motion_kwargs(old))
    p = next(old.parameters())
    new.load_state_dict(old.state_dict(), strict=True)
    return new.to(device=p.device, dtype=p.dtype).eval()


def _convert_resnet(old: nn.Module) -> B200ResnetBlock3D:
    per_frame = type(old.norm1).__name__ == "InflatedGroupNorm"
    new = B200ResnetBlock3D(in_channels=old.in_channels, out_channels=old.out_channels,
                            temb_channels=old.time_emb_proj.in_features if old.time_emb_proj is not None else None,
                            groups=old.norm1.num_groups, groups_out=old.norm2.num_groups, eps=old.norm1.eps,
                            output_scale_factor=old.output_scale_factor, use_in_shortcut=old.conv_shortcut is not None,
                            use_inflated_groupnorm=per_frame)
    if getattr(old, "time_embedding_norm", "default") != "default":
        raise ValueError("time_embedding_norm='scale_shift' is not used by the shipped configs")
    p = next(old.parameters())
    new.load_state_dict(old.state_dict(), strict=True)
    return new.to(device=p.device, dtype=p.dtype).eval()


def _controlnet_call(self, control_model_input, t, controlnet_prompt_embeds, frame_count, image_embeds=None,
                     do_classifier_free_guidance=True, guess_mode=True):
    """Replacement for `MultiControlNetResidualsPipeline.__call__` (reference modules/controlresiduals_pipeline.py:278-316).
    Same inputs; every ControlNet is evaluated with conditioning_scale 1 (RAW residuals) and the scale-and-sum of diffusers'
    MultiControlNetModel, the 13 rearranges of :304-312 and the UNet's skip adds are left to kernel (3): the return value is
    a tuple of `LazyResidual` proxies that the UNMODIFIED reference UNet consumes through `skip + residual`."""
    import torch
    from .residuals import ResidualSet
    b, c, f, h, w = control_model_input.shape
    x = control_model_input.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)                      # :287
    prompt = torch.cat([controlnet_prompt_embeds] * frame_count)                                # :292 (tiling kept as is)
    nets = list(getattr(self.controlnet, "nets", [self.controlnet]))
    dtype = next(nets[0].parameters()).dtype if hasattr(nets[0], "parameters") and any(True for _ in nets[0].parameters()) else x.dtype
    images = self.prep_images if isinstance(self.prep_images, (list, tuple)) else [self.prep_images]
    per_net = []
    for k, net in enumerate(nets):
        down, mid = net(x.to(dtype), t, encoder_hidden_states=prompt.to(dtype), controlnet_cond=images[k], conditioning_scale=1.0,
                        guess_mode=False, return_dict=False)                                    # :294-302, unscaled
        per_net.append(list(down) + [mid])
    scales = self.cond_scale if isinstance(self.cond_scale, (list, tuple)) else [self.cond_scale] * len(nets)
    return ResidualSet(per_net, list(scales), frame_count, bool(guess_mode)).lazy_tuple()


def install_controlnet_pipeline(pipe) -> None:
    """Boundary B3 producer side: make `pipe(...)` (a MultiControlNetResidualsPipeline) return lazy residual sets."""
    cls = type(pipe)
    if getattr(cls, "_ca_b200", False):
        return
    pipe.__class__ = type("B200" + cls.__name__, (cls,), {"__call__": _controlnet_call, "_ca_b200": True})


def install(unet: nn.Module, controlnet_pipeline=None, *, processors: bool = True, motion_modules: bool = True,
            resnets: bool = True) -> Dict[str, int]:
    """Mutates `unet` (and, when given, `controlnet_pipeline`) in place; returns how many modules of each kind were
    replaced.  `install(pipe.pipeline.unet, pipe.pipeline.controlnet)` covers boundaries B1-B4."""
    counts = dict(processors=0, motion_modules=0, resnets=0, controlnet_pipeline=0)
    if controlnet_pipeline is not None:
        install_controlnet_pipeline(controlnet_pipeline)
        counts["controlnet_pipeline"] = 1
    if motion_modules or resnets:
        for parent in list(unet.modules()):
            for name, child in list(parent.named_children()):
                cls = type(child).__name__
                if motion_modules and cls == "VanillaTemporalModule":
                    setattr(parent, name, _convert_motion_module(child))
                    counts["motion_modules"] += 1
                elif resnets and cls == "ResnetBlock3D":
                    setattr(parent, name, _convert_resnet(child))
                    counts["resnets"] += 1
            if isinstance(parent, nn.ModuleList):
                for i, child in enumerate(list(parent)):
                    cls = type(child).__name__
                    if motion_modules and cls == "VanillaTemporalModule":
                        parent[i] = _convert_motion_module(child)
                        counts["motion_modules"] += 1
                    elif resnets and cls == "ResnetBlock3D":
                        parent[i] = _convert_resnet(child)
                        counts["resnets"] += 1
    if processors:
        for m in unet.modules():
            if _is_versatile_attention(m):
                m.set_processor(B200TemporalAttnProcessor())
                counts["processors"] += 1
    return counts


def verify_installed(unet: nn.Module) -> bool:
    """True iff every temporal attention still dispatches to a B200 processor (nothing overwrote them)."""
    mods = [m for m in unet.modules() if _is_versatile_attention(m)]
    return bool(mods) and all(isinstance(m.get_processor(), B200TemporalAttnProcessor) for m in mods)


def set_ip_adapter(unet: nn.Module, scale: float = 1.0, num_tokens: int = 4) -> Dict[str, nn.Module]:
    """Mirror of `IPAdapter.set_ip_adapter` (reference modules/ip_adapter.py:95-126) with B200 processors: every
    cross-attention (`attn2`) gets a `B200IPAttnProcessor` (dual-KV kernel path), every other attention keeps what it has
    (the reference installs a plain AttnProcessor2_0 there, which the B200 temporal processor is equivalent to).
    Returns {attn_processors key: processor} of the installed IP processors, in `attn_processors` order, so that
    `torch.nn.ModuleList(procs.values())` can be handed to the reference's weight loader (ip_adapter.py:184)."""
    cross_dim = unet.config["cross_attention_dim"] if isinstance(unet.config, dict) else unet.config.cross_attention_dim
    procs = {}
    for name, m in unet.named_modules():
        if name.endswith("attn2") and hasattr(m, "set_processor"):
            hidden = m.to_q.weight.shape[0]
            p = B200IPAttnProcessor(hidden_size=hidden, cross_attention_dim=cross_dim, scale=scale, num_tokens=num_tokens)
            p = p.to(device=m.to_q.weight.device, dtype=m.to_q.weight.dtype)
            m.set_processor(p)
            procs[f"{name}.processor"] = p
    return procs
