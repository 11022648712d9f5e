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

from .layers import B200MotionModule, B200ResnetBlock3D, B200TemporalAttnProcessor


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


def install(unet: nn.Module, *, processors: bool = True, motion_modules: bool = True, resnets: bool = True) -> Dict[str, int]:
    """Mutates `unet` in place; returns how many modules of each kind were replaced."""
    counts = dict(processors=0, motion_modules=0, resnets=0)
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
