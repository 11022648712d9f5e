"""Model construction helpers: fast random initialisation on the device (there are no checkpoints offline)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn


# Architecture of the headline workload: SD1.5 UNet3D + mm_sd_v15_v2 motion modules (reference
# configs/inference/inference-v2.yaml:1-22 on top of the SD1.5 UNet config).
MOTION_MODULE_KWARGS_V2 = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
                               temporal_position_encoding=True, temporal_position_encoding_max_len=32, temporal_attention_dim_div=1)


def sd15_unet3d_config(time_cond_proj_dim=None) -> dict:
    """Ctor kwargs of `UNet3DConditionModel` for SD1.5 + v2 motion modules (time_cond_proj_dim=256 for LCM checkpoints)."""
    cfg = dict(sample_size=64, in_channels=4, out_channels=4,
               down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
               up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
               block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, cross_attention_dim=768, attention_head_dim=8,
               norm_num_groups=32, norm_eps=1e-5, use_inflated_groupnorm=True, unet_use_cross_frame_attention=False,
               unet_use_temporal_attention=False, use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8),
               motion_module_mid_block=True, motion_module_decoder_only=False, motion_module_type="Vanilla",
               motion_module_kwargs=dict(MOTION_MODULE_KWARGS_V2))
    if time_cond_proj_dim:
        cfg["time_cond_proj_dim"] = time_cond_proj_dim
    return cfg


def analytic_pe(max_len: int, dim: int, device=None) -> torch.Tensor:
    """PositionalEncoding buffer, reference motion_module.py:236-244."""
    position = torch.arange(max_len, device=device).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, device=device) * (-math.log(10000.0) / dim))
    pe = torch.zeros(1, max_len, dim, device=device)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


@torch.no_grad()
def random_init_(module: nn.Module, seed: int = 0) -> nn.Module:
    """N(0, 1/fan_in) weights, 1+0.1 N norm gains, 0.1 N biases, analytic PE — generated on the module's device.
    (Zero-initialised layers of the reference are re-randomised so the benchmark does real work.)"""
    dev = next(module.parameters()).device
    gen = torch.Generator(device=dev).manual_seed(seed)
    for name, p in module.state_dict().items():
        if name.endswith(".pe"):
            p.copy_(analytic_pe(p.shape[1], p.shape[2], device=p.device))
            continue
        leaf = name.rsplit(".", 1)[-1]
        owner = name.split(".")[-2] if "." in name else ""
        if p.dim() >= 2:
            fan_in = p[0].numel()
            p.copy_(torch.randn(p.shape, generator=gen, device=dev, dtype=torch.float32) * fan_in ** -0.5)
        elif leaf == "weight" and ("norm" in owner or owner.isdigit()):
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=gen, device=dev, dtype=torch.float32))
        else:
            p.copy_(0.1 * torch.randn(p.shape, generator=gen, device=dev, dtype=torch.float32))
    return module


def build_on_device(ctor, device, dtype, seed: int = 0) -> nn.Module:
    """Construct `ctor()` on the meta device, materialise on `device` in `dtype`, random-init there."""
    with torch.device("meta"):
        m = ctor()
    m = m.to_empty(device=device)
    for mod in m.modules():  # PE tables stay fp32 (layers._PositionalEncoding._apply)
        pass
    m = m.to(dtype)
    random_init_(m, seed)
    return m.eval()
