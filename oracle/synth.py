"""Deterministic synthetic weights and inputs (TEST INFRASTRUCTURE — see oracle/README.md).

No real checkpoints or videos exist offline, so every tensor on the hot path is generated from a
seed *and the tensor's own state_dict key*; the reference model (imported through the shim in
`oracle/gen_golden.py`), the oracle restatement (`oracle/ref_*.py`) and the CUDA product
(`controlanimate_b200/`) can therefore be filled with bit-identical fp32 weights without shipping a
state_dict.  numpy's PCG64 stream is used because it is stable across numpy versions/platforms.

Zero-initialised layers of the reference (`motion_module.py:76-77` proj_out; diffusers ControlNet
zero-convs) are re-randomised like every other weight — otherwise parity tests would be vacuous
(SURVEY.md §7 "Hard parts").
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, Tuple

import numpy as np
import torch


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.default_rng([int(seed) & 0x7FFFFFFF, zlib.crc32(key.encode())])


def tensor(seed: int, key: str, shape: Iterable[int], scale: float = 1.0, dtype=torch.float32) -> torch.Tensor:
    """N(0, scale²) tensor that depends only on (seed, key, shape)."""
    shape = tuple(int(s) for s in shape)
    a = _rng(seed, key).standard_normal(shape, dtype=np.float32) * np.float32(scale)
    return torch.from_numpy(a).to(dtype)


def _is_norm_key(key: str) -> bool:
    parts = key.split(".")
    owner = parts[-2] if len(parts) >= 2 else ""
    if owner.startswith("norm") or owner in ("ff_norm", "conv_norm_out", "norm_out"):
        return True
    # `norms.0.weight` (TemporalTransformerBlock.norms ModuleList, motion_module.py:203)
    return len(parts) >= 3 and parts[-3] == "norms"


def synth_value(seed: int, key: str, ref: torch.Tensor) -> torch.Tensor:
    """Synthetic value for state_dict entry `key` with the shape of `ref` (fp32)."""
    shape = tuple(ref.shape)
    leaf = key.rsplit(".", 1)[-1]
    if ref.ndim >= 2:
        fan_in = int(np.prod(shape[1:]))
        return tensor(seed, key, shape, 1.0 / math.sqrt(fan_in))
    if leaf == "weight" and _is_norm_key(key):
        return 1.0 + tensor(seed, key, shape, 0.1)
    return tensor(seed, key, shape, 0.1)  # biases (and any other 1-D parameter)


def fill_state_dict(sd: Dict[str, torch.Tensor], seed: int) -> Dict[str, torch.Tensor]:
    """Return a new fp32 state_dict with the same keys/shapes, synthetic values.

    Buffers named `...pos_encoder.pe` keep their analytic value (`motion_module.py:236-244`).
    """
    out = {}
    for k, v in sd.items():
        if k.endswith(".pe"):
            out[k] = v.detach().clone().float()
        else:
            out[k] = synth_value(seed, k, v)
    return out


def fill_module_(module: torch.nn.Module, seed: int) -> torch.nn.Module:
    sd = fill_state_dict(module.state_dict(), seed)
    with torch.no_grad():
        for k, v in module.state_dict().items():
            v.copy_(sd[k].to(v.dtype))
    return module


# Architecture kwargs = /root/reference/configs/inference/inference-v2.yaml:1-22 (mm_sd_v15_v2)
MOTION_MODULE_KWARGS_V2 = dict(
    num_attention_heads=8,
    num_transformer_block=1,
    attention_block_types=("Temporal_Self", "Temporal_Self"),
    temporal_position_encoding=True,
    temporal_position_encoding_max_len=32,
    temporal_attention_dim_div=1,
)

UNET_ADDITIONAL_KWARGS_V2 = dict(
    use_inflated_groupnorm=True,
    unet_use_cross_frame_attention=False,
    unet_use_temporal_attention=False,
    use_motion_module=True,
    motion_module_resolutions=(1, 2, 4, 8),
    motion_module_mid_block=True,
    motion_module_decoder_only=False,
    motion_module_type="Vanilla",
    motion_module_kwargs=MOTION_MODULE_KWARGS_V2,
)


def unet_config(tiny: bool = False) -> dict:
    """SD1.5 UNet3D (+v2 motion modules) ctor kwargs; `tiny` = same topology, 1/10 width."""
    cfg = dict(
        sample_size=64, in_channels=4, out_channels=4,
        down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
        up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
        block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, cross_attention_dim=768,
        attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5,
    )
    if tiny:
        cfg.update(block_out_channels=(32, 64, 128, 128), cross_attention_dim=64, sample_size=16)
    cfg.update(UNET_ADDITIONAL_KWARGS_V2)
    return cfg


# The 13 ControlNet residual shapes per (b·f) sample, as (channels, latent divisor)
# (order = UNet `down_block_res_samples`, unet.py:550-562, then mid).
def residual_shapes(block_out_channels=(320, 640, 1280, 1280), layers_per_block=2) -> Tuple[Tuple[int, int], ...]:
    shapes = [(block_out_channels[0], 1)]
    div = 1
    for i, c in enumerate(block_out_channels):
        shapes += [(c, div)] * layers_per_block
        if i != len(block_out_channels) - 1:
            div *= 2
            shapes.append((c, div))
    shapes.append((block_out_channels[-1], div))  # mid
    return tuple(shapes)
