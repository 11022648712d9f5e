"""Model construction helpers: fast random initialisation on the device (there are no checkpoints offline)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn


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
