"""GEGLU restated (diffusers 0.23.0 models/activations.py): proj -> chunk -> h * gelu_erf(gate)."""
import torch.nn.functional as F
from torch import nn

from .lora import LoRACompatibleLinear


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = LoRACompatibleLinear(dim_in, dim_out * 2)

    def forward(self, hidden_states, scale: float = 1.0):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class GELU(nn.Module):  # name only (unused by the v2 configs)
    pass


class ApproximateGELU(nn.Module):  # name only
    pass
