"""diffusers.models.attention symbols: Attention re-export, FeedForward, AdaLayerNorm (name only)."""
from torch import nn

from .activations import GEGLU
from .attention_processor import Attention  # noqa: F401
from .lora import LoRACompatibleLinear


class AdaLayerNorm(nn.Module):  # name only
    pass


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False):
        super().__init__()
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        assert activation_fn == "geglu"
        self.net = nn.ModuleList([GEGLU(dim, inner_dim), nn.Dropout(dropout), LoRACompatibleLinear(inner_dim, dim_out)])

    def forward(self, hidden_states, scale: float = 1.0):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states
