"""`Attention` module restated (diffusers 0.23.0 models/attention_processor.py).

Only the constructor surface and helpers the reference touches.  The arithmetic of the processors
themselves is NOT restated here: golden generation installs the reference's own
`/root/reference/modules/attention_processor.py` processors.
"""
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .lora import LoRACompatibleLinear


class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 dropout: float = 0.0, bias=False, upcast_attention: bool = False, upcast_softmax: bool = False,
                 cross_attention_norm=None, cross_attention_norm_num_groups=32, added_kv_proj_dim=None,
                 norm_num_groups=None, spatial_norm_dim=None, out_bias: bool = True, scale_qk: bool = True,
                 only_cross_attention: bool = False, eps: float = 1e-5, rescale_output_factor: float = 1.0,
                 residual_connection: bool = False, _from_deprecated_attn_block=False, processor=None):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention = upcast_attention
        self.upcast_softmax = upcast_softmax
        self.rescale_output_factor = rescale_output_factor
        self.residual_connection = residual_connection
        self.dropout = dropout
        self.scale_qk = scale_qk
        self.scale = dim_head ** -0.5 if scale_qk else 1.0
        self.heads = heads
        self.sliceable_head_dim = heads
        self.added_kv_proj_dim = added_kv_proj_dim
        self.only_cross_attention = only_cross_attention
        self.group_norm = None
        self.spatial_norm = None
        self.norm_cross = None
        self.to_q = LoRACompatibleLinear(query_dim, self.inner_dim, bias=bias)
        self.to_k = LoRACompatibleLinear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = LoRACompatibleLinear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([LoRACompatibleLinear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.set_processor(processor if processor is not None else AttnProcessor())

    def set_use_memory_efficient_attention_xformers(self, *a, **k):  # pragma: no cover
        raise RuntimeError("xformers is not available in the shim")

    def set_processor(self, processor, _remove_lora=False):
        if hasattr(self, "processor") and isinstance(self.processor, nn.Module) and not isinstance(processor, nn.Module):
            self._modules.pop("processor")
        self.processor = processor

    def get_processor(self, return_deprecated_lora: bool = False):
        return self.processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)

    def batch_to_head_dim(self, tensor):
        head_size = self.heads
        batch_size, seq_len, dim = tensor.shape
        tensor = tensor.reshape(batch_size // head_size, head_size, seq_len, dim)
        return tensor.permute(0, 2, 1, 3).reshape(batch_size // head_size, seq_len, dim * head_size)

    def head_to_batch_dim(self, tensor, out_dim=3):
        head_size = self.heads
        batch_size, seq_len, dim = tensor.shape
        tensor = tensor.reshape(batch_size, seq_len, head_size, dim // head_size).permute(0, 2, 1, 3)
        if out_dim == 3:
            tensor = tensor.reshape(batch_size * head_size, seq_len, dim // head_size)
        return tensor

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        if attention_mask is None:
            baddbmm_input = torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device)
            beta = 0
        else:
            baddbmm_input, beta = attention_mask, 1
        scores = torch.baddbmm(baddbmm_input, query, key.transpose(-1, -2), beta=beta, alpha=self.scale)
        if self.upcast_softmax:
            scores = scores.float()
        return scores.softmax(dim=-1).to(dtype)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        if attention_mask is None:
            return attention_mask
        raise NotImplementedError("attention masks are not used on the hot path")


class AttnProcessor:
    """Placeholder default; golden generation always installs the reference's own processors."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        raise RuntimeError("install /root/reference/modules/attention_processor.py processors first")


class AttnProcessor2_0(AttnProcessor):
    pass


class AttnAddedKVProcessor(AttnProcessor):
    pass


AttentionProcessor = AttnProcessor
ADDED_KV_ATTENTION_PROCESSORS = (AttnAddedKVProcessor,)
CROSS_ATTENTION_PROCESSORS = (AttnProcessor, AttnProcessor2_0)
