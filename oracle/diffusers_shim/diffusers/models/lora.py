"""LoRACompatible* restated: plain Conv2d/Linear whose forward takes an ignored `scale` when no
LoRA layer is attached (diffusers 0.23.0 models/lora.py)."""
import torch.nn as nn
import torch.nn.functional as F


class LoRALinearLayer(nn.Module):  # name only
    pass


class LoRACompatibleConv(nn.Conv2d):
    def __init__(self, *args, lora_layer=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer = lora_layer

    def forward(self, hidden_states, scale: float = 1.0):
        return F.conv2d(hidden_states, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


class LoRACompatibleLinear(nn.Linear):
    def __init__(self, *args, lora_layer=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer = lora_layer

    def forward(self, hidden_states, scale: float = 1.0):
        return F.linear(hidden_states, self.weight, self.bias)
