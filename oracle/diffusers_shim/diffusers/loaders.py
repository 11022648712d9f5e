class UNet2DConditionLoadersMixin:
    """Empty: LoRA / attn-proc loading is off the hot path."""
