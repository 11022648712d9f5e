"""controlanimate_b200 — B200-native (sm_100a) kernels for ControlAnimate's denoising hot path.

Host side mirrors the reference's operator surface (AttentionProcessor, motion-module forward,
ControlNet residual contract); arithmetic runs in hand-written CUDA behind the C ABI of
include/controlanimate_b200.h.  See DESIGN.md / INTEGRATION.md.
"""
__version__ = "0.1.0"
