"""Minimal stand-in for `diffusers==0.23.0` (reference pin: /root/reference/env.yml:120).

TEST INFRASTRUCTURE ONLY.  diffusers is not installed in this image and there is no network, so
this package supplies just the symbols `/root/reference/animatediff/models/*.py` import, restated
from the published diffusers 0.23.0 semantics, so that the reference's own Python can be imported
UNMODIFIED by `oracle/gen_golden.py` to produce the fixtures under `tests/golden/`.
Nothing under `controlanimate_b200/` may import this package.
"""
__version__ = "0.23.0+shim"


class ControlNetModel:  # name only (third-party model; restated in oracle/ref_unet3d.py, parity unpinned)
    @classmethod
    def from_pretrained(cls, *a, **k):
        raise RuntimeError("no checkpoints offline")
