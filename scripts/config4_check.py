#!/usr/bin/env python
"""BASELINE config 4 at full width on one B200 (no oracle at this size: finiteness, shape, determinism and cross-checks only):
UNet3D forward with f = 32 frames (PE max_len), non-square 96x54 latents (pyramid 54 -> 27 -> 14 -> 7: `forward_upsample_size`),
IP-Adapter-length prompts [b, 81, 768], plus the same step with the own GroupNorm / attention paths forced to their alternatives
through the environment of a child process (the noise predictions must agree to bf16 accuracy)."""
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlanimate_b200 import _lib, unet as un, utils  # noqa: E402
from oracle import synth  # noqa: E402  (configuration tables only)


def run():
    _lib.load(build_if_missing=False)
    dev = torch.device("cuda")
    cfg = synth.unet_config(tiny=False)
    net = utils.build_on_device(lambda: un.UNet3DConditionModel(**cfg), dev, torch.bfloat16, seed=1)
    g = torch.Generator(device="cpu").manual_seed(4)
    x = torch.randn(1, 4, 32, 54, 96, generator=g).to(dev, torch.bfloat16)
    ctx = torch.randn(1, 81, 768, generator=g).to(dev, torch.bfloat16)
    with torch.no_grad():
        y = net(x, 501, ctx).sample
        y2 = net(x, 501, ctx).sample
    torch.cuda.synchronize()
    assert y.shape == x.shape, y.shape
    assert torch.isfinite(y.float()).all()
    assert torch.equal(y, y2), "not deterministic"
    return y.float().cpu()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        torch.save(run(), sys.argv[2])
        sys.exit(0)
    base = run()
    print("config 4 forward ok:", tuple(base.shape), "rms", float(base.pow(2).mean().sqrt()))
    for tag, env in (("gn ring instead of slab", {"CA_GN_SLAB": "0"}), ("gn split launches", {"CA_GN_SLAB": "0", "CA_GN_RING": "0"}),
                     ("layernorm persistent", {"CA_LN_MODE": "persist"})):
        path = f"/tmp/config4_{abs(hash(tag))}.pt"
        subprocess.run([sys.executable, __file__, "child", path], env=dict(os.environ, **env), check=True)
        other = torch.load(path)
        cos = float(torch.dot(base.flatten().double(), other.flatten().double()) / (base.norm().double() * other.norm().double()))
        print(f"  vs {tag}: cosine {cos:.6f}")
        assert cos >= 0.999, (tag, cos)
    print("config 4 cross-checks ok")
