"""Generate tests/golden/*.npz by running the REFERENCE'S OWN PYTHON (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python -m oracle.gen_golden            # from the repo root

The reference modules are imported UNMODIFIED from /root/reference through `oracle/diffusers_shim`
(diffusers/controlnet_aux are not installed and there is no network).  Weights and inputs come from
`oracle.synth` keyed by state_dict name, so tests regenerate identical inputs from the seeds stored
in each fixture; only outputs (fp32) are stored.  Attention arithmetic comes from the reference's
own `modules/attention_processor.py` processors (math path :7-77 and SDPA path :186-272).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = os.environ.get("CA_REFERENCE_ROOT", "/root/reference")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    if not os.path.isdir(REFERENCE):
        raise SystemExit(f"{REFERENCE} not found: golden generation only runs in the build container")
    sys.path.insert(0, os.path.join(HERE, "diffusers_shim"))
    sys.path.insert(0, REFERENCE)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)


def _save(name, **arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().float().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def _set_processors(module, proc):
    for m in module.modules():
        if hasattr(m, "set_processor"):
            m.set_processor(proc)


def main():
    _import_reference()
    from oracle import synth
    from animatediff.models.resnet import InflatedGroupNorm, ResnetBlock3D
    from animatediff.models.motion_module import VanillaTemporalModule
    from animatediff.models.unet import UNet3DConditionModel
    from modules.attention_processor import AttnProcessor, AttnProcessor2_0
    import torch.nn.functional as F

    torch.set_grad_enabled(False)
    SEED = 1234

    # ---- (2) GroupNorm + SiLU: resnet.py:23-31 + :191-192; v1 (non-inflated) variant :150-151 --------------
    b, c, f, h, w = 2, 64, 3, 6, 5
    x = synth.tensor(SEED, "gn.x", (b, c, f, h, w)) * 1.7 + 0.3
    temb = synth.tensor(SEED, "gn.temb", (b, c))
    out = {}
    for groups in (32, 8):
        for per_frame in (True, False):
            norm = (InflatedGroupNorm if per_frame else torch.nn.GroupNorm)(num_groups=groups, num_channels=c, eps=1e-5, affine=True)
            synth.fill_module_(norm, SEED)
            gamma, beta = norm.weight.clone(), norm.bias.clone()
            out[f"y_g{groups}_pf{int(per_frame)}"] = F.silu(norm(x))
            out[f"y_g{groups}_pf{int(per_frame)}_temb"] = F.silu(norm(x + temb[:, :, None, None, None]))
            out[f"y_g{groups}_pf{int(per_frame)}_nosilu"] = norm(x)
    _save("groupnorm_silu", seed=SEED, shape=np.array([b, c, f, h, w]), **out)

    # ---- ResnetBlock3D.forward resnet.py:188-218 ---------------------------------------------------------
    out = {}
    for name, (cin, cout) in {"same": (64, 64), "widen": (96, 64)}.items():
        for per_frame in (True, False):
            blk = ResnetBlock3D(in_channels=cin, out_channels=cout, temb_channels=128, groups=32, eps=1e-5,
                                non_linearity="silu", use_inflated_groupnorm=per_frame)
            synth.fill_module_(blk, SEED)
            xin = synth.tensor(SEED, f"resnet.{name}.x", (2, cin, 3, 6, 5))
            te = synth.tensor(SEED, f"resnet.{name}.temb", (2, 128))
            out[f"{name}_pf{int(per_frame)}"] = blk(xin, te)
    _save("resnet_block3d", seed=SEED, **out)

    # ---- (1) motion module: motion_module.py:79-160, :212-224, :272-329 -----------------------------------
    out = {}
    for cname, (c, f, h, w) in {"c64_f8": (64, 8, 4, 3), "c128_f16": (128, 16, 3, 3), "c64_f5": (64, 5, 2, 2)}.items():
        mm = VanillaTemporalModule(in_channels=c, **synth.MOTION_MODULE_KWARGS_V2)
        synth.fill_module_(mm, SEED)
        xin = synth.tensor(SEED, f"mm.{cname}.x", (2, c, f, h, w))
        _set_processors(mm, AttnProcessor())
        y_math = mm(xin, None, None)
        _set_processors(mm, AttnProcessor2_0())
        y_sdpa = mm(xin, None, None)
        assert torch.allclose(y_math, y_sdpa, atol=2e-5, rtol=1e-5), (y_math - y_sdpa).abs().max()
        out[cname] = y_math
        # the AttentionProcessor boundary alone (B1): processor(attn, hidden_states) on [(b d), f, C]
        attn = mm.temporal_transformer.transformer_blocks[0].attention_blocks[0]
        xa = synth.tensor(SEED, f"mm.{cname}.proc_x", (6, f, c))
        out[cname + "_proc"] = AttnProcessor()(attn, xa)
        # VersatileAttention.forward incl. rearranges + PE (motion_module.py:272-329)
        xv = synth.tensor(SEED, f"mm.{cname}.va_x", (2 * f, h * w, c))
        attn.set_processor(AttnProcessor())
        out[cname + "_va"] = attn(xv, video_length=f)
    _save("motion_module", seed=SEED, **out)

    # ---- whole UNet3D forward unet.py:458-621 (tiny width, full topology) ---------------------------------
    cfg = synth.unet_config(tiny=True)
    unet = UNet3DConditionModel(**cfg)
    synth.fill_module_(unet, SEED)
    unet.set_attn_processor(AttnProcessor())
    out = {}
    n_proc = len(unet.attn_processors)
    for cname, (b, f, hh, ww) in {"sq": (2, 4, 16, 16), "odd": (1, 3, 12, 10)}.items():
        sample = synth.tensor(SEED, f"unet.{cname}.sample", (b, 4, f, hh, ww))
        ctx = synth.tensor(SEED, f"unet.{cname}.ctx", (b, 7, cfg["cross_attention_dim"]))
        res = []
        sh, sw = hh, ww
        div_prev = 1
        for i, (ch, div) in enumerate(synth.residual_shapes(cfg["block_out_channels"])):
            while div_prev < div:
                sh, sw = (sh + 1) // 2, (sw + 1) // 2
                div_prev *= 2
            res.append(synth.tensor(SEED, f"unet.{cname}.res{i}", (b, ch, f, sh, sw), 0.1))
        t = 501
        out[cname + "_plain"] = unet(sample, t, encoder_hidden_states=ctx).sample
        out[cname + "_ctrl"] = unet(sample, t, encoder_hidden_states=ctx, down_block_additional_residuals=tuple(res[:-1]),
                                    mid_block_additional_residual=res[-1]).sample
    # guess-mode + CFG broadcast: residual batch 1 against UNet batch 2 (unet.py:572)
    sample = synth.tensor(SEED, "unet.sq.sample", (2, 4, 4, 16, 16))
    ctx = synth.tensor(SEED, "unet.sq.ctx", (2, 7, cfg["cross_attention_dim"]))
    res1 = []
    sh = 16
    div_prev = 1
    for i, (ch, div) in enumerate(synth.residual_shapes(cfg["block_out_channels"])):
        while div_prev < div:
            sh = (sh + 1) // 2
            div_prev *= 2
        res1.append(synth.tensor(SEED, f"unet.sq.res{i}", (2, ch, 4, sh, sh), 0.1)[:1])
    out["sq_ctrl_bcast"] = unet(sample, 501, encoder_hidden_states=ctx, down_block_additional_residuals=tuple(res1[:-1]),
                                mid_block_additional_residual=res1[-1]).sample
    _save("unet3d_tiny", seed=SEED, n_attn_processors=n_proc, **out)

    # ---- (3) residual layout contract: controlresiduals_pipeline.py:278-316 with a recorded ControlNet -----
    from modules.controlresiduals_pipeline import MultiControlNetResidualsPipeline

    class RecordedControlNet:
        """Stands in for diffusers' MultiControlNetModel: returns pre-made (b f)-batched residuals."""

        def __init__(self, down, mid):
            self.down, self.mid, self.calls = down, mid, []

        def __call__(self, sample, t, encoder_hidden_states=None, controlnet_cond=None, conditioning_scale=None,
                     guess_mode=None, return_dict=False):
            self.calls.append(dict(sample=tuple(sample.shape), ehs=tuple(encoder_hidden_states.shape),
                                   scale=list(conditioning_scale), guess_mode=guess_mode))
            return self.down, self.mid

    b, f = 2, 3
    raw = [synth.tensor(SEED, f"cn.res{i}", (b * f, ch, max(8 // div, 1), max(8 // div, 1)))
           for i, (ch, div) in enumerate(synth.residual_shapes((32, 64, 128, 128)))]
    pipe = object.__new__(MultiControlNetResidualsPipeline)
    pipe.controlnet = RecordedControlNet(raw[:-1], raw[-1])
    pipe.prep_images, pipe.cond_scale = None, [1.0, 0.5]
    down, mid = pipe(torch.zeros(b, 4, f, 8, 8), 501, torch.zeros(b, 7, 64), f, guess_mode=False)
    call = pipe.controlnet.calls[0]
    assert call["sample"] == (b * f, 4, 8, 8) and call["ehs"] == (b * f, 7, 64)
    out = {f"down{i}": d for i, d in enumerate(down)}
    out["mid"] = mid
    _save("residual_layout", seed=SEED, b=b, f=f, **out)
    gen_ip_adapter()


def gen_ip_adapter():
    """IP-Adapter dual-KV cross-attention: the reference's own IPAttnProcessor2_0 / IPAttnProcessor
    (modules/attention_processor.py:367-492, :80-183) driving the shim's Attention module (config 4: 77 + 4 tokens)."""
    _import_reference()
    from oracle import synth
    from diffusers.models.attention_processor import Attention
    from modules.attention_processor import IPAttnProcessor, IPAttnProcessor2_0
    torch.set_grad_enabled(False)
    SEED = 4321
    out = {}
    for cname, (c, cross, heads, n, d, L, ntok, scale) in {"c64": (64, 48, 8, 3, 20, 11, 4, 1.0), "c320": (320, 768, 8, 2, 12, 81, 4, 0.6)}.items():
        attn = Attention(query_dim=c, cross_attention_dim=cross, heads=heads, dim_head=c // heads)
        synth.fill_module_(attn, SEED)
        proc = IPAttnProcessor2_0(hidden_size=c, cross_attention_dim=cross, scale=scale, num_tokens=ntok)
        synth.fill_module_(proc, SEED + 1)
        x = synth.tensor(SEED, f"ip.{cname}.x", (n, d, c))
        ctx = synth.tensor(SEED, f"ip.{cname}.ctx", (n, L, cross))
        y = proc(attn, x, encoder_hidden_states=ctx)
        proc_math = IPAttnProcessor(hidden_size=c, cross_attention_dim=cross, scale=scale, num_tokens=ntok)
        proc_math.load_state_dict(proc.state_dict())
        y_math = proc_math(attn, x, encoder_hidden_states=ctx)
        assert torch.allclose(y, y_math, atol=2e-5, rtol=1e-5), (y - y_math).abs().max()
        out[cname] = y
    _save("ip_adapter", seed=SEED, **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ip":
        gen_ip_adapter()
    else:
        main()
