#!/usr/bin/env python
"""Headline benchmark: denoised frames/s of the ControlAnimate hot loop on B200 (BASELINE.json configs[1]).

Workload (config 2): SD1.5 UNet3D + mm_sd_v15_v2 motion modules + 2 ControlNets, 16 frames 512x512 (latent 64x64),
CFG 7.5 (b=2), DDIM 20 steps, bf16 storage / fp32 accumulate, random-init weights, synthetic latents / prompt
embeddings / control images.  One bench "step" = ONE denoising step of the loop (ControlNets -> residual merge ->
UNet3D -> CFG -> DDIM, reference controlanimation_pipeline.py:793-849); frames/s = frames / (20 * step time).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework (sm_100a kernels)
  python bench.py --impl reference ...                           # CPU arm: the oracle restatement of the reference's
                                                                 # own PyTorch path on the host cores (bounded sample)
N > 1 (torchrun), --parallelism:
  windows (default)  every rank denoises its own 16-frame window of one long clip (windows overlap by 4 latent frames,
                     exchanged and blended each step with NCCL send/recv) -> weak scaling; value counts the UNIQUE frames of the
                     clip (N * 16 - (N - 1) * 4), `frames_counted_per_s` also the overlap frames every neighbour denoises twice
  cfg                N = 2: one window, each rank one CFG row (strong scaling: step latency)
  cfg+controlnet     N = 4, 6: CFG halves x (UNet rank + ControlNet ranks); the ControlNet ranks' raw residuals are read over
                     NVLink by kernel (3) on the UNet rank (symmetric memory), overlapping the UNet's down path
  controlnet         N = 2, 3: UNet rank + ControlNet ranks (no CFG split)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DDIM_STEPS = 20
FRAMES = 16
LATENT = 64
GUIDANCE = 7.5
COND_SCALE = [1.0, 0.5]
OVERLAP = 4
# --workload config3 (BASELINE configs[2]): LCM-LoRA branch (b = 1, guidance through timestep_cond), 4 ControlNets with
# SampleConfig.yaml's scales, 4 steps, 16-frame windows with 4-frame overlap (5 windows = the 64-frame clip)
CONFIG3 = dict(steps=4, guidance=1.1, cond_scale=[1.0, 0.35, 1.0, 0.4])


def ncu_traffic(family: str):
    """DRAM bytes per launch of `family` from the newest committed ncu launch list (profiles/r*_traffic.json, written by
    scripts/launch_traffic.py from the same bench command under ncu); None when no capture is committed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return None
    try:
        with open(files[-1]) as fh:
            d = json.load(fh)
        return float(d["families"][family]["traffic_per_launch"])
    except (KeyError, ValueError, OSError):
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]), tf_sustained=float(d["bf16_tflops_sustained"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": reasons, "samples": len(rows)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (CPU restatement of the reference's PyTorch path)
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_factory(frames_sample: int, latent: int, n_nets: int = 2, seed: int = 0):
    """Return (step_fn, description, kind).  One denoising step of config 2 restricted to `frames_sample` (>= 4, so that the
    temporal attention is real) of the 16 frames at full resolution.  The UNet3D is the REFERENCE'S OWN module
    (animatediff/models/unet.py through oracle/diffusers_shim, shipped under baseline/_ref by build()) when it can be
    imported, else the oracle port; the ControlNets are diffusers code (third party, absent) -> always the oracle port."""
    from oracle import ref_import
    from oracle import ref_ops as R
    from oracle import ref_unet3d as U
    from oracle import synth
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = synth.unet_config(tiny=False)
    gen = torch.Generator().manual_seed(seed)

    def fill(shapes):
        sd = {}
        for k, shp in shapes.items():
            if k.endswith(".pe"):
                sd[k] = R.positional_encoding(shp[1], shp[2])
            elif len(shp) >= 2:
                fan_in = 1
                for s in shp[1:]:
                    fan_in *= s
                sd[k] = torch.randn(shp, generator=gen) * fan_in ** -0.5
            elif k.endswith("weight") and "norm" in k:
                sd[k] = 1 + 0.1 * torch.randn(shp, generator=gen)
            else:
                sd[k] = 0.1 * torch.randn(shp, generator=gen)
        return sd

    sd_u = fill(U.unet3d_shapes(cfg))
    sd_c = [fill(U.controlnet_shapes(cfg)) for _ in range(n_nets)]
    ref_unet, kind = None, "port"
    try:
        ref_import.import_reference()
        from animatediff.models.unet import UNet3DConditionModel          # the reference's own class
        from modules.attention_processor import AttnProcessor2_0
        with torch.device("meta"):
            ref_unet = UNet3DConditionModel(**cfg)
        ref_unet = ref_unet.to_empty(device="cpu")
        missing = ref_unet.load_state_dict(sd_u, strict=False)            # strays: BasicTransformerBlock's unused Attention params
        with torch.no_grad():
            for k in missing.missing_keys:
                ref_unet.state_dict()[k].zero_()
        ref_unet.set_attn_processor(AttnProcessor2_0())
        ref_unet.eval()
        kind = "reference"
    except Exception as e:  # noqa: BLE001 - the port is the documented fallback
        sys.stderr.write(f"[bench] reference modules unavailable ({type(e).__name__}: {e}); timing the oracle port\n")
        ref_unet = None
    f = frames_sample
    latents = torch.randn(1, 4, f, latent, latent, generator=gen)
    prompt = torch.randn(2, 77, 768, generator=gen)
    images = [torch.randn(2 * f, 3, latent * 8, latent * 8, generator=gen) for _ in range(n_nets)]
    acp = R.ddim_alphas_cumprod()
    t = R.ddim_timesteps(DDIM_STEPS)[0]
    state = {"latents": latents}

    @torch.no_grad()
    def step():
        lat = state["latents"]
        model_in = torch.cat([lat] * 2)
        x2d = model_in.permute(0, 2, 1, 3, 4).reshape(2 * f, 4, latent, latent)
        ctx = torch.cat([prompt] * f)
        per_net = [U.controlnet_forward(sd_c[k], cfg, x2d, t, ctx, images[k]) for k in range(n_nets)]
        down, mid = R.merge_controlnet_residuals(per_net, COND_SCALE[:n_nets], f)
        if ref_unet is not None:
            noise = ref_unet(model_in, t, encoder_hidden_states=prompt, down_block_additional_residuals=down,
                             mid_block_additional_residual=mid).sample
        else:
            noise = U.unet3d_forward(sd_u, cfg, model_in, t, prompt, down, mid)
        state["latents"] = R.ddim_step(R.cfg_combine(noise, GUIDANCE), t, lat, acp, DDIM_STEPS)

    what = ("the reference's own UNet3DConditionModel (animatediff/models/unet.py via oracle/diffusers_shim, AttnProcessor2_0) + oracle-port "
            "ControlNets (diffusers, third party)") if kind == "reference" else "oracle port (torch fp32 CPU restatement of the reference path)"
    desc = (f"{what}, fp32, one denoising step of config 2 on {f} of {FRAMES} frames at full 512x512 (latent {latent}x{latent}), b=2 CFG, "
            f"{n_nets} ControlNets; frames/s = {f}/(20*step_s)")
    return step, desc, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    f = args.cpu_frames
    step, desc, kind = cpu_reference_step_factory(f, args.latent)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = f / (DDIM_STEPS * dt)
    line = {
        "impl": "reference", "metric": "denoised frames/s", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, parallelism="cpu"),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, parallelism):
    return {"workload": f"SD1.5 UNet3D + mm_sd_v15_v2 motion modules + 2 ControlNets, {FRAMES} frames 512x512 (latent "
                        f"{args.latent}x{args.latent}), CFG {GUIDANCE} (b=2), DDIM {DDIM_STEPS} steps; one bench step = one denoising step",
            "frames_per_window": FRAMES, "ddim_steps": DDIM_STEPS, "controlnets": 2, "cond_scale": COND_SCALE,
            "parallelism": parallelism, "window_overlap_frames": OVERLAP,
            "l2": "per-step working set (>10 GB of activations and 4 GB of weights) exceeds the 126 MB L2; no explicit flush"}


def torch_eager_yardstick(dev, latent, frames=FRAMES, steps=2):
    """The reference's op sequence under stock PyTorch on the same GPU (SURVEY §2.1's B200 bar): the oracle restatement run
    in bf16 on CUDA with F.scaled_dot_product_attention (what the reference's AttnProcessor2_0 dispatches to), cuDNN
    convolutions, native GroupNorm / LayerNorm arithmetic, one launch per op.  A yardstick under `extra`, never the
    reference arm and never a product path."""
    from oracle import ref_ops as R
    from oracle import ref_unet3d as U
    from oracle import synth
    cfg = synth.unet_config(tiny=False)
    gen = torch.Generator(device=dev).manual_seed(7)

    def fill(shapes):
        sd = {}
        for k, shp in shapes.items():
            if k.endswith(".pe"):
                sd[k] = R.positional_encoding(shp[1], shp[2]).to(dev)
            elif len(shp) >= 2:
                fan_in = 1
                for v in shp[1:]:
                    fan_in *= v
                sd[k] = (torch.randn(shp, generator=gen, device=dev) * fan_in ** -0.5).bfloat16()
            elif k.endswith("weight") and "norm" in k:
                sd[k] = (1 + 0.1 * torch.randn(shp, generator=gen, device=dev)).bfloat16()
            else:
                sd[k] = (0.1 * torch.randn(shp, generator=gen, device=dev)).bfloat16()
        return sd

    sd_u, sd_c = fill(U.unet3d_shapes(cfg)), [fill(U.controlnet_shapes(cfg)) for _ in range(2)]
    f = frames
    lat = torch.randn(1, 4, f, latent, latent, generator=gen, device=dev).bfloat16()
    prompt = torch.randn(2, 77, 768, generator=gen, device=dev).bfloat16()
    images = [torch.randn(2 * f, 3, latent * 8, latent * 8, generator=gen, device=dev).bfloat16() for _ in range(2)]
    acp, t = R.ddim_alphas_cumprod().to(dev), R.ddim_timesteps(DDIM_STEPS)[0]

    @torch.no_grad()
    def step(x):
        model_in = torch.cat([x] * 2)
        x2d = model_in.permute(0, 2, 1, 3, 4).reshape(2 * f, 4, latent, latent)
        ctx = torch.cat([prompt] * f)
        per_net = [U.controlnet_forward(sd_c[k], cfg, x2d, t, ctx, images[k]) for k in range(2)]
        down, mid = R.merge_controlnet_residuals(per_net, COND_SCALE, f)
        noise = U.unet3d_forward(sd_u, cfg, model_in, t, prompt, down, mid)
        return R.ddim_step(R.cfg_combine(noise.float(), GUIDANCE), t, x.float(), acp, DDIM_STEPS).bfloat16()

    R.USE_SDPA = True
    try:
        lat = step(lat)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            lat = step(lat)
        e1.record()
        torch.cuda.synchronize()
    finally:
        R.USE_SDPA = False
    ms = e0.elapsed_time(e1) / steps
    del sd_u, sd_c
    torch.cuda.empty_cache()
    return {"ms_per_step": ms, "frames_per_s": f / (DDIM_STEPS * ms * 1e-3),
            "what": "oracle port of the reference path in bf16 on this GPU, stock PyTorch eager (cuBLAS, cuDNN, SDPA), same workload"}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from controlanimate_b200 import _lib, parallel, pipeline, profiler, unet as un, utils

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load(build_if_missing=False)
    if lib.ca_device_sm() != 100:
        raise SystemExit(f"sm_{lib.ca_device_sm()} device: these kernels are sm_100a only")
    torch.backends.cudnn.benchmark = True

    mode = args.parallelism if world > 1 else "single"
    c3 = args.workload in ("config3", "config3-cfg")           # config 3: 4 ControlNets, 4 steps
    lcm = args.workload == "config3"                           # ... in the LCM branch (b = 1); "config3-cfg": CFG 1.1, b = 2
    n_steps = CONFIG3["steps"] if c3 else DDIM_STEPS
    cond_scale = CONFIG3["cond_scale"] if c3 else COND_SCALE
    guidance = CONFIG3["guidance"] if c3 else GUIDANCE
    n_nets = len(cond_scale)
    cfg = utils.sd15_unet3d_config(time_cond_proj_dim=256 if lcm else None)
    dtype = torch.bfloat16
    unet = utils.build_on_device(lambda: un.UNet3DConditionModel(**cfg), dev, dtype, seed=1)
    nets = [utils.build_on_device(lambda: un.ControlNetModel(), dev, dtype, seed=2 + k) for k in range(n_nets)]
    mc = pipeline.MultiControlNetResiduals(nets, cond_scale)
    mc.overlap = bool(args.stream_overlap) and not args.no_graph
    mc.hoist_cond_embedding = bool(args.hoist_cond_embedding)
    sched = pipeline.DDIMScheduler()
    timesteps = sched.set_timesteps(n_steps)
    step_par = parallel.StepParallel(mode, rank, world, n_nets=n_nets) if mode in parallel.StepParallel.MODES else None
    clip = None
    if mode == "clip":
        n_win = args.windows or (5 if world == 8 else max(1, world // 2))
        clip = parallel.ClipLayout(rank, world, n_win, n_nets, args.local_nets or max(1, n_nets // 2))
    else:
        loop = pipeline.DenoisingLoop(unet, mc, sched, guidance_scale=guidance, use_cuda_graph=not args.no_graph, parallel=step_par,
                                      use_lcm=lcm)
    windows = parallel.WindowParallel(rank, world, FRAMES, OVERLAP) if mode == "windows" else None

    f, lat = FRAMES, args.latent
    # windows: every rank has its own window of the clip; step parallelism: every rank works on THE SAME window
    g = torch.Generator(device="cpu").manual_seed(100 + (rank if mode in ("windows", "clip") else 0))
    # host-side inputs (pinned): what a caller of the public API holds
    h_latents = torch.randn(1, 4, f, lat, lat, generator=g).pin_memory()
    rows = 1 if lcm else 2                                      # CFG duplicates the batch; the LCM branch does not
    h_prompt = torch.randn(rows, 77, 768, generator=g).to(dtype).pin_memory()
    h_images = [torch.randn(rows * f, 3, lat * 8, lat * 8, generator=g).to(dtype) for _ in range(n_nets)]
    mc.prep_images = [im.to(dev) for im in h_images]          # prepared once per window (prep_control_images)
    d_prompt = h_prompt.to(dev)
    h_out = torch.empty_like(h_latents).pin_memory()
    if clip is not None:
        # the control video of job (window, net) lives on the rank that evaluates it
        images = {}
        for (w_, k_) in clip.jobs[rank]:
            gi = torch.Generator(device="cpu").manual_seed(1000 + 10 * w_ + k_)
            images[(w_, k_)] = torch.randn(rows * f, 3, lat * 8, lat * 8, generator=gi).to(dtype).to(dev)
        mc.prep_images = None
        loop = pipeline.ClipLoop(unet, mc, sched, clip, images, (1, 4, f, lat, lat), guidance_scale=guidance, use_lcm=lcm,
                                 use_cuda_graph=not args.no_graph, overlap=OVERLAP)

    def one_step(latents, i):
        t = timesteps[i % n_steps]
        if clip is not None:
            return loop.step(latents if clip.is_unet_rank else None, t, d_prompt)
        latents = loop.step(latents, t, d_prompt)
        if windows is not None:
            latents = windows.exchange(latents)
        return latents

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident run (value) -------------------------------------------------------------------------
    latents = h_latents.to(dev)
    for i in range(args.warmup):
        latents = one_step(latents, i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    profiler.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ncu_range = os.environ.get("CA_NCU_RANGE") == "1"      # `ncu --profile-from-start off`: capture only the timed steps
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    mc._cond_cache.clear()       # (--hoist-cond-embedding) the once-per-window embedding is evaluated INSIDE the timed region
    for i in range(args.steps):
        latents = one_step(latents, i)
    e1.record()
    barrier()
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStop()
    ms = e0.elapsed_time(e1) / args.steps
    launches = profiler.launch_count()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end through the public API with HOST buffers (e2e) ----------------------------------------------
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    mc._cond_cache.clear()
    for i in range(args.steps):
        if clip is not None and not clip.is_unet_rank:       # a ControlNet server: its inputs arrive over NVLink
            loop.step(None, timesteps[i % n_steps], d_prompt)
            continue
        d_lat = h_latents.to(dev, non_blocking=True)
        d_p = h_prompt.to(dev, non_blocking=True)
        out = loop.step(d_lat, timesteps[i % n_steps], d_p)
        if windows is not None:
            out = windows.exchange(out)
        h_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller consumes the latents on the host
        h_latents.copy_(h_out)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3) / args.steps

    # max over ranks
    if world > 1:
        tms = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tms[0]), float(tms[1])

    # ---- per-kernel timing pass (CUDA events around every C-ABI launch, on the launching stream) ---------------
    # every rank runs it (the step contains collectives when N > 1); rank 0 reports its own kernels
    loop.use_cuda_graph = False                # eager: every launch bracketed by events, GPU kept backlogged
    lat2 = one_step(latents, 0)
    torch.cuda.synchronize()
    profiler.enable(True, backlog_ms=400.0)
    lat2 = one_step(lat2, 1)
    torch.cuda.synchronize()
    kern = profiler.summary(measured_peaks())
    profiler.enable(False)
    # shares are of the TIMED (graph-replayed) step, not of the instrumented pass whose wall clock holds the backlog
    for fam in kern["families"].values():
        fam["share_of_step"] = fam["ms_total"] / ms
    kern["dominant"]["share_of_step"] = kern["families"][kern["dominant"]["kernel"]]["share_of_step"]
    kern["own_share"] = sum(fam["ms_total"] for fam in kern["families"].values()) / ms
    if rank == 0 and args.torch_profile and world == 1:
        # attribution aid (never a bench value): which aten ops the non-library, non-own kernels of one eager step come from
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
            lat2 = loop.step(lat2, timesteps[2], d_prompt)
            torch.cuda.synchronize()
        with open(args.torch_profile, "w") as fh:
            fh.write(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=120,
                                                                         max_name_column_width=60, max_shapes_column_width=90))
    barrier()

    if rank == 0:
        peaks = measured_peaks()
        if mode == "clip":
            unique = clip.n_windows * f - (clip.n_windows - 1) * OVERLAP
            counted = clip.n_windows * f
            scaling = "strong"
            par_desc = (f"clip layout: {clip.n_windows} UNet ranks (one {f}-frame window each, {clip.local_nets} ControlNet(s) local) + "
                        f"{clip.n_servers} ControlNet server rank(s) for the other {sum(len(clip.jobs[r]) for r in range(clip.n_windows, world))} "
                        f"(window, net) jobs; raw residuals read over NVLink by kernel (3); {unique} unique frames")
        elif mode == "windows" or world == 1:
            unique = world * f - (world - 1) * OVERLAP          # adjacent windows share OVERLAP frames
            counted = world * f
            scaling = "weak"
            par_desc = "single" if world == 1 else f"frame-windows x{world} (overlap {OVERLAP}: {unique} unique frames of {counted} denoised)"
        else:
            unique = counted = f                                # one window, split over the ranks
            scaling = "strong"
            par_desc = {"cfg": "CFG halves x2 (one window)", "controlnet": f"UNet rank + {world - 1} ControlNet rank(s) (one window)",
                        "cfg+controlnet": f"CFG halves x (UNet rank + {world // 2 - 1} ControlNet rank(s)) (one window)"}[mode]
        value = unique / (n_steps * ms * 1e-3)
        e2e = unique / (n_steps * ms_e2e * 1e-3)
        top = kern["dominant"]
        top["traffic"] = ncu_traffic(top["kernel"])
        top["traffic_source"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu launch list of the same step (profiles/)"
        line = {
            "metric": "denoised frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, parallelism=par_desc) if not c3 else {
                "workload": f"config 3: SD1.5 UNet3D{' (time_cond_proj_dim 256)' if lcm else ''} + motion modules + 4 ControlNets {cond_scale}, "
                            + (f"LCM branch (b=1, guidance {guidance} through timestep_cond), " if lcm else f"CFG {guidance} (b=2), ")
                            + f"{FRAMES}-frame windows 512x512 (latent {args.latent}x{args.latent}), "
                            f"{n_steps} steps; one bench step = one denoising step; scheduler arithmetic = DDIM update (the LCM scheduler is "
                            "outside the hot path)", "frames_per_window": FRAMES, "steps": n_steps, "controlnets": n_nets,
                "cond_scale": cond_scale, "parallelism": par_desc, "window_overlap_frames": OVERLAP},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h_latents.numel() * 4 + h_prompt.numel() * 2, "d2h_bytes_per_step": h_out.numel() * 4},
            "frames_unique": unique, "frames_counted": counted,
            "frames_counted_per_s": counted / (n_steps * ms * 1e-3),
            # kernels of libca_b200.so executed inside the timed region (counted per launch in eager mode; with CUDA-graph
            # replay = launches recorded in one captured step x timed steps)
            "gpu_launches": launches if args.no_graph else kern["launches_per_step"] * args.steps,
            "cuda_graph": not args.no_graph,
            "controlnet_streams": n_nets if mc.overlap else 0,
            "cond_embedding": ("hoisted: once per window, evaluated once inside each timed region" if mc.hoist_cond_embedding
                               else "per step (as the reference)"),
            "roofline": top,
            "kernels": kern["families"],
            "own_kernel_share_of_step": kern["own_share"],
            "peaks": peaks,
        }
        if step_par is not None and step_par.g > 1:
            # bytes kernel (3) pulls over NVLink per step on each UNet rank: the raw residual sets of the sharded ControlNets
            rows_cn = (1 if (step_par.halves == 2 or lcm) else 2) * f
            per_net = sum(c * (lat // d) * (lat // d) for c, d in ((320, 1),) * 3 + ((320, 2),) + ((640, 2),) * 2 + ((640, 4),)
                          + ((1280, 4),) * 2 + ((1280, 8),) * 4) * rows_cn * 2
            line["nvlink_bytes_per_step_per_unet_rank"] = per_net * n_nets
        extra = {}
        if args.eager_yardstick and world == 1 and not c3:
            del loop, unet, nets, mc
            torch.cuda.empty_cache()
            try:
                extra["torch_eager_gpu"] = torch_eager_yardstick(dev, args.latent)
            except Exception as e:  # noqa: BLE001 - a yardstick must never fail the bench
                extra["torch_eager_gpu"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        if extra:
            line["extra"] = extra
        if args.cpu_baseline and world == 1 and not c3:
            step, desc, kind = cpu_reference_step_factory(args.cpu_frames, args.latent)
            t0 = time.perf_counter()
            step()
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": args.cpu_frames / (DDIM_STEPS * dt), "unit": "frames/s", "cores": torch.get_num_threads(),
                                    "kind": kind, "sample": desc, "ms_per_step": dt * 1e3}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--latent", type=int, default=LATENT)
    ap.add_argument("--cpu-frames", type=int, default=4, help="frames of the bounded CPU sample (>= 4: temporal attention is real)")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config3-cfg"],
                    help="config2 = the headline (BASELINE configs[1]); config3 = LCM + 4 ControlNets, 4 steps (configs[2])")
    ap.add_argument("--windows", type=int, default=0, help="clip layout: UNet ranks = windows of the clip (default: 5 of 8 ranks, "
                    "else half the ranks); the other ranks serve ControlNet jobs")
    ap.add_argument("--local-nets", type=int, default=0, help="clip layout: ControlNets a UNet rank evaluates itself "
                    "(0 = half of them: measured best for config 3 on 8 GPUs, profiles/r02d_*)")
    ap.add_argument("--parallelism", default="windows", choices=["windows", "cfg", "controlnet", "cfg+controlnet", "clip"],
                    help="how N > 1 GPUs are used (see the module docstring)")
    ap.add_argument("--no-eager-yardstick", dest="eager_yardstick", action="store_false",
                    help="skip the torch bf16 eager yardstick (the reference's op sequence under stock PyTorch on the same GPU)")
    ap.add_argument("--hoist-cond-embedding", type=int, default=0,
                    help="1: evaluate every ControlNet's conditioning embedding (a function of the control video only) once per "
                         "window instead of once per step; it is recomputed once inside each timed region.  Default 0: a step "
                         "does exactly the work of a reference step")
    ap.add_argument("--stream-overlap", type=int, default=int(os.environ.get("CA_STREAM_OVERLAP", "0")),
                    help="1: every ControlNet on its own CUDA stream next to the UNet encoder (inside the captured graph)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--torch-profile", default="", help="write a torch.profiler op table of one eager step to this path")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
