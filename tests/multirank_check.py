"""Multi-rank GPU check (run by tests/test_gpu_multirank.py under torchrun, one rank per GPU, NCCL):

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multirank_check.py <mode>

mode "windows": WindowParallel.exchange over NCCL send/recv against the single-process blend of all windows.
mode "cfg" / "controlnet" / "cfg+controlnet": one denoising step of a small-width UNet3D + 2 ControlNets split over the
ranks (parallel.StepParallel; the ControlNet ranks' residuals are read over NVLink through symmetric memory) against the
SAME step computed by every rank alone with the same seeds — eager and CUDA-graph replay.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from controlanimate_b200 import _lib, parallel, pipeline, unet as un, utils  # noqa: E402


def cosine(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()))


def main():
    mode = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    _lib.load(build_if_missing=False)
    try:
        if mode == "windows":
            frames, ov = 16, 4
            starts = parallel.window_starts(world, frames, ov)
            g = torch.Generator().manual_seed(5)
            clip = torch.randn(1, 4, starts[-1] + frames, 8, 8, generator=g)
            wins = [clip[:, :, s:s + frames] + 0.1 * w for w, s in enumerate(starts)]
            want = parallel.blend_windows_reference(wins, ov)[rank].to(dev)
            got = parallel.WindowParallel(rank, world, frames, ov).exchange(wins[rank].to(dev))
            torch.cuda.synchronize()
            assert torch.allclose(got, want, atol=1e-6), float((got - want).abs().max())
        elif mode == "clip":
            # parallel.ClipLayout: world // 2 windows (UNet ranks) + the rest ControlNet servers, 4 nets, LCM branch (b = 1)
            # and CFG (b = 2); every UNet rank must end each step with what the windows computed alone and blended give
            n_win, n_nets, f, hh, ov = max(1, world // 2), 4, 8, 16, 2
            cfg = utils.sd15_unet3d_config(time_cond_proj_dim=256)
            cfg.update(block_out_channels=(64, 128, 256, 256), cross_attention_dim=64)
            dtype = torch.bfloat16
            unet = utils.build_on_device(lambda: un.UNet3DConditionModel(**cfg), dev, dtype, seed=1)
            nets = [utils.build_on_device(lambda: un.ControlNetModel(block_out_channels=cfg["block_out_channels"], cross_attention_dim=64),
                                          dev, dtype, seed=2 + k) for k in range(n_nets)]
            scales = [1.0, 0.35, 1.0, 0.4]
            sched = pipeline.DDIMScheduler()
            ts = sched.set_timesteps(4)
            for lcm in (True, False):
                rows = 1 if lcm else 2
                g = torch.Generator().manual_seed(3)                    # identical inputs on every rank
                lats = [torch.randn(1, 4, f, hh, hh, generator=g).to(dev) for _ in range(n_win)]
                prompt = torch.randn(rows, 7, 64, generator=g).to(dev, dtype)
                images = {(w, k): torch.randn(rows * f, 3, hh * 8, hh * 8, generator=g).to(dev, dtype)
                          for w in range(n_win) for k in range(n_nets)}
                want = [x.clone() for x in lats]
                for t in ts[:3]:                                        # every window alone, then the overlap blend
                    nxt = []
                    for w in range(n_win):
                        mc = pipeline.MultiControlNetResiduals(nets, scales)
                        mc.prep_images = [images[(w, k)] for k in range(n_nets)]
                        alone = pipeline.DenoisingLoop(unet, mc, sched, guidance_scale=7.5, use_lcm=lcm)
                        nxt.append(alone.step(want[w], t, prompt))
                    want = parallel.blend_windows_reference(nxt, ov)
                for graph in (False, True):
                    layout = parallel.ClipLayout(rank, world, n_win, n_nets, local_nets=1)
                    mc = pipeline.MultiControlNetResiduals(nets, scales)
                    loop = pipeline.ClipLoop(unet, mc, sched, layout, {j: images[j] for j in layout.jobs[rank]}, (1, 4, f, hh, hh),
                                             guidance_scale=7.5, use_lcm=lcm, use_cuda_graph=graph, overlap=ov)
                    cur = lats[rank].clone() if layout.is_unet_rank else None
                    for t in ts[:3]:
                        cur = loop.step(cur, t, prompt)
                    torch.cuda.synchronize()
                    if layout.is_unet_rank:
                        c = cosine(cur, want[rank])
                        assert c >= 0.9999, (mode, lcm, graph, rank, c)
                    dist.barrier()
        else:
            cfg = utils.sd15_unet3d_config()
            cfg.update(block_out_channels=(64, 128, 256, 256), cross_attention_dim=64)
            dtype = torch.bfloat16
            unet = utils.build_on_device(lambda: un.UNet3DConditionModel(**cfg), dev, dtype, seed=1)
            nets = [utils.build_on_device(lambda: un.ControlNetModel(block_out_channels=cfg["block_out_channels"], cross_attention_dim=64),
                                          dev, dtype, seed=2 + k) for k in range(2)]
            f, hh = 8, 16
            g = torch.Generator().manual_seed(3)                        # identical inputs on every rank
            latents = torch.randn(1, 4, f, hh, hh, generator=g).to(dev)
            prompt = torch.randn(2, 7, 64, generator=g).to(dev, dtype)
            images = [torch.randn(2 * f, 3, hh * 8, hh * 8, generator=g).to(dev, dtype) for _ in range(2)]
            sched = pipeline.DDIMScheduler()
            ts = sched.set_timesteps(4)

            def make(par, graph):
                mc = pipeline.MultiControlNetResiduals(nets, [1.0, 0.5])
                mc.prep_images = images
                return pipeline.DenoisingLoop(unet, mc, sched, guidance_scale=7.5, use_cuda_graph=graph, parallel=par)

            alone = make(None, False)
            want = [alone.step(latents, t, prompt) for t in ts[:3]]
            sp = parallel.StepParallel(mode, rank, world, n_nets=2)
            for graph in (False, True):
                loop = make(sp, graph)
                got = [loop.step(latents, t, prompt).clone() for t in ts[:3]]
                torch.cuda.synchronize()
                for a, b in zip(got, want):
                    c = cosine(a, b)
                    assert c >= 0.9999, (mode, graph, rank, c)
                # every rank ends the step with the same latents (the scheduler step is replicated)
                ref = got[-1].clone()
                dist.broadcast(ref, 0)
                assert torch.equal(ref, got[-1]), (mode, graph, rank)
        dist.barrier()
        if rank == 0:
            print(f"multirank {mode} ok on {world} GPUs")
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
