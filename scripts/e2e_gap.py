"""python scripts/e2e_gap.py: where the e2e step (host buffers, one synchronisation per step) loses time against the
device-resident step: CPU time before / inside / after the CUDA-graph launch of one DenoisingLoop.step (development aid)."""
import os, sys, time, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from controlanimate_b200 import pipeline, unet as un, utils
dev, dtype = torch.device("cuda"), torch.bfloat16
torch.backends.cudnn.benchmark = True
cfg = utils.sd15_unet3d_config()
unet = utils.build_on_device(lambda: un.UNet3DConditionModel(**cfg), dev, dtype, seed=1)
nets = [utils.build_on_device(lambda: un.ControlNetModel(), dev, dtype, seed=2 + k) for k in range(2)]
mc = pipeline.MultiControlNetResiduals(nets, [1.0, 0.5])
sched = pipeline.DDIMScheduler()
ts = sched.set_timesteps(20)
loop = pipeline.DenoisingLoop(unet, mc, sched, guidance_scale=7.5, use_cuda_graph=True)
g = torch.Generator().manual_seed(0)
h_lat = torch.randn(1, 4, 16, 64, 64, generator=g).pin_memory()
h_p = torch.randn(2, 77, 768, generator=g).to(dtype).pin_memory()
mc.prep_images = [torch.randn(32, 3, 512, 512, generator=g).to(dtype).to(dev) for _ in range(2)]
h_out = torch.empty_like(h_lat).pin_memory()
acc = {"replay_cpu": 0.0, "replays": 0}
orig = torch.cuda.CUDAGraph.replay
def timed(self):
    t = time.perf_counter(); orig(self); acc["replay_cpu"] += time.perf_counter() - t; acc["replays"] += 1
torch.cuda.CUDAGraph.replay = timed
lat = h_lat.to(dev)
for i in range(3):
    lat = loop.step(lat, ts[i], h_p.to(dev))
torch.cuda.synchronize()
# device-resident
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
d_p = h_p.to(dev)
e0.record()
for i in range(8):
    lat = loop.step(lat, ts[i], d_p)
e1.record(); torch.cuda.synchronize()
resident = e0.elapsed_time(e1) / 8
acc.update(replay_cpu=0.0, replays=0)
seg = {"pre": 0.0, "step_cpu": 0.0, "wait": 0.0}
t_all = time.perf_counter()
for i in range(8):
    t0 = time.perf_counter()
    d_lat = h_lat.to(dev, non_blocking=True); d_p = h_p.to(dev, non_blocking=True)
    t1 = time.perf_counter()
    out = loop.step(d_lat, ts[i], d_p)
    h_out.copy_(out, non_blocking=True)
    t2 = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    h_lat.copy_(h_out)
    t3 = time.perf_counter()
    seg["pre"] += t1 - t0; seg["step_cpu"] += t2 - t1; seg["wait"] += t3 - t2
total = (time.perf_counter() - t_all) / 8 * 1e3
print(json.dumps(dict(resident_ms=resident, e2e_ms=total, cpu_pre_ms=seg["pre"] / 8 * 1e3, cpu_step_ms=seg["step_cpu"] / 8 * 1e3,
                      of_which_graph_launch_ms=acc["replay_cpu"] / 8 * 1e3, replays_per_step=acc["replays"] / 8,
                      wait_ms=seg["wait"] / 8 * 1e3)))
