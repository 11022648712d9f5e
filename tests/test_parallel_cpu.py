"""world_size-2/3 gloo tests (CPU) of the multi-GPU host logic in controlanimate_b200/parallel.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from controlanimate_b200 import parallel as P


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, world):
    port = _free_port()
    mp.spawn(_entry, args=(fn, world, port), nprocs=world, join=True)


def _entry(rank, fn, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world)
    finally:
        dist.destroy_process_group()


def _full_clip(n_frames):
    g = torch.Generator().manual_seed(5)
    return torch.randn(1, 4, n_frames, 6, 5, generator=g)


def _window_worker(rank, world):
    frames, ov = 8, 2
    starts = P.window_starts(world, frames, ov)
    wins = []
    for w, s in enumerate(starts):       # every window perturbs its copy so the overlap frames really differ
        wins.append(_full_clip(starts[-1] + frames)[:, :, s:s + frames] + 0.1 * w)
    want = P.blend_windows_reference(wins, ov)
    got = P.WindowParallel(rank, world, frames, ov).exchange(wins[rank])
    assert torch.allclose(got, want[rank], atol=1e-6)
    # interior frames are untouched
    assert torch.equal(got[:, :, ov:frames - ov], wins[rank][:, :, ov:frames - ov])


@pytest.mark.parametrize("world", [2, 3])
def test_window_parallel_exchange(world):
    _run(_window_worker, world)


def _cfg_worker(rank, world):
    g = torch.Generator().manual_seed(9)
    noise = torch.randn(2, 4, 3, 4, 4, generator=g)
    cfg = P.CFGParallel(rank, world)
    out = cfg.combine(cfg.my_rows(noise), 7.5)
    u, c = noise.chunk(2)
    assert torch.allclose(out, u + 7.5 * (c - u), atol=1e-6)


def test_cfg_parallel():
    _run(_cfg_worker, 2)


def _cn_worker(rank, world):
    n_nets = 3
    g = torch.Generator().manual_seed(11)
    sets = [[torch.randn(4, 8, 3, 3, generator=g), torch.randn(4, 16, 2, 2, generator=g)] for _ in range(n_nets)]
    cp = P.ControlNetParallel(rank, world, n_nets)
    assert sorted(sum(([k for k in range(n_nets) if k % world == r] for r in range(world)), [])) == [0, 1, 2]
    got = cp.gather([sets[k] for k in cp.my_nets()])
    for k in range(n_nets):
        for a, b in zip(got[k], sets[k]):
            assert torch.equal(a, b)


def test_controlnet_parallel_gather():
    _run(_cn_worker, 2)


def test_window_starts_match_config3():
    assert P.window_starts(5, 16, 4) == [0, 12, 24, 36, 48]


def test_step_parallel_role_tables():
    """Rank layouts of parallel.StepParallel (no process group needed)."""
    sp = [P.StepParallel("cfg+controlnet", r, 4, n_nets=2) for r in range(4)]
    assert [s.half for s in sp] == [0, 0, 1, 1] and [s.role for s in sp] == [0, 1, 0, 1]
    assert sp[0].my_nets() == [] and sp[1].my_nets() == [0, 1] and sp[3].my_nets() == [0, 1]
    assert sp[2].unet_rank == 2 and sp[3].unet_rank == 2 and sp[0].unet_ranks() == [0, 2]
    assert [sp[2].owner_of(k) for k in range(2)] == [3, 3]
    sp6 = [P.StepParallel("cfg+controlnet", r, 6, n_nets=2) for r in range(6)]
    assert [s.my_nets() for s in sp6] == [[], [0], [1], [], [0], [1]]
    assert [sp6[3].owner_of(k) for k in range(2)] == [4, 5]
    cfg = [P.StepParallel("cfg", r, 2, n_nets=2) for r in range(2)]
    assert cfg[0].my_nets() == [0, 1] and cfg[1].is_unet_rank and cfg[1].half == 1
    cn = [P.StepParallel("controlnet", r, 3, n_nets=4) for r in range(3)]
    assert [s.my_nets() for s in cn] == [[], [0, 2], [1, 3]] and all(s.half == 0 for s in cn)
    with pytest.raises(ValueError):
        P.StepParallel("cfg", 0, 3, n_nets=2)
    with pytest.raises(ValueError):
        P.StepParallel("controlnet", 0, 4, n_nets=2)       # 3 ControlNet ranks for 2 nets
    with pytest.raises(ValueError):
        P.ControlNetParallel(0, 3, 2)


def _combine_worker(rank, world):
    g = torch.Generator().manual_seed(13)
    noise = torch.randn(2, 4, 3, 4, 4, generator=g)
    u, c = noise.chunk(2)
    want = u + 7.5 * (c - u)
    like = torch.zeros(1, 4, 3, 4, 4)
    if world == 4:      # CFG halves x (UNet rank + ControlNet rank)
        sp = P.StepParallel("cfg+controlnet", rank, world, n_nets=2)
        local = sp.rows(noise) if sp.is_unet_rank else None
    else:               # UNet rank (both rows) + one ControlNet rank
        sp = P.StepParallel("controlnet", rank, world, n_nets=2)
        local = noise if sp.is_unet_rank else None
    out = sp.combine_noise(local, like, 7.5)
    assert torch.allclose(out, want, atol=1e-6)


@pytest.mark.parametrize("world", [2, 4])
def test_step_parallel_combine_noise(world):
    _run(_combine_worker, world)


def test_clip_layout_is_the_survey_plan_for_config_3():
    """8 GPUs, 5 windows x 4 ControlNets -> ranks 0-4 = UNet_w + one ControlNet of window w, ranks 5-7 = the other
    15 jobs, 5 each; every job has exactly one owner and a UNet rank's remote sets come from at most two servers."""
    from controlanimate_b200.parallel import ClipLayout
    lay = [ClipLayout(r, 8, 5, 4) for r in range(8)]
    jobs = [j for r in range(8) for j in lay[0].jobs[r]]
    assert sorted(jobs) == [(w, k) for w in range(5) for k in range(4)]
    assert all(lay[0].jobs[w] == [(w, 0)] for w in range(5)) and all(len(lay[0].jobs[s]) == 5 for s in (5, 6, 7))
    assert lay[0].max_slots() == 5 and [lay[w].is_unet_rank for w in range(8)] == [True] * 5 + [False] * 3
    for w in range(5):
        assert 1 <= len(lay[0].servers_of(w)) <= 2
        for k in range(1, 4):
            o = lay[0].owner_of(w, k)
            assert lay[0].jobs[o][lay[0].slot_of(w, k)] == (w, k) and w in lay[0].windows_of(o)
    # no server ranks: plain windows, every net local;  more local nets: fewer served jobs
    assert ClipLayout(0, 4, 4, 4).jobs[2] == [(2, k) for k in range(4)] and ClipLayout(0, 4, 4, 4).servers_of(1) == []
    two = ClipLayout(0, 8, 5, 4, local_nets=2)
    assert sorted(len(two.jobs[s]) for s in (5, 6, 7)) == [3, 3, 4]
    with pytest.raises(ValueError):
        ClipLayout(0, 4, 5, 4)
    with pytest.raises(ValueError):
        ClipLayout(0, 8, 1, 4, local_nets=2)       # 7 servers for 2 served jobs
