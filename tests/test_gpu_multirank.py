"""-m gpu tests that need SEVERAL GPUs (NCCL over NVLink, symmetric memory): window exchange, CFG split, ControlNet sharding.
Each case launches tests/multirank_check.py under torchrun; cases needing more GPUs than the box has are skipped."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode,world", [("windows", 2), ("cfg", 2), ("controlnet", 2), ("controlnet", 3), ("cfg+controlnet", 4), ("clip", 3), ("clip", 4)])
def test_multirank(mode, world):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world + len(mode)), os.path.join(ROOT, "tests", "multirank_check.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and f"multirank {mode} ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
