import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A bare `pytest` on a machine without a GPU skips the gpu-marked tests instead of failing them (an explicit
    `-m gpu` still runs — and fails loudly — there: a GPU box with a broken driver must not look green)."""
    import torch
    if torch.cuda.is_available() or "gpu" in (config.getoption("-m") or ""):
        return
    skip = pytest.mark.skip(reason="no CUDA device: gpu-marked tests need a B200")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))

    return load


def pytest_sessionfinish(session, exitstatus):
    """Kernel parity tests record, next to the asserted error, the plain (no half-ulp subtraction) numbers of every comparison:
    dump them where a gpurun call brings them back (gpurun_out/kernel_relerr.tsv), worst first."""
    mod = sys.modules.get("test_gpu_kernels") or sys.modules.get("tests.test_gpu_kernels")
    rows = getattr(mod, "PLAIN", None) if mod is not None else None
    if not rows:
        return
    out = os.path.join(ROOT, "gpurun_out")
    try:                                 # a report, never a reason for the session to fail (read-only checkout, full disk)
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "kernel_relerr.tsv"), "w") as fh:
            fh.write("test\tdtype\tasserted (max|err| - half ulp) / max|ref|\tplain max|err| / max|ref|\tworst per-element |err| / max(|ref|, max|ref|/64)\n")
            for r in sorted(rows, key=lambda r: -r[3]):
                fh.write(f"{r[0]}\t{r[1]}\t{(r[2] if r[2] is not None else float('nan')):.3e}\t{r[3]:.3e}\t{r[4]:.3e}\n")
    except OSError:
        pass
