"""BASELINE config 4 at full width (f = 32, 96x54 latents, 81 prompt tokens) on one B200: scripts/config4_check.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_config4_full_width_forward_and_alternative_paths():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "config4_check.py")], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "config 4 cross-checks ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
