"""Locate and import the REFERENCE'S OWN PYTHON (TEST INFRASTRUCTURE).

The reference needs diffusers 0.23 / controlnet_aux, which are not installed: `oracle/diffusers_shim` provides the symbols
it imports.  In the build container the sources lie under /root/reference; `__graft_entry__.build()` copies the few
modules of the hot path into the git-ignored `baseline/_ref/` so that they travel to the GPU box with the snapshot
(nothing of the reference is ever committed).  Only tests/, bench.py's CPU / yardstick legs and oracle/gen_golden.py
import this.
"""
from __future__ import annotations

import os
import shutil
import sys
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SHIPPED = os.path.join(ROOT, "baseline", "_ref")
NEEDED = ("animatediff/models", "modules")


def reference_root() -> Optional[str]:
    for cand in (os.environ.get("CA_REFERENCE_ROOT"), "/root/reference", SHIPPED):
        if cand and os.path.isfile(os.path.join(cand, "animatediff", "models", "unet.py")):
            return cand
    return None


def ship_reference(src: str = "/root/reference") -> Optional[str]:
    """Copy the hot-path modules of the reference into baseline/_ref (git-ignored).  Returns the destination or None."""
    if not os.path.isdir(src):
        return None
    for sub in NEEDED:
        d = os.path.join(SHIPPED, sub)
        os.makedirs(d, exist_ok=True)
        for f in os.listdir(os.path.join(src, sub)):
            if f.endswith(".py"):
                shutil.copy2(os.path.join(src, sub, f), os.path.join(d, f))
    return SHIPPED


def import_reference() -> str:
    """Put the shim and the reference on sys.path; returns the root used.  Raises RuntimeError if no copy exists."""
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference sources are neither at /root/reference nor shipped under baseline/_ref "
                           "(run __graft_entry__.build() in the build container)")
    for pth in (root, os.path.join(HERE, "diffusers_shim")):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return root
