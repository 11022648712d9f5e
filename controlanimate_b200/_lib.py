"""ctypes binding of libca_b200.so — the C ABI declared in include/controlanimate_b200.h.

There is NO fallback: if the library cannot be loaded (or a call fails) a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CA_LIB_PATH") or os.path.join(HERE, "libca_b200.so")   # CA_LIB_PATH: development builds only

CA_BF16, CA_F16, CA_F32 = 0, 1, 2
CA_LAYOUT_NCFHW, CA_LAYOUT_BFHWC = 0, 1
CA_EPI_NONE, CA_EPI_GEGLU = 0, 1
CA_MAX_NETS, CA_MAX_RESIDUALS = 8, 16

# name -> (restype, argtypes); mirrors include/controlanimate_b200.h one to one
_vp, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
SIGNATURES = {
    "ca_version": (C.c_char_p, []),
    "ca_last_error": (C.c_char_p, []),
    "ca_device_sm": (_i, []),
    "ca_groupnorm_workspace_bytes": (_sz, [_i] * 9),
    "ca_groupnorm_silu": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _i, _f, _i, _i, _i, _i, _vp, _sz, _vp]),
    "ca_residual_merge": (_i, [C.POINTER(_vp), C.POINTER(_f), C.POINTER(_vp), C.POINTER(_i), _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ca_layernorm_pe": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _f, _i, _vp]),
    "ca_temporal_attn_core": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _i, _f, _i, _vp]),
    "ca_temporal_attn_fused": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _i, _vp]),
    "ca_cross_attn_core": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _ll, _ll, _vp, _f, _i, _vp]),
    "ca_bias_act_residual": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _f, _i, _i, _vp]),
    "ca_linear": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _ll, _ll, _i, _i, _vp]),
    "ca_spatial_attn_core": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _f, _i, _vp]),
    "ca_upsample_nearest": (_i, [_vp, _vp, _ll, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ca_concat_channels": (_i, [_vp, _vp, _vp, _ll, _i, _i, _i, _vp]),
    "ca_row_stats": (_i, [_vp, _vp, _ll, _i, _ll, _f, _i, _vp]),
    "ca_cfg_ddim_step": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _f, _f, _f, _f, _f, _i, _i, _vp]),
    "ca_linear_ln": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _ll, _i, _i, _ll, _ll, _i, _i, _vp]),
}

_lock = threading.Lock()
_lib = None


def load(build_if_missing: bool = True):
    """Load (building first if the .so is absent and nvcc exists). Raises RuntimeError on failure."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH) and build_if_missing:
            from . import build as _build
            _build.build()
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m controlanimate_b200.build` (no CPU fallback exists)")
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise RuntimeError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise RuntimeError(f"{LIB_PATH} does not export {name}") from e
            fn.restype, fn.argtypes = res, args
        _lib = lib
        return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().ca_last_error().decode(errors="replace")
        if status in (1, 2):
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what}: {msg} (status {status})")
