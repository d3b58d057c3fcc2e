"""ctypes binding of libmvster_b200.so (the C ABI declared in include/mvster_b200.h).

The library is the product: there is no Python/torch fallback.  If it has not been built, or no CUDA device is
present when a compute entry point is called, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_uint64, c_uint8, c_void_p

from ._build import LIB_PATH, build_library

_lock = threading.Lock()
_lib = None

_P = c_void_p

_SIGNATURES = {
    "mvster_version": (c_int, []),
    "mvster_last_error": (c_char_p, []),
    "mvster_launch_count": (c_uint64, []),
    "mvster_compose_homographies": (c_int, [_P, _P, c_int, c_int, _P]),
    "mvster_compose_homography_pair": (c_int, [_P, _P, _P, c_int, _P]),
    "mvster_epi_fwd": (c_int, [_P, POINTER(_P), _P, _P, _P, _P, _P] + [c_int] * 9 + [c_float, c_int, _P]),
    "mvster_epi_fwd_mode": (c_int, [_P, POINTER(_P), _P, _P, _P] + [c_int] * 9 + [c_float, c_int, c_int, c_int, _P]),
    "mvster_epi_fwd_mode_ex": (c_int, [_P, POINTER(_P), _P, _P, _P, _P] + [c_int] * 9 +
                               [c_float, c_int, c_int, c_int, _P]),
    "mvster_epi_bwd_mode": (c_int, [_P, POINTER(_P), _P, _P, _P, _P, _P, _P, POINTER(_P)] + [c_int] * 9 +
                            [c_float, c_int, c_int, c_int, _P]),
    "mvster_epi_bwd": (c_int, [_P, POINTER(_P), _P, _P, _P, _P, _P, _P, POINTER(_P)] + [c_int] * 9 +
                       [c_float, c_int, _P]),
    "mvster_homo_warp": (c_int, [_P, _P, _P, _P] + [c_int] * 7 + [c_int, _P]),
    "mvster_init_inverse_range": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _P]),
    "mvster_schedule_inverse_range": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "mvster_tail": (c_int, [_P, _P, c_float, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "mvster_regtail": (c_int, [_P, _P, _P, _P, _P, c_float, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "mvster_conv3d_small": (c_int, [_P, _P, _P, _P, _P] + [c_int] * 9 + [_P]),
    "mvster_conv3d_small_slice": (c_int, [_P, _P, _P, _P, _P] + [c_int] * 10 + [_P]),
    "mvster_conv2d_mid5": (c_int, [_P, _P, _P, _P] + [c_int] * 6 + [_P]),
    "mvster_conv3d_mid": (c_int, [_P, _P, _P, _P] + [c_int] * 8 + [_P]),
    "mvster_conv2d_small": (c_int, [_P, _P, _P, _P] + [c_int] * 10 + [_P]),
    "mvster_conv3d_mid_tc": (c_int, [_P, _P, _P, _P, _P] + [c_int] * 8 + [_P]),
    "mvster_conv3d_mid_umma": (c_int, [_P, _P, _P, _P] + [c_int] * 8 + [_P]),
    "mvster_umma_pack_weights": (c_int, [_P, _P, c_int, c_int, c_int, _P]),
    "mvster_tf32_split": (c_int, [_P, _P, _P, ctypes.c_longlong, _P]),
    "mvster_fpn_topdown": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P] + [c_int] * 7 + [_P]),
    "mvster_fpn_topdown_ex": (c_int, [_P, _P, _P, _P, _P, c_int, _P, _P, _P] + [c_int] * 7 + [_P]),
    "mvster_fpn_topdown_lin": (c_int, [_P, c_int, c_int, _P, _P, c_int, _P, _P] + [c_int] * 5 + [_P]),
    "mvster_fpn_project_up": (c_int, [_P, c_int, c_int, _P, _P, _P, _P] + [c_int] * 5 + [_P]),
    "mvster_bn_train_workspace_bytes": (ctypes.c_longlong, [c_int, c_int, ctypes.c_longlong]),
    "mvster_bn_train_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_float, c_float, c_int, c_int, c_int,
                                    ctypes.c_longlong, _P, _P]),
    "mvster_bn_train_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, ctypes.c_longlong, _P, _P]),
    "mvster_conv3d_wgrad_workspace_bytes": (ctypes.c_longlong, [c_int] * 7),
    "mvster_conv3d_wgrad": (c_int, [_P, _P, _P] + [c_int] * 10 + [_P, _P]),
    "mvster_conv1x1_wgrad_workspace_bytes": (ctypes.c_longlong, [c_int, c_int, ctypes.c_longlong]),
    "mvster_conv1x1_wgrad": (c_int, [_P, _P, _P, _P, c_int, c_int, ctypes.c_longlong, _P, _P]),
    "mvster_tail_bwd": (c_int, [_P, _P, _P, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, _P]),
    "mvster_geo_check_pair": (c_int, [_P, POINTER(c_double), POINTER(c_double), _P, POINTER(c_double),
                                      POINTER(c_double), c_double, c_double, _P, _P, _P, _P, c_int, c_int, _P]),
    "mvster_geo_filter": (c_int, [_P, _P, POINTER(c_double), POINTER(c_double), POINTER(c_int32), c_int, c_int, c_int,
                                  c_double, c_double, c_double, c_int, _P, _P, _P, _P, _P, c_int, c_int, _P]),
    "mvster_depth2pts": (c_int, [_P, POINTER(c_double), POINTER(c_double), _P, c_int, c_int, _P]),
    "mvster_sinkhorn_blocks": (c_int, [c_int] * 7),
    "mvster_sinkhorn_fwd": (c_int, [_P, _P, _P, _P, c_int, c_float, c_int, c_int, _P, _P, _P, _P, c_int, c_int, c_int,
                                    c_int, _P]),
    "mvster_sinkhorn_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "mvster_nchw_to_nhwc": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))


def library_path() -> str:
    return LIB_PATH


def load(build_if_missing: bool = False) -> ctypes.CDLL:
    """Load (once) and return the shared library; raises RuntimeError when it is not built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("MVSTER_B200_LIB", LIB_PATH)  # experiments: an alternative build of the same ABI
        if path != LIB_PATH:
            lib = ctypes.CDLL(path)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
            return lib
        if not os.path.exists(LIB_PATH):
            if build_if_missing:
                build_library()
            else:
                raise RuntimeError(
                    "libmvster_b200.so is not built (%s). Build it with "
                    "`python -c 'import __graft_entry__ as g; g.build()'` - there is no CPU/PyTorch fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.mvster_version() != 1:
            raise RuntimeError("libmvster_b200.so ABI version %d != 1; rebuild" % lib.mvster_version())
        _lib = lib
        return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().mvster_last_error()
        raise RuntimeError("mvster_b200 error %d: %s" % (status, msg.decode() if msg else "?"))


def launch_count() -> int:
    return int(load().mvster_launch_count())
