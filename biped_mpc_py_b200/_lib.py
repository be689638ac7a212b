"""ctypes binding of ``libbiped_mpc_b200.so`` (C ABI in include/biped_mpc_b200.h)."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint8, c_void_p

from .params import BmpcParams

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libbiped_mpc_b200.so")

EXPORTS = ["bmpc_create", "bmpc_destroy", "bmpc_step", "bmpc_solve", "bmpc_lowlevel", "bmpc_foot_positions", "bmpc_rollout", "bmpc_warm_start",
           "bmpc_debug_assemble", "bmpc_set_option", "bmpc_launch_count", "bmpc_enable_timing", "bmpc_last_timing", "bmpc_measure_fma_peak", "bmpc_last_error",
           "bmpc_abi_version"]

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load():
    """Load the shared library; raise loudly if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH
    if not os.path.exists(path):
        raise LibraryMissing(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). biped_mpc_py_b200 has no CPU fallback.")
    lib = ctypes.CDLL(path)
    dp, ip, up, vp = POINTER(c_double), POINTER(c_int32), POINTER(c_uint8), c_void_p
    lib.bmpc_create.argtypes = [POINTER(BmpcParams), c_int, c_int, POINTER(c_void_p)]
    lib.bmpc_create.restype = c_int
    lib.bmpc_destroy.argtypes = [c_void_p]
    lib.bmpc_destroy.restype = c_int
    # device pointers are passed as integers (tensor.data_ptr())
    lib.bmpc_step.argtypes = [c_void_p, c_int] + [vp] * 15 + [vp]
    lib.bmpc_step.restype = c_int
    lib.bmpc_solve.argtypes = [c_void_p, c_int] + [vp] * 10 + [vp]
    lib.bmpc_solve.restype = c_int
    lib.bmpc_lowlevel.argtypes = [c_void_p, c_int] + [vp] * 8 + [vp]
    lib.bmpc_lowlevel.restype = c_int
    lib.bmpc_foot_positions.argtypes = [c_void_p, c_int, vp, vp, vp, vp]
    lib.bmpc_foot_positions.restype = c_int
    lib.bmpc_rollout.argtypes = [c_void_p, c_int, c_int] + [vp] * 6 + [c_int, c_int] + [vp] * 5 + [vp]
    lib.bmpc_rollout.restype = c_int
    lib.bmpc_warm_start.argtypes = [c_void_p, c_int]
    lib.bmpc_warm_start.restype = c_int
    lib.bmpc_debug_assemble.argtypes = [c_void_p] + [vp] * 7 + [vp]
    lib.bmpc_debug_assemble.restype = c_int
    lib.bmpc_set_option.argtypes = [c_void_p, c_char_p, c_int]
    lib.bmpc_set_option.restype = c_int
    lib.bmpc_launch_count.argtypes = [c_void_p]
    lib.bmpc_launch_count.restype = c_int64
    lib.bmpc_enable_timing.argtypes = [c_void_p, c_int]
    lib.bmpc_enable_timing.restype = c_int
    lib.bmpc_last_timing.argtypes = [c_void_p, POINTER(ctypes.c_float)]
    lib.bmpc_last_timing.restype = c_int
    lib.bmpc_measure_fma_peak.argtypes = [c_int, c_int, POINTER(c_double)]
    lib.bmpc_measure_fma_peak.restype = c_int
    lib.bmpc_last_error.argtypes = []
    lib.bmpc_last_error.restype = c_char_p
    lib.bmpc_abi_version.argtypes = []
    lib.bmpc_abi_version.restype = c_int
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError("biped_mpc_b200: " + load().bmpc_last_error().decode("utf-8", "replace"))
