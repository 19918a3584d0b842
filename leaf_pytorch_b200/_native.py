"""ctypes binding of libleafk.so (C ABI in include/leafk.h).

The library is built in-tree by ``leaf_pytorch_b200/csrc/build.sh`` (or ``__graft_entry__.build()``)
into ``leaf_pytorch_b200/lib/libleafk.so``.  There is deliberately NO fallback: if the library is
missing or a call fails, a ``LeafNativeError`` is raised -- the frontend never silently runs torch
ops instead of the sm_100a kernels.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

# LEAFK_LIB (development only): load another build of the same library, e.g. a timing-experiment variant
LIB_PATH = os.environ.get("LEAFK_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libleafk.so")

ALGO_AUTO, ALGO_FP32, ALGO_TC = 0, 1, 2
TC_NOPRUNE = 32        # LEAFK_TC_NOPRUNE: every filter over all taps (no support pruning of the k-steps)
ALGOS = {"auto": ALGO_AUTO, "fp32": ALGO_FP32, "tc": ALGO_TC, "tc_full": ALGO_TC | TC_NOPRUNE}
REUSE_BANKS = 64       # LEAFK_REUSE_BANKS: the workspace still holds the banks of the same parameters (chunked clips)

SYMBOLS = (
    "leafk_version", "leafk_last_error", "leafk_num_frames", "leafk_same_padding",
    "leafk_workspace_bytes", "leafk_forward", "leafk_forward_window", "leafk_backward",
    "leafk_backward_workspace_bytes", "leafk_forward_host", "leafk_launch_count", "leafk_tc_supported", "leafk_profile_begin", "leafk_profile_end", "leafk_profile_k1_clock", "leafk_profile_tc_schedule",
    "leafk_forward_host_async", "leafk_event_create", "leafk_event_destroy", "leafk_event_synchronize",
    "leafk_train_supported", "leafk_train_workspace_bytes", "leafk_forward_train", "leafk_backward_saved",
    "leafk_backward_saved_workspace_bytes", "leafk_async_status", "leafk_status_message", "leafk_peak_divisors", "leafk_clip_minimum", "leafk_describe_plan",
    "leafk_preemp_forward", "leafk_preemp_backward", "leafk_preemp_backward_workspace_bytes", "leafk_instnorm_forward",
    "leafk_instnorm_backward",
)


class LeafNativeError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("kernel", "pool_w", "pool_b", "alpha", "delta", "root", "ema_w")]


class Grads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("kernel", "pool_w", "pool_b", "alpha", "delta", "root", "ema_w")]


class ClipPrep(C.Structure):
    """leafk_clip_prep: per-clip crop start / raw length / peak divisor applied while the kernels stage the waveform."""
    _fields_ = [("start", C.c_void_p), ("length", C.c_void_p), ("divisor", C.c_void_p), ("ld", C.c_longlong),
                ("pad_mode", C.c_int), ("pad_value", C.c_void_p)]


PAD_MODES = {"zero": 0, "wrap": 1, "edge": 2, "min": 3}


class Config(C.Structure):
    _fields_ = [("F", C.c_int), ("K", C.c_int), ("H", C.c_int), ("pcen_floor", C.c_float),
                ("clamp_min", C.c_float), ("compression", C.c_int), ("algo", C.c_int), ("input_format", C.c_int),
                ("output_format", C.c_int), ("prep", C.POINTER(ClipPrep))]


_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise loudly when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise LeafNativeError(
                f"{LIB_PATH} not found: build it with `sh leaf_pytorch_b200/csrc/build.sh` "
                "(needs nvcc, targets sm_100a). There is no CPU / torch fallback for this path.")
        L = C.CDLL(LIB_PATH)
        vp, ll, i, sz = C.c_void_p, C.c_longlong, C.c_int, C.c_size_t
        L.leafk_version.restype = i
        L.leafk_last_error.restype = C.c_char_p
        L.leafk_num_frames.restype = i
        L.leafk_num_frames.argtypes = [i, i, i]
        L.leafk_same_padding.restype = None
        L.leafk_same_padding.argtypes = [i, C.POINTER(i), C.POINTER(i)]
        L.leafk_workspace_bytes.restype = sz
        L.leafk_workspace_bytes.argtypes = [C.POINTER(Config), i, i]
        L.leafk_forward.restype = i
        L.leafk_forward.argtypes = [C.POINTER(Config), C.POINTER(Params), vp, i, i, vp, vp, vp, sz, vp]
        L.leafk_forward_window.restype = i
        L.leafk_forward_window.argtypes = [C.POINTER(Config), C.POINTER(Params), vp, i, ll, ll, ll, i, i, i,
                                           vp, vp, vp, vp, ll, ll, vp, sz, vp]
        L.leafk_backward.restype = i
        L.leafk_backward.argtypes = [C.POINTER(Config), C.POINTER(Params), vp, i, i, vp, vp, C.POINTER(Grads),
                                     vp, vp, sz, vp]
        L.leafk_backward_workspace_bytes.restype = sz
        L.leafk_backward_workspace_bytes.argtypes = [C.POINTER(Config), i, i]
        L.leafk_forward_host.restype = i
        L.leafk_forward_host.argtypes = [C.POINTER(Config), C.POINTER(Params), vp, i, i, vp, i, vp, vp, vp, sz,
                                         vp, vp, vp]
        L.leafk_tc_supported.restype = i
        L.leafk_tc_supported.argtypes = [i, i, i]
        L.leafk_profile_begin.restype = None
        L.leafk_profile_begin.argtypes = []
        L.leafk_profile_end.restype = i
        L.leafk_profile_end.argtypes = [C.POINTER(C.c_float)] * 3
        L.leafk_profile_k1_clock.restype = i
        L.leafk_profile_k1_clock.argtypes = [C.POINTER(Config), i, i, vp, sz, C.POINTER(ll), C.POINTER(ll)]
        L.leafk_profile_tc_schedule.restype = i
        L.leafk_profile_tc_schedule.argtypes = [C.POINTER(Config), i, i, vp, sz, C.POINTER(i), C.POINTER(i), C.POINTER(i),
                                                C.POINTER(i), i]
        L.leafk_forward_host_async.restype = i
        L.leafk_forward_host_async.argtypes = [C.POINTER(Config), C.POINTER(Params), vp, i, i, vp, i, vp, vp, vp, sz,
                                               vp, vp, vp, vp, vp, vp]
        L.leafk_event_create.restype = vp
        L.leafk_event_create.argtypes = []
        L.leafk_event_destroy.restype = None
        L.leafk_event_destroy.argtypes = [vp]
        L.leafk_event_synchronize.restype = i
        L.leafk_event_synchronize.argtypes = [vp]
        L.leafk_train_supported.restype = i
        L.leafk_train_supported.argtypes = [i, i, i]
        L.leafk_train_workspace_bytes.restype = sz
        L.leafk_train_workspace_bytes.argtypes = [C.POINTER(Config), i, i]
        L.leafk_forward_train.restype = i
        L.leafk_forward_train.argtypes = [C.POINTER(Config), C.POINTER(Params), vp, i, i, vp, vp, vp, sz, vp]
        L.leafk_backward_saved_workspace_bytes.restype = sz
        L.leafk_backward_saved_workspace_bytes.argtypes = [C.POINTER(Config), i, i, i]
        L.leafk_backward_saved.restype = i
        L.leafk_backward_saved.argtypes = [C.POINTER(Config), C.POINTER(Params), vp, i, i, vp, vp, C.POINTER(Grads),
                                           vp, vp, sz, vp]
        L.leafk_describe_plan.restype = i
        L.leafk_describe_plan.argtypes = [i, i, i, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i)]
        L.leafk_clip_minimum.restype = i
        L.leafk_clip_minimum.argtypes = [C.POINTER(Config), vp, i, i, vp, vp]
        L.leafk_peak_divisors.restype = i
        L.leafk_peak_divisors.argtypes = [C.POINTER(Config), vp, i, i, i, vp, vp]
        L.leafk_preemp_forward.restype = i
        L.leafk_preemp_forward.argtypes = [vp, vp, i, i, vp, vp]
        L.leafk_preemp_backward_workspace_bytes.restype = sz
        L.leafk_preemp_backward_workspace_bytes.argtypes = []
        L.leafk_preemp_backward.restype = i
        L.leafk_preemp_backward.argtypes = [vp, vp, vp, i, i, vp, vp, vp, sz, vp]
        L.leafk_instnorm_forward.restype = i
        L.leafk_instnorm_forward.argtypes = [vp, ll, i, C.c_float, vp, vp, vp]
        L.leafk_instnorm_backward.restype = i
        L.leafk_instnorm_backward.argtypes = [vp, vp, vp, ll, i, vp, vp]
        L.leafk_async_status.restype = i
        L.leafk_async_status.argtypes = [vp]
        L.leafk_status_message.restype = i
        L.leafk_status_message.argtypes = [i]
        L.leafk_launch_count.restype = ll
        L.leafk_launch_count.argtypes = [i]
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().leafk_last_error().decode("utf-8", "replace")
        raise LeafNativeError(f"{what} failed (code {rc}): {msg}")
