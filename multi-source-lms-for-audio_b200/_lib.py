"""ctypes binding of libvqb_b200.so (C ABI in include/vqb.h).

The shared library is the product; this module only loads it and declares the signatures.  If it has not been built
the import of any compute entry point fails loudly - there is no Python / PyTorch fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvqb_b200.so")

# flags / error codes mirrored from include/vqb.h
PREC_FP32, PREC_BF16, PREC_TF32 = 0x00, 0x01, 0x02
WANT_Q, WANT_RESID = 0x10, 0x20
UNIQUE_ID_BYTES = 128
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "tf32": PREC_TF32}

_c_f32p = C.c_void_p
_SIGNATURES = {
    "vqb_version": (C.c_int, []),
    "vqb_last_error": (C.c_char_p, []),
    "vqb_workspace_bytes": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "vqb_workspace_bytes_bw": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "vqb_forward": (C.c_int, [_c_f32p, _c_f32p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, _c_f32p, _c_f32p,
                              C.c_void_p, C.c_size_t, C.c_void_p]),
    "vqb_finalize": (C.c_int, [_c_f32p, C.c_int, C.c_int, C.c_float, _c_f32p, C.c_void_p]),
    "vqb_backward": (C.c_int, [_c_f32p, _c_f32p, C.c_void_p, _c_f32p, _c_f32p, _c_f32p, _c_f32p, C.c_float, C.c_int, C.c_int,
                               C.c_int64, C.c_int, _c_f32p, _c_f32p, C.c_void_p]),
    "vqb_ema_update": (C.c_int, [_c_f32p, _c_f32p, _c_f32p, _c_f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "vqb_onehot": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, _c_f32p, C.c_void_p]),
    "vqb_gather": (C.c_int, [_c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, _c_f32p, C.c_void_p]),
    "vqb_window_indices": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_void_p, _c_f32p, C.c_void_p]),
    "vqb_forward_host": (C.c_int, [_c_f32p, _c_f32p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, _c_f32p, _c_f32p,
                                   C.c_int, C.c_void_p]),
    "vqb_host_release": (C.c_int, []),
    "vqb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "vqb_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "vqb_allreduce_stats": (C.c_int, [C.c_void_p, _c_f32p, C.c_size_t, C.c_void_p]),
    "vqb_comm_destroy": (C.c_int, [C.c_void_p]),
    "vqb_debug_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "vqb_debug_launch_count": (C.c_longlong, [C.c_int]),
    "vqb_debug_reload_env": (C.c_int, []),
    "vqb_debug_tail3_lanes": (C.c_int, [C.c_int]),
    "vqb_debug_tail3_perm_pos": (C.c_int, [C.c_int, C.c_int]),
    "vqb_debug_kernel_timing": (C.c_int, [C.c_int]),
    "vqb_debug_kernel_time_ms": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "vqb_debug_stage_time_ms": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "vqb_debug_tc_scores": (C.c_int, [_c_f32p, _c_f32p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, _c_f32p, C.c_void_p,
                                      C.c_size_t, C.c_void_p]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


class VqbError(RuntimeError):
    """A non-zero return from libvqb_b200.so (code and vqb_last_error() text)."""

    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed with code {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python multi-source-lms-for-audio_b200/build.py` "
                "(or __graft_entry__.build()).  The vector-quantiser hot path is CUDA-only; there is no fallback.")
        import torch  # noqa: F401  (brings libcudart.so.12 into the process: the library links the CUDA runtime dynamically)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError here means the .so is stale
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(fn: str, rc: int) -> None:
    if rc != 0:
        raise VqbError(fn, rc, lib().vqb_last_error().decode("utf-8", "replace"))


def stats_len(K: int, D: int) -> int:
    return K * (D + 1) + 2
