"""Loads libme_modal.so (the C ABI declared in include/me_modal.h) through ctypes.

There is no CPU fallback and no alternative backend: if the library is missing or there is no CUDA device,
calls fail loudly (ImportError here, MeError with ME_CUDA_ERROR from the compute entry points).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libme_modal.so")

ME_OK, ME_BAD_ARG, ME_CUDA_ERROR, ME_CANCELLED, ME_NOT_CONVERGED, ME_NO_MODES, ME_FACTOR_FAILED, ME_QUEUE_FULL, ME_OUT_OF_MEMORY = range(9)
STATUS_NAMES = ["ME_OK", "ME_BAD_ARG", "ME_CUDA_ERROR", "ME_CANCELLED", "ME_NOT_CONVERGED", "ME_NO_MODES", "ME_FACTOR_FAILED", "ME_QUEUE_FULL", "ME_OUT_OF_MEMORY"]


class MeError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"{STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else status}: {text}")
        self.status = status


class MeModalEvent(C.Structure):
    """include/me_modal.h MeModalEvent == the reference's ModalEvent (ModalAudio.h:28-37)."""

    _fields_ = [("kind", C.c_uint32), ("object", C.c_uint32), ("ex_pos", C.c_uint32)] + [
        (n, C.c_float) for n in ("jx", "jy", "jz", "pulse_step", "pulse_gamma", "accel_amp", "click_b0", "click_a1", "click_a2")
    ]


class MeRenderStats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint32), ("resonator_kernel_ms", C.c_float), ("total_device_ms", C.c_float),
        ("mode_samples", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("time_segments", C.c_uint32), ("scan_fallbacks", C.c_uint32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m mesheditor_b200.build` (nvcc, sm_100a). There is no fallback path.")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, f32, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_int
    L.me_last_error.restype = C.c_char_p
    L.me_build_info.restype = C.c_char_p
    L.me_device_count.restype = i32
    sig = {
        "me_bank_create": [f32, i32, C.POINTER(vp)],
        "me_bank_add_object": [vp, u32, u32, vp, vp, vp, u32, C.POINTER(u32)],
        "me_bank_tune_object": [vp, u32, vp, vp, u32, f32],
        "me_bank_set_object_shapes": [vp, u32, u32, u32, vp],
        "me_bank_set_gain": [vp, u32, f32, f32],
        "me_bank_set_click_gain": [vp, f32],
        "me_bank_set_max_impacts": [vp, u32],
        "me_bank_set_time_segments": [vp, u32],
        "me_bank_install": [vp],
        "me_bank_enqueue": [vp, C.POINTER(MeModalEvent)],
        "me_bank_render": [vp, vp, u32],
        "me_bank_render_offline": [vp, vp, vp, u32, u64, u32, vp],
        "me_bank_render_offline_device": [vp, vp, vp, u32, u64, u32, vp, vp],
        "me_bank_get_mode_column": [vp, i32, vp],
        "me_bank_get_object_layout": [vp, u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(f32)],
        "me_bank_get_object_status": [vp, u32, C.POINTER(u32), C.POINTER(u32)],
        "me_bank_last_render_stats": [vp, C.POINTER(MeRenderStats)],
        "me_measure_fp32_fma_rate": [i32, i32, i32, C.POINTER(C.c_double)],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = i32
    L.me_bank_free.argtypes = [vp]
    L.me_bank_free.restype = None
    for name in ("me_bank_object_count", "me_bank_mode_total", "me_bank_active_impacts"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = u32
    L.me_bank_events_dropped.argtypes = [vp]
    L.me_bank_events_dropped.restype = u64
    _lib = L
    return L


def check(status):
    if status != ME_OK:
        raise MeError(status, lib().me_last_error().decode())
