"""ctypes front-ends for the two resonator oracles. TEST INFRASTRUCTURE ONLY.

RefScene  -> oracle/_ref/libme_ref_audio.so : the unmodified reference ModalAudio.cpp (kind "reference").
PortBank  -> oracle/_build/libme_oracle.so  : this repo's C restatement, oracle/resonator_oracle.c (kind "port").
Both expose the same methods so tests can run either against the CUDA path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libme_ref_audio.so")
PORT_SO = os.path.join(HERE, "_build", "libme_oracle.so")

F32P = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
U32P = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


class Event(C.Structure):
    """Layout of the reference's ModalEvent (ModalAudio.h:28-37)."""

    _fields_ = [("Kind", C.c_uint32), ("Object", C.c_uint32), ("ExPos", C.c_uint32)] + [
        (n, C.c_float) for n in ("Jx", "Jy", "Jz", "PulseStep", "PulseGamma", "AccelAmp", "ClickB0", "ClickA1", "ClickA2")
    ]


def impact_event(obj, impulse, ex_pos=0, pulse_step=1.0 / 300.0, gamma=20.0, accel_amp=0.0, click=(0.0, 0.0, 0.0)):
    """tests/ModalBench.h:42-44 ImpactEvent."""
    return Event(0, obj, ex_pos, impulse, 0.5 * impulse, 0.0, np.float32(pulse_step), gamma, accel_amp, *click)


def build_port(force=False):
    if force or not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < max(
        os.path.getmtime(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith("_oracle.c")
    ):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return PORT_SO


def build_ref():
    """Builds oracle/_ref when /root/reference is present (this container). On the GPU box the prebuilt .so travels."""
    if os.path.isdir("/root/reference/src/audio"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return os.path.exists(REF_SO)


def have_ref():
    return os.path.exists(REF_SO)


def _as_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class _Common:
    MODE_COLS = ["CoeffRe", "CoeffIm", "StateRe", "StateIm", "RadiationGain", "RadiationArea", "OutPhaseIm", "OutPhaseRe", "DeflectionGain", "QuadCompliance", "QuadDriveScale"]
    OBJ_U32 = ["ModeOffset", "ModeCount", "TunedModeCount", "LiveModeCount", "Ringing"]
    OBJ_F32 = ["RadiantRadius", "DeflectionScale", "OutGain", "ListenerGain"]

    def add_modes(self, modes, out_gain=1.0, radius_scale=1.0):
        return self.add_object(modes["freqs"], modes["t60s"], modes["shapes"], modes["positions"], modes["indices"], out_gain, radius_scale)

    def render_blocks(self, blocks, frames=512):
        out = np.zeros(blocks * frames, np.float32)
        for b in range(blocks):
            self.render(out[b * frames:(b + 1) * frames])
        return out


class RefScene(_Common):
    kind = "reference"

    def __init__(self, sample_rate=48000.0, renderers=1):
        L = self.L = C.CDLL(REF_SO)
        L.ref_scene_create.restype = C.c_void_p
        L.ref_scene_create.argtypes = [C.c_float, C.c_uint32]
        L.ref_scene_free.argtypes = [C.c_void_p]
        L.ref_scene_add_object.restype = C.c_uint32
        L.ref_scene_add_object.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, F32P, F32P, F32P, F32P, C.c_void_p, C.c_uint32, C.c_float, C.c_float]
        L.ref_scene_retune.argtypes = [C.c_void_p, C.c_uint32, F32P, F32P, C.c_uint32, C.c_float]
        L.ref_scene_set_gain.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float]
        L.ref_scene_set_click_gain.argtypes = [C.c_void_p, C.c_float]
        L.ref_scene_set_max_impacts.argtypes = [C.c_void_p, C.c_uint32]
        L.ref_scene_install.argtypes = [C.c_void_p, C.c_uint32]
        L.ref_scene_enqueue.argtypes = [C.c_void_p, C.POINTER(Event)]
        L.ref_scene_render.argtypes = [C.c_void_p, F32P, C.c_uint32]
        for f in ("mode_total", "object_count", "active_impacts"):
            getattr(L, "ref_scene_" + f).restype = C.c_uint32
            getattr(L, "ref_scene_" + f).argtypes = [C.c_void_p]
        L.ref_scene_events_dropped.restype = C.c_uint64
        L.ref_scene_events_dropped.argtypes = [C.c_void_p]
        L.ref_scene_get_mode_column.argtypes = [C.c_void_p, C.c_uint32, F32P]
        L.ref_scene_get_object_column_u32.argtypes = [C.c_void_p, C.c_uint32, U32P]
        L.ref_scene_get_object_column_f32.argtypes = [C.c_void_p, C.c_uint32, F32P]
        L.ref_click_filter.argtypes = [C.c_double] * 4 + [F32P]
        self.h = L.ref_scene_create(sample_rate, renderers)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_scene_free(self.h)
            self.h = None

    def add_object(self, freqs, t60s, shapes, positions, indices, out_gain=1.0, radius_scale=1.0):
        freqs, t60s, shapes, positions = map(_as_f32, (freqs, t60s, shapes, positions))
        idx = np.ascontiguousarray(indices, np.uint32)
        return self.L.ref_scene_add_object(self.h, len(freqs), shapes.shape[0], freqs, t60s, shapes, positions, idx.ctypes.data, idx.size, out_gain, radius_scale)

    def retune(self, slot, freqs, t60s, radius_scale=1.0):
        freqs, t60s = _as_f32(freqs), _as_f32(t60s)
        self.L.ref_scene_retune(self.h, slot, freqs, t60s, len(freqs), radius_scale)

    def set_gain(self, slot, out_gain, listener_gain=1.0):
        self.L.ref_scene_set_gain(self.h, slot, out_gain, listener_gain)

    def set_click_gain(self, g):
        self.L.ref_scene_set_click_gain(self.h, g)

    def set_max_impacts(self, n):
        self.L.ref_scene_set_max_impacts(self.h, n)

    def install(self, discard_frames=512):
        self.L.ref_scene_install(self.h, discard_frames)

    def enqueue(self, ev):
        self.L.ref_scene_enqueue(self.h, C.byref(ev))

    def render(self, out):
        self.L.ref_scene_render(self.h, out, out.size)

    def mode_column(self, name):
        out = np.zeros(self.L.ref_scene_mode_total(self.h), np.float32)
        self.L.ref_scene_get_mode_column(self.h, self.MODE_COLS.index(name), out)
        return out

    def object_column(self, name):
        n = self.L.ref_scene_object_count(self.h)
        if name in self.OBJ_U32:
            out = np.zeros(n, np.uint32)
            self.L.ref_scene_get_object_column_u32(self.h, self.OBJ_U32.index(name), out)
        else:
            out = np.zeros(n, np.float32)
            self.L.ref_scene_get_object_column_f32(self.h, self.OBJ_F32.index(name), out)
        return out

    def active_impacts(self):
        return self.L.ref_scene_active_impacts(self.h)

    def events_dropped(self):
        return self.L.ref_scene_events_dropped(self.h)

    def click_filter(self, radius, volume, mass, sample_rate):
        out = np.zeros(3, np.float32)
        self.L.ref_click_filter(radius, volume, mass, sample_rate, out)
        return tuple(float(x) for x in out)


class PortBank(_Common):
    kind = "port"

    def __init__(self, sample_rate=48000.0, renderers=1):
        L = self.L = C.CDLL(build_port())
        L.or_bank_create.restype = C.c_void_p
        L.or_bank_create.argtypes = [C.c_float]
        L.or_bank_free.argtypes = [C.c_void_p]
        L.or_bank_add_object.restype = C.c_uint32
        L.or_bank_add_object.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, F32P, F32P, C.c_void_p, C.c_uint32]
        L.or_bank_tune_object.argtypes = [C.c_void_p, C.c_uint32, F32P, F32P, C.c_uint32, C.c_float]
        L.or_bank_set_gain.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float]
        L.or_bank_set_click_gain.argtypes = [C.c_void_p, C.c_float]
        L.or_bank_set_max_impacts.argtypes = [C.c_void_p, C.c_uint32]
        L.or_bank_set_cull.argtypes = [C.c_void_p, C.c_int]
        L.or_bank_install.argtypes = [C.c_void_p]
        L.or_bank_enqueue.restype = C.c_int
        L.or_bank_enqueue.argtypes = [C.c_void_p, C.POINTER(Event)]
        L.or_bank_render.argtypes = [C.c_void_p, F32P, C.c_uint32]
        L.or_bank_render_exact.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"), C.c_uint32]
        for f in ("mode_total", "object_count", "active_impacts"):
            getattr(L, "or_bank_" + f).restype = C.c_uint32
            getattr(L, "or_bank_" + f).argtypes = [C.c_void_p]
        L.or_bank_events_dropped.restype = C.c_uint64
        L.or_bank_events_dropped.argtypes = [C.c_void_p]
        L.or_bank_get_mode_column.argtypes = [C.c_void_p, C.c_uint32, F32P]
        L.or_bank_get_object_column_u32.argtypes = [C.c_void_p, C.c_uint32, U32P]
        L.or_bank_get_object_column_f32.argtypes = [C.c_void_p, C.c_uint32, F32P]
        self.h = L.or_bank_create(sample_rate)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.or_bank_free(self.h)
            self.h = None

    def add_object(self, freqs, t60s, shapes, positions, indices, out_gain=1.0, radius_scale=1.0):
        freqs, t60s, shapes, positions = map(_as_f32, (freqs, t60s, shapes, positions))
        idx = np.ascontiguousarray(indices, np.uint32)
        slot = self.L.or_bank_add_object(self.h, len(freqs), shapes.shape[0], shapes, positions, idx.ctypes.data, idx.size)
        self.L.or_bank_tune_object(self.h, slot, freqs, t60s, len(freqs), radius_scale)
        self.L.or_bank_set_gain(self.h, slot, out_gain, 1.0)
        return slot

    def retune(self, slot, freqs, t60s, radius_scale=1.0):
        freqs, t60s = _as_f32(freqs), _as_f32(t60s)
        self.L.or_bank_tune_object(self.h, slot, freqs, t60s, len(freqs), radius_scale)

    def set_gain(self, slot, out_gain, listener_gain=1.0):
        self.L.or_bank_set_gain(self.h, slot, out_gain, listener_gain)

    def set_click_gain(self, g):
        self.L.or_bank_set_click_gain(self.h, g)

    def set_max_impacts(self, n):
        self.L.or_bank_set_max_impacts(self.h, n)

    def set_cull(self, cull):
        self.L.or_bank_set_cull(self.h, int(cull))

    def install(self, discard_frames=512):
        self.L.or_bank_install(self.h)
        if discard_frames:
            self.render(np.zeros(discard_frames, np.float32))

    def enqueue(self, ev):
        return self.L.or_bank_enqueue(self.h, C.byref(ev))

    def render(self, out):
        self.L.or_bank_render(self.h, out, out.size)

    def render_exact(self, out):
        """FP64 arbiter: the same recurrence over the same float32 parameters with states and sums in double. A bank is
        rendered EITHER with render() OR with render_exact() for its whole life (the two keep separate states)."""
        self.L.or_bank_render_exact(self.h, out, out.size)

    def mode_column(self, name):
        out = np.zeros(self.L.or_bank_mode_total(self.h), np.float32)
        self.L.or_bank_get_mode_column(self.h, self.MODE_COLS.index(name), out)
        return out

    def object_column(self, name):
        n = self.L.or_bank_object_count(self.h)
        if name in self.OBJ_U32:
            out = np.zeros(n, np.uint32)
            self.L.or_bank_get_object_column_u32(self.h, self.OBJ_U32.index(name), out)
        else:
            out = np.zeros(n, np.float32)
            self.L.or_bank_get_object_column_f32(self.h, self.OBJ_F32.index(name), out)
        return out

    def active_impacts(self):
        return self.L.or_bank_active_impacts(self.h)

    def events_dropped(self):
        return self.L.or_bank_events_dropped(self.h)


def make_modes(mode_count, longest_t60, shape_scale=1.0, sample_points=4):
    """tests/ModalBench.h:18-40 MakeModes + SampleStrip, in float32 like the reference."""
    f32 = np.float32
    k = np.arange(1, mode_count + 1, dtype=np.float32)
    freqs = (f32(40.0) * k) * f32(1.031)
    t60s = f32(longest_t60) / k
    positions = np.zeros((sample_points, 3), np.float32)
    for p in range(sample_points):
        positions[p] = (f32(p) * f32(0.01), 0.0, 0.02 if p % 2 else 0.0)
    indices = np.array([[p, p + 1, p + 2] for p in range(sample_points - 2)], np.uint32).ravel()
    shapes = np.zeros((sample_points, mode_count, 3), np.float32)
    for p in range(sample_points):
        a = (k * f32(0.37) + f32(p)).astype(np.float32)
        # std::sin/std::cos on float arguments: evaluate in float32.
        v = np.stack([np.sin(a), np.cos((a * f32(1.7)).astype(np.float32)), np.sin((a * f32(2.3)).astype(np.float32))], -1).astype(np.float32)
        shapes[p] = (v * f32(0.01)) * f32(shape_scale)
    return dict(freqs=freqs.astype(np.float32), t60s=t60s.astype(np.float32), shapes=shapes, positions=positions, indices=indices)
