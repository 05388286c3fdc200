"""Python face of the resonator bank C ABI, shaped like the reference's API
(AddModalObject / TuneModalObject / InstallModalBank / EnqueueModalEvent / RenderModal, src/audio/ModalAudio.h:297-315)
so the parity tests read like the reference's own tests (tests/ModalBench.h, tests/ModalRenderTest.cpp)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import ME_QUEUE_FULL, MeModalEvent, MeRenderStats, MeRetune, check, lib

MODE_COLUMNS = ["CoeffRe", "CoeffIm", "StateRe", "StateIm", "RadiationGain", "RadiationArea", "OutPhaseIm", "OutPhaseRe", "DeflectionGain", "QuadCompliance", "QuadDriveScale"]


def impact_event(obj, impulse, ex_pos=0, pulse_step=1.0 / 300.0, gamma=20.0, accel_amp=0.0, click=(0.0, 0.0, 0.0)):
    """tests/ModalBench.h:42-44 ImpactEvent."""
    return MeModalEvent(0, obj, ex_pos, impulse, 0.5 * impulse, 0.0, np.float32(pulse_step), gamma, accel_amp, *click)


def silence_event(obj):
    return MeModalEvent(1, obj, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def uniform_scale_ratio(world_scale, baked_scale=(1.0, 1.0, 1.0)) -> float:
    """UniformScaleRatio (ContactScene.h:97-101); world_scale None = no world transform."""
    world = None if world_scale is None else (C.c_float * 3)(*world_scale)
    return lib().me_uniform_scale_ratio(world, (C.c_float * 3)(*baked_scale))


def listener_gain(distance) -> float:
    """UpdateListenerGains (AudioSystem.cpp:232-243)."""
    return lib().me_listener_gain(distance)


def monitor_frames(frames, sample_rate=48000.0, envelope=0.0):
    """MonitorFrames (AudioSystem.cpp:1177-1189): pressure -> device units under the monitor limiter. Returns (frames, envelope)."""
    out, env = np.array(frames, np.float32).reshape(-1), C.c_float(envelope)
    check(lib().me_monitor_frames(out.ctypes.data, len(out), sample_rate, C.byref(env)))
    return out, env.value


def wav_bytes(frames, sample_rate=48000, normalize_max=None) -> bytes:
    """WriteWav (AudioSystem.cpp:1244-1250): mono float32 RIFF/WAVE, optionally normalised to `normalize_max`."""
    f, out, size = _f32(frames).reshape(-1), C.c_void_p(), C.c_uint64()
    check(lib().me_wav_encode(f.ctypes.data, len(f), int(sample_rate), 0.0 if normalize_max is None else normalize_max, C.byref(out), C.byref(size)))
    try:
        return C.string_at(out, size.value)
    finally:
        lib().me_bytes_free(out)


def wav_frames(data: bytes):
    """The frames and sample rate of a mono float32 (or int16) RIFF/WAVE file."""
    buf, out, n, rate = np.frombuffer(data, np.uint8), C.c_void_p(), C.c_uint64(), C.c_uint32()
    check(lib().me_wav_decode(buf.ctypes.data, len(buf), C.byref(out), C.byref(n), C.byref(rate)))
    try:
        frames = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_float)), (n.value,)).copy() if n.value else np.zeros(0, np.float32)
    finally:
        lib().me_bytes_free(out)
    return frames, rate.value


def retuning(scale=1.0, fundamental=0.0, t60_scale=1.0, alpha=None, modal_level=1.0, gain=1.0) -> MeRetune:
    """What RetuneModalObject (AudioSystem.cpp:263-311) looks up in the scene: size ratio, ModalTuning, the material's Rayleigh alpha
    (None: no AcousticMaterial), ModalControls::ModalLevel, ModalGain::Value."""
    return MeRetune(scale, fundamental, t60_scale, int(alpha is not None), 0.0 if alpha is None else float(alpha), modal_level, gain)


def retune_modes(freqs, t60s, rt: MeRetune):
    """The tuned (freqs, t60s) RetuneModalObject hands to TuneModalObject (AudioSystem.cpp:271, 299-308)."""
    freqs, t60s = _f32(freqs), _f32(t60s)
    n = min(len(freqs), len(t60s))
    out_f, out_t = np.zeros(n, np.float32), np.zeros(n, np.float32)
    check(lib().me_retune_modes(freqs.ctypes.data, t60s.ctypes.data, n, C.byref(rt), out_f.ctypes.data, out_t.ctypes.data))
    return out_f, out_t


def modal_out_gain(rt: MeRetune) -> float:
    """ModalOutGain (AudioSystem.cpp:221-224)."""
    return lib().me_modal_out_gain(C.byref(rt))


class ModalBank:
    """One resonator bank resident on one B200."""

    kind = "cuda"

    def __init__(self, sample_rate=48000.0, device=0):
        self._h = C.c_void_p()
        check(lib().me_bank_create(sample_rate, device, C.byref(self._h)))
        self.sample_rate = sample_rate
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().me_bank_free(self._h)
            self._h = None

    __del__ = close

    # --- building -------------------------------------------------------------------------------------------
    def add_object(self, freqs, t60s, shapes, positions, indices, out_gain=1.0, radius_scale=1.0):
        """AddModalObject + TuneModalObject + OutGain, as ModalScene does (tests/ModalBench.h:59-63)."""
        freqs, t60s, shapes, positions = _f32(freqs), _f32(t60s), _f32(shapes), _f32(positions)
        idx = np.ascontiguousarray(indices, np.uint32)
        slot = C.c_uint32()
        check(lib().me_bank_add_object(self._h, len(freqs), shapes.shape[0], shapes.ctypes.data, positions.ctypes.data, idx.ctypes.data, idx.size, C.byref(slot)))
        self.tune(slot.value, freqs, t60s, radius_scale)
        self.set_gain(slot.value, out_gain, 1.0)
        return slot.value

    def add_modes(self, modes, out_gain=1.0, radius_scale=1.0):
        return self.add_object(modes["freqs"], modes["t60s"], modes["shapes"], modes["positions"], modes["indices"], out_gain, radius_scale)

    def tune(self, slot, freqs, t60s, radius_scale=1.0):
        freqs, t60s = _f32(freqs), _f32(t60s)
        check(lib().me_bank_tune_object(self._h, slot, freqs.ctypes.data, t60s.ctypes.data, min(len(freqs), len(t60s)), radius_scale))

    retune = tune

    def retune_object(self, slot, freqs, t60s, rt: MeRetune):
        """RetuneModalObject (AudioSystem.cpp:263-311): tuned modes -> TuneModalObject at the size ratio -> OutGain."""
        freqs, t60s = _f32(freqs), _f32(t60s)
        check(lib().me_bank_retune_object(self._h, slot, freqs.ctypes.data, t60s.ctypes.data, min(len(freqs), len(t60s)), C.byref(rt)))

    def set_shapes(self, slot, shapes):
        shapes = _f32(shapes)
        check(lib().me_bank_set_object_shapes(self._h, slot, shapes.shape[1], shapes.shape[0], shapes.ctypes.data))

    def set_gain(self, slot, out_gain, listener_gain=1.0):
        check(lib().me_bank_set_gain(self._h, slot, out_gain, listener_gain))

    def set_click_gain(self, g):
        check(lib().me_bank_set_click_gain(self._h, g))

    def set_max_impacts(self, n):
        check(lib().me_bank_set_max_impacts(self._h, n))

    def set_time_segments(self, n):
        check(lib().me_bank_set_time_segments(self._h, n))

    def set_render_path(self, path):
        """0 automatic, 1 FP32 sample loop, 2 tensor-core form wherever the span allows it."""
        check(lib().me_bank_set_render_path(self._h, path))

    def install(self, discard_frames=512):
        """InstallModalBank, then the one discard block ModalScene renders (tests/ModalBench.h:64-69)."""
        check(lib().me_bank_install(self._h))
        if discard_frames:
            self.render(np.zeros(discard_frames, np.float32))

    # --- events and rendering ---------------------------------------------------------------------------------
    def enqueue(self, ev):
        status = lib().me_bank_enqueue(self._h, C.byref(ev))
        if status == ME_QUEUE_FULL:
            return False
        check(status)
        return True

    def render(self, out):
        """RenderModal: adds out.size frames into `out` (float32, host)."""
        assert out.dtype == np.float32 and out.flags.c_contiguous
        check(lib().me_bank_render(self._h, out.ctypes.data, out.size))

    def render_blocks(self, blocks, frames=512):
        out = np.zeros(blocks * frames, np.float32)
        for b in range(blocks):
            self.render(out[b * frames:(b + 1) * frames])
        return out

    @staticmethod
    def pack_events(events, frames):
        """Timeline as the C ABI takes it: a contiguous MeModalEvent array and ascending frame numbers."""
        if isinstance(events, tuple) and len(events) == 3:
            return events
        n = len(events)
        arr = (MeModalEvent * max(n, 1))(*events)
        fr = np.ascontiguousarray(frames, np.uint64)
        assert fr.size == n
        return arr, fr, n

    _pack_events = pack_events

    def render_offline(self, events, frames, total_frames, block_frames=512, out=None):
        arr, fr, n = self._pack_events(events, frames)
        if out is None:
            out = np.zeros(total_frames, np.float32)
        check(lib().me_bank_render_offline(self._h, C.cast(arr, C.c_void_p), fr.ctypes.data, n, total_frames, block_frames, out.ctypes.data))
        return out

    def render_offline_device(self, events, frames, total_frames, block_frames, out_ptr, stream_ptr=None):
        """Result left in device memory at `out_ptr` (e.g. torch_tensor.data_ptr()); enqueued on `stream_ptr`."""
        arr, fr, n = self._pack_events(events, frames)
        check(lib().me_bank_render_offline_device(self._h, C.cast(arr, C.c_void_p), fr.ctypes.data, n, total_frames, block_frames, C.c_void_p(out_ptr), C.c_void_p(stream_ptr)))

    # --- introspection ------------------------------------------------------------------------------------------
    def object_count(self):
        return lib().me_bank_object_count(self._h)

    def mode_total(self):
        return lib().me_bank_mode_total(self._h)

    def active_impacts(self):
        return lib().me_bank_active_impacts(self._h)

    def events_dropped(self):
        return lib().me_bank_events_dropped(self._h)

    def mode_column(self, name):
        out = np.zeros(self.mode_total(), np.float32)
        check(lib().me_bank_get_mode_column(self._h, MODE_COLUMNS.index(name), out.ctypes.data))
        return out

    def object_layout(self, slot):
        off, cnt, tuned, radius = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_float()
        check(lib().me_bank_get_object_layout(self._h, slot, C.byref(off), C.byref(cnt), C.byref(tuned), C.byref(radius)))
        return dict(ModeOffset=off.value, ModeCount=cnt.value, TunedModeCount=tuned.value, RadiantRadius=radius.value)

    def object_status(self, slot):
        live, ringing = C.c_uint32(), C.c_uint32()
        check(lib().me_bank_get_object_status(self._h, slot, C.byref(live), C.byref(ringing)))
        return dict(LiveModeCount=live.value, Ringing=ringing.value)

    def stats(self):
        s = MeRenderStats()
        check(lib().me_bank_last_render_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in MeRenderStats._fields_}


def measure_fp32_fma_rate(device=0, packed=True, iters=20):
    out = C.c_double()
    check(lib().me_measure_fp32_fma_rate(device, int(packed), iters, C.byref(out)))  # packed doubles as the probe mode 0..4
    return out.value
