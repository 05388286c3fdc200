"""TEST INFRASTRUCTURE ONLY (oracle). Slices of BASELINE.json configs[4] (1024 voices x 500 modes x 10 s, mesheditor_b200.
workloads.c5_modes / c5_timeline) rendered on the host two ways, for the parity tests and for bench.py's `parity` key:

  * `ref`   - the reference's own RenderModal (oracle/_ref when it is built, else the C restatement that is bit-equal to it),
              float32, 512-frame blocks, the strikes of a block enqueued before it (tests/ModalBench.h:76-80);
  * `exact` - the FP64 arbiter (oracle/resonator_oracle.c or_bank_render_exact): the same recurrence over the same float32
              parameters with states and sums in double. Voices are independent and the mix is linear, so the arbiter renders
              one voice per bank on a thread pool and sums the doubles.

Why both: on this workload nothing decays (T60s of 20 s .. 10 000 s), and the reference's sequential float32 recurrence itself
random-walks ~1e-5 of peak away from the recurrence's exact value within seconds (measured: 0.7e-5 at 1 s, 0.9e-5 at 2 s,
1.1e-5 at 5 s, 1.75e-5 at 10 s on the 8-voice slice). A renderer that is MORE exact than the reference therefore cannot stay
within 1e-5 of peak of the reference for the whole 10 s; the arbiter shows which side moved.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from mesheditor_b200 import workloads as wl

from . import resonator as orc

GATE = 1e-5  # of peak amplitude: BASELINE.json north_star; tests/ModalRenderTest.cpp:36,48


def c5_slice_timeline(voices, frames, total_voices=1024):
    """The strikes of the first `voices` voices of the full 1024-voice timeline (so a slice is a literal part of the bench's input)."""
    events, ev_frames, ev_voice = wl.c5_timeline(total_voices, frames)
    keep = ev_voice < voices
    return [e for e, k in zip(events, keep) if k], ev_frames[keep]


def _render_blocks(bank, events, ev_frames, frames, exact=False, block=wl.BLOCK):
    out, k = np.zeros(frames, np.float64 if exact else np.float32), 0
    for begin in range(0, frames, block):
        while k < len(events) and ev_frames[k] == begin:
            v, impulse, ex = events[k]
            bank.enqueue(orc.impact_event(v, impulse, ex))
            k += 1
        (bank.render_exact if exact else bank.render)(out[begin:min(begin + block, frames)])
    return out


def reference_render(voices, frames, modes=None, threads=1, events=None, ev_frames=None):
    """The reference bank's float32 render of the slice (kind says which build answered)."""
    modes = modes if modes is not None else wl.c5_modes()
    if events is None:
        events, ev_frames = c5_slice_timeline(voices, frames)
    cls = orc.RefScene if orc.have_ref() else orc.PortBank
    bank = cls(wl.SAMPLE_RATE, threads if cls is orc.RefScene else 1)
    for _ in range(voices):
        bank.add_modes(modes)
    bank.install()
    return _render_blocks(bank, events, ev_frames, frames), cls.kind


def exact_render(voices, frames, modes=None, threads=None, events=None, ev_frames=None):
    """The FP64 arbiter's render of the slice, one voice per bank on `threads` host threads."""
    modes = modes if modes is not None else wl.c5_modes()
    if events is None:
        events, ev_frames = c5_slice_timeline(voices, frames)
    orc.build_port()

    def one(v):
        bank = orc.PortBank(wl.SAMPLE_RATE, 1)
        bank.add_modes(modes)
        bank.install()
        mine = [(0, impulse, ex) for (vv, impulse, ex) in events if vv == v]
        mine_frames = np.asarray([f for (vv, _, _), f in zip(events, ev_frames) if vv == v], np.uint64)
        return _render_blocks(bank, mine, mine_frames, frames, exact=True)

    with ThreadPoolExecutor(threads or min(voices, os.cpu_count() or 1)) as pool:
        total = np.zeros(frames, np.float64)
        for part in pool.map(one, range(voices)):
            total += part
    return total


def compare(gpu, ref, exact):
    """The three distances of the parity report, each as a fraction of the reference render's peak amplitude, plus the
    last frame up to which the reference itself is still within half the gate of the exact value."""
    peak = float(np.abs(ref).max())
    ref_drift = np.abs(ref.astype(np.float64) - exact) / peak
    within = np.nonzero(ref_drift > 0.5 * GATE)[0]
    calm = int(within[0]) if len(within) else len(ref)
    g = gpu.astype(np.float64)
    return {
        "peak": peak,
        "gpu_vs_reference": float(np.abs(g - ref).max() / peak),
        "gpu_vs_exact": float(np.abs(g - exact).max() / peak),
        "reference_vs_exact": float(ref_drift.max()),
        "calm_frames": calm,  # frames before the reference's own float32 drift first exceeds half the gate
        "gpu_vs_reference_while_calm": float(np.abs(g[:calm] - ref[:calm]).max() / peak) if calm else 0.0,
    }
