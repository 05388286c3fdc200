"""Synthetic workloads named by BASELINE.json / SURVEY.md §8(d), shared by bench.py and the tests.

Nothing here computes audio or eigenpairs: these only build inputs (mode tables, strike timelines, tet meshes).
"""
from __future__ import annotations

import numpy as np

from ._lib import MeModalEvent

SAMPLE_RATE = 48000.0
BLOCK = 512


def make_modes(mode_count, longest_t60, shape_scale=1.0, sample_points=4):
    """The reference harness's mode table (tests/ModalBench.h:18-40 MakeModes + SampleStrip), float32."""
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
        v = np.stack([np.sin(a), np.cos((a * f32(1.7)).astype(np.float32)), np.sin((a * f32(2.3)).astype(np.float32))], -1).astype(np.float32)
        shapes[p] = (v * f32(0.01)) * f32(shape_scale)
    return dict(freqs=freqs.astype(np.float32), t60s=t60s.astype(np.float32), shapes=shapes, positions=positions, indices=indices)


def c5_modes(n_modes=500, shape_scale=100.0, t60_floor=20.0):
    """Config 5 voice: MakeModes(500) with every T60 floored at `t60_floor` >= 10 s (T60_k = floor * n / k) and a
    shape scale that keeps every 8-mode chunk above the reference's SilentEnergy, so its audibility culling
    (ModalAudio.cpp:139-146) never removes work and nominal mode-samples equal rendered mode-samples."""
    return make_modes(n_modes, t60_floor * n_modes, shape_scale)


def c5_timeline(voices, frames, sample_rate=SAMPLE_RATE, block=BLOCK, restrike_hz=2.0, seed=12345):
    """One strike per voice at frame 0 plus Poisson re-strikes at `restrike_hz` per voice, quantised to render blocks
    (the reference drains its event queue once per RenderModal block). Returns (events, frames, voice ids)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    blocks = (frames + block - 1) // block
    hits = rng.random((blocks, voices)) < (restrike_hz * block / sample_rate)
    hits[0, :] = True
    impulses = rng.uniform(0.2, 1.0, size=(blocks, voices))
    ex_pos = rng.integers(0, 4, size=(blocks, voices))
    events, ev_frames, ev_voice = [], [], []
    for b in range(blocks):
        for v in np.nonzero(hits[b])[0]:
            impulse = 1.0 if b == 0 else float(impulses[b, v])
            events.append((int(v), impulse, 0 if b == 0 else int(ex_pos[b, v])))
            ev_frames.append(b * block)
            ev_voice.append(int(v))
    return events, np.asarray(ev_frames, np.uint64), np.asarray(ev_voice, np.int64)


def impact(obj, impulse, ex_pos=0, pulse_step=1.0 / 300.0, gamma=20.0):
    """tests/ModalBench.h:42-44 ImpactEvent as an MeModalEvent."""
    return MeModalEvent(0, obj, ex_pos, impulse, 0.5 * impulse, 0.0, np.float32(pulse_step), gamma, 0.0, 0.0, 0.0, 0.0)


def shard_voices(voices, world, rank):
    """LPT deal of identical voices == contiguous even split (DealObjects ModalAudio.cpp:430-461)."""
    lo = voices * rank // world
    hi = voices * (rank + 1) // world
    return lo, hi
