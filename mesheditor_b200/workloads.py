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


# ------------------------------------------------------------------------------------------------ analysis workloads
def kuhn_block(nx, ny, nz, size=(1.0, 1.0, 1.0)):
    """MakeBarTets of the reference's solver test (tests/ModalSolverTest.cpp:38-69): an nx x ny x nz grid of cells, each
    split into six positively oriented tets around its main diagonal. Returns (points f64 [V,3], tets uint32 [T,4])."""
    vy, vz = ny + 1, nz + 1
    i, j, k = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    points = np.stack([size[0] * i / nx, size[1] * j / ny, size[2] * k / nz], -1).reshape(-1, 3).astype(np.float64)
    ci, cj, ck = (a.ravel() for a in np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"))

    def vid(a, b, c):
        return (a * vy + b) * vz + c

    c = np.stack([vid(ci, cj, ck), vid(ci + 1, cj, ck), vid(ci, cj + 1, ck), vid(ci + 1, cj + 1, ck),
                  vid(ci, cj, ck + 1), vid(ci + 1, cj, ck + 1), vid(ci, cj + 1, ck + 1), vid(ci + 1, cj + 1, ck + 1)], -1)
    corners = np.array([[0, 1, 3, 7], [0, 3, 2, 7], [0, 2, 6, 7], [0, 6, 4, 7], [0, 4, 5, 7], [0, 5, 1, 7]])
    return points, c[:, corners].reshape(-1, 4).astype(np.uint32)


def torus_mesh(n_u=208, n_c=13, major=0.10, minor=0.03):
    """Config 2: a solid torus as a swept structured grid — n_c x n_c cells over the circular cross
    section (square grid mapped onto the disc) times n_u periodic cells around the ring — each cell split into six Kuhn
    tets. 6 * n_u * n_c^2 tets (210,912 at the defaults)."""
    g = np.linspace(-1.0, 1.0, n_c + 1)
    sx, sy = np.meshgrid(g, g, indexing="ij")
    dx, dy = sx * np.sqrt(1 - sy * sy / 2), sy * np.sqrt(1 - sx * sx / 2)  # square -> disc
    u = 2 * np.pi * np.arange(n_u) / n_u
    rad = np.broadcast_to(major + minor * dx[None], (n_u, n_c + 1, n_c + 1))
    points = np.stack([rad * np.cos(u)[:, None, None], rad * np.sin(u)[:, None, None], np.broadcast_to(minor * dy[None], rad.shape)], -1).reshape(-1, 3)
    vy = vz = n_c + 1

    def vid(a, b, c):
        return ((a % n_u) * vy + b) * vz + c

    ci, cj, ck = (a.ravel() for a in np.meshgrid(np.arange(n_u), np.arange(n_c), np.arange(n_c), indexing="ij"))
    c = np.stack([vid(ci, cj, ck), vid(ci + 1, cj, ck), vid(ci, cj + 1, ck), vid(ci + 1, cj + 1, ck),
                  vid(ci, cj, ck + 1), vid(ci + 1, cj, ck + 1), vid(ci, cj + 1, ck + 1), vid(ci + 1, cj + 1, ck + 1)], -1)
    corners = np.array([[0, 1, 3, 7], [0, 3, 2, 7], [0, 2, 6, 7], [0, 6, 4, 7], [0, 4, 5, 7], [0, 5, 1, 7]])
    tets = c[:, corners].reshape(-1, 4)
    p = points[tets]
    det = np.einsum("ij,ij->i", p[:, 3] - p[:, 0], np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]))
    flip = det < 0  # TetMesh wants positively oriented tets (src/mesh/TetMesh.h:10-13)
    tets[flip] = tets[flip][:, [0, 2, 1, 3]]
    return points.astype(np.float64), tets.astype(np.uint32)


def config1_mesh():
    """Config 1: the IcoSphere (radius 0.1 m, 4 subdivisions) tetrahedralised by the reference's own tetrahedralizer with
    Quality on; a committed fixture (tests/golden/make_golden.py make_config1), 8,987 tets."""
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "meshes", "icosphere_c1.npz"))
    return z["points"], z["tets"], z["surface"]


def bench_excitations(points, count=10):
    """ModalSolverBench's excitation rule (tests/ModalSolverBench.cpp:220-221): vertices i * V / count."""
    idx = (np.arange(count) * len(points)) // count
    return points[idx].astype(np.float32)


def config4_dims(count=64, lo=10_000, hi=500_000):
    """Config 4: `count` Kuhn blocks with tet counts log-spaced lo..hi (deterministic dims table, cubes of edge n cells)."""
    targets = np.exp(np.linspace(np.log(lo), np.log(hi), count))
    return [max(2, int(round((t / 6.0) ** (1.0 / 3.0)))) for t in targets]


def lpt_assign(costs, world):
    """Biggest-first greedy deal to the least-loaded rank (DealObjects, ModalAudio.cpp:430-461; ModalSolverBench.cpp:475-478)."""
    order = np.argsort(-np.asarray(costs, np.float64), kind="stable")
    load, owner = np.zeros(world), np.zeros(len(costs), np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += costs[i]
    return owner
