"""GPU probe: FMA issue-rate ceilings and a first timing of the resonator kernel on the C5-shaped bank."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesheditor_b200 import ModalBank, impact_event, measure_fp32_fma_rate  # noqa: E402
from oracle import resonator as orc  # noqa: E402

res = {}
for packed in (0, 1):
    res[f"fma_rate_packed{packed}"] = measure_fp32_fma_rate(0, bool(packed), 20)
print(json.dumps(res))

voices, n_modes, seconds = int(os.environ.get("VOICES", 1024)), 500, float(os.environ.get("SECONDS", 2))
frames = int(48000 * seconds) // 512 * 512
modes = orc.make_modes(n_modes, 10.0 * n_modes, shape_scale=float(os.environ.get('SHAPE', 3000)))  # T60_k = 10 s * n/k >= 10 s
t0 = time.time()
bank = ModalBank(48000.0, 0)
for v in range(voices):
    bank.add_modes(modes)
bank.install(0)
print("build+install s", time.time() - t0)
rng = np.random.default_rng(12345)
events, ev_frames = [], []
for v in range(voices):
    events.append(impact_event(v, 1.0))
    ev_frames.append(0)
blocks = frames // 512
strike_blocks = rng.random((blocks, voices)) < (2.0 * 512 / 48000)
for b in range(1, blocks):
    for v in np.nonzero(strike_blocks[b])[0]:
        events.append(impact_event(int(v), float(rng.uniform(0.2, 1.0)), int(rng.integers(0, 4))))
        ev_frames.append(b * 512)
print("events", len(events))
for it in range(3):
    t0 = time.time()
    out = bank.render_offline(events, ev_frames, frames, 512)
    wall = time.time() - t0
    st = bank.stats()
    live = [bank.object_status(v)["LiveModeCount"] for v in range(0, voices, max(1, voices // 16))]
    print("live mode counts (sample):", min(live), max(live))
    print(json.dumps(dict(it=it, wall_s=wall, **st, ms_per_s=st["resonator_kernel_ms"] / seconds, mode_samples_per_s_kernel=st["mode_samples"] / (st["resonator_kernel_ms"] * 1e-3), peak=float(np.abs(out).max()))))
