import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesheditor_b200 import ModalBank, MeModalEvent
from oracle import resonator as orc
def to_me(ev):
    return MeModalEvent(ev.Kind, ev.Object, ev.ExPos, ev.Jx, ev.Jy, ev.Jz, ev.PulseStep, ev.PulseGamma, ev.AccelAmp, ev.ClickB0, ev.ClickA1, ev.ClickA2)
rng = np.random.default_rng(5)
n_obj, blocks = 6, 40
modes = orc.make_modes(96, float(os.environ.get("T60", 1.5)))
events, frames = [], []
for b in range(blocks):
    for obj in range(n_obj):
        if b == 0 or rng.random() < 0.08:
            events.append(orc.impact_event(obj, float(rng.uniform(0.2, 1.0)), int(rng.integers(0, 4)), float(1.0 / rng.integers(40, 1500))))
            frames.append(b * 512)
o = orc.PortBank(48000.0, 1)
if os.environ.get("NOCULL"): o.set_cull(0)
for _ in range(n_obj): o.add_modes(modes)
o.install()
ref = np.zeros(blocks * 512, np.float32)
k = 0
for b in range(blocks):
    while k < len(events) and frames[k] == b * 512:
        o.enqueue(events[k]); k += 1
    o.render(ref[b * 512:(b + 1) * 512])
pk = np.abs(ref).max()
for seg in (1, 2, 5, 0):
    g = ModalBank(48000.0, 0)
    for _ in range(n_obj): g.add_modes(modes)
    g.install()
    g.set_time_segments(seg)
    off = g.render_offline([to_me(e) for e in events], frames, blocks * 512, 512)
    d = np.abs(off - ref)
    print("segments", seg, "err", d.max() / pk, "at", int(d.argmax()), "stats", g.stats()["time_segments"])
    blk = d.reshape(blocks, 512).max(1) / pk
    print("  per-block err:", " ".join("%.0e" % x for x in blk))
