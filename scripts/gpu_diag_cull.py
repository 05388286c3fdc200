import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesheditor_b200 import ModalBank, MeModalEvent
from oracle import resonator as orc
def to_me(ev):
    return MeModalEvent(ev.Kind, ev.Object, ev.ExPos, ev.Jx, ev.Jy, ev.Jz, ev.PulseStep, ev.PulseGamma, ev.AccelAmp, ev.ClickB0, ev.ClickA1, ev.ClickA2)
for (n_obj, n_modes, blocks, t60) in [(1, 64, 8, 0.2), (16, 64, 32, 0.2), (3, 500, 16, 2.0)]:
    modes = orc.make_modes(n_modes, t60)
    outs = {}
    for name in ("cull", "nocull", "gpu"):
        b = ModalBank(48000.0, 0) if name == "gpu" else orc.PortBank(48000.0, 1)
        if name == "nocull": b.set_cull(0)
        for _ in range(n_obj): b.add_modes(modes)
        b.install()
        for o in range(n_obj):
            ev = orc.impact_event(o, 1.0 + 0.1 * o, ex_pos=o % 4)
            b.enqueue(to_me(ev) if name == "gpu" else ev)
        outs[name] = b.render_blocks(blocks)
    pk = np.abs(outs["cull"]).max()
    print(n_obj, n_modes, blocks, t60, "peak", pk, "gpu-vs-cull", np.abs(outs["gpu"] - outs["cull"]).max() / pk, "gpu-vs-nocull", np.abs(outs["gpu"] - outs["nocull"]).max() / pk, "cull-vs-nocull", np.abs(outs["cull"] - outs["nocull"]).max() / pk)
