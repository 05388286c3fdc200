import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["ME_RESONATOR_DEBUG"] = "1"
from oracle import resonator as orc
from test_resonator_gpu import build_pair, to_me, rel_err
o, g = build_pair(orc.PortBank, 1, orc.make_modes(64, 0.2))
g.set_render_path(2)
ev = orc.impact_event(0, 1.0, ex_pos=0)
o.enqueue(ev); g.enqueue(to_me(ev))
for b in range(8):
    a = np.zeros(512, np.float32); c = np.zeros(512, np.float32)
    o.render(a); g.render(c)
    print(b, g.stats(), "err", np.abs(a - c).max(), "peak", np.abs(a).max())
