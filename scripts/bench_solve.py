#!/usr/bin/env python
"""Times me_modal_solve on the analysis configs of BASELINE.json and prints the SolveProfile columns the reference's
ModalSolverBench prints (tests/ModalSolverBench.cpp:413-420). Usage: python scripts/bench_solve.py c1 c2 c3 [kuhn:N:order:modes]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mesheditor_b200 import mesh2modes, solver_config  # noqa: E402
from mesheditor_b200 import workloads as wl  # noqa: E402


def run(name):
    if name == "c1":
        points, tets, _ = wl.config1_mesh()
        mat, order, modes = "Steel", 2, 30
    elif name == "c2":
        points, tets = wl.torus_mesh()
        mat, order, modes = "Ceramic", 1, 100
    elif name == "c2p2":
        points, tets = wl.torus_mesh()
        mat, order, modes = "Ceramic", 2, 100
    elif name == "c3":
        points, tets = wl.kuhn_block(55, 55, 55, (0.3, 0.3, 0.3))
        mat, order, modes = "Steel", 1, 200
    else:
        _, n, order, modes = name.split(":")
        points, tets = wl.kuhn_block(int(n), int(n), int(n), (0.3, 0.3, 0.3))
        mat, order, modes = "Steel", int(order), int(modes)
    ex = wl.bench_excitations(points)
    cfg = solver_config(num_modes=modes, element_order=order, max_mode_freq=1e9)
    t0 = time.perf_counter()
    r = mesh2modes(points, tets, mat, ex, config=cfg)
    dt = time.perf_counter() - t0
    p = r.profile
    print(json.dumps({"case": name, "tets": len(tets), "order": order, "modes": modes, "status": r.status, "seconds": dt, "f1": float(r.freqs[0]) if len(r.freqs) else None,
                      "kept_modes": len(r.freqs), **{k: (round(v, 5) if isinstance(v, float) else v) for k, v in p.items()}}), flush=True)


if __name__ == "__main__":
    for name in sys.argv[1:] or ["c1"]:
        run(name)
        run(name) if name == "c1" else None
