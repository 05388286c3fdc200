"""A/B of the 8-wide panel sweeps on the 1M-tet factor inside ONE process (box-to-box noise is ~5 %): macro blocks off / on.
Prints per setting the device time of a panel application (CUDA events, 12 repetitions) and the difference between a panel
solve's first column and the single-vector sweep of the same right-hand side."""
import math, os, sys
import numpy as np
sys.path.insert(0, ".")
from mesheditor_b200 import Factor, FemSystem, workloads as wl
from mesheditor_b200.modal import MATERIALS

n = int(sys.argv[1]) if len(sys.argv) > 1 else 55
points, tets = wl.kuhn_block(n, n, n, (0.3, 0.3, 0.3))
fem = FemSystem(points, tets, MATERIALS["Steel"], 1)
dofs = fem.info["dofs"]
b8 = np.asfortranarray(np.random.default_rng(1).standard_normal((dofs, 8)))
for setting in sys.argv[2:] or ["1", "4", "1", "4"]:
    os.environ["ME_MACRO_PANELS"] = setting.rstrip("b")  # "4": four panels per macro block, forward sweep only; "4b": both sweeps
    os.environ.pop("ME_MACRO_BACKWARD", None)
    if setting.endswith("b"):
        os.environ["ME_MACRO_BACKWARD"] = "1"
    f = Factor(fem, -((2 * math.pi * 20.0) ** 2))
    ms = []
    for _ in range(13):
        x8 = f.solve(b8)
        ms.append(f.info["last_solve_device_ms"])
    x1 = f.solve(b8[:, 0].copy())
    fi = f.info
    bytes_ = 16 * fi["factor_nonzeros"] + 16 * dofs * 8
    best, med = min(ms[1:]), float(np.median(ms[1:]))
    print(f"macro {setting}: panel application min {best:.3f} ms median {med:.3f} ms = {bytes_ / (med * 1e-3) / 1e9:.0f} GB/s; factor {fi['factor_device_ms']:.1f} ms; "
          f"|panel - single| / |single| = {np.linalg.norm(x8[:, 0] - x1) / np.linalg.norm(x1):.2e}", flush=True)
    del f
