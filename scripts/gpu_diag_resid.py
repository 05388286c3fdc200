"""Diagnostic: shift-invert residual of the float32 basis columns of the 1M-tet solve, with and without removing the
rigid-body components of the float32 rounding noise."""
import math, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesheditor_b200 import Factor, FemSystem, mesh2modes, solver_config
from mesheditor_b200 import workloads as wl

points, tets = wl.kuhn_block(55, 55, 55, (0.3, 0.3, 0.3))
ex = wl.bench_excitations(points)
cfg = solver_config(num_modes=200, element_order=1, max_mode_freq=1e9)
for rep in range(2):
    r = mesh2modes(points, tets, "Steel", ex, config=cfg, keep_basis=True)
    lam = r.eigenvalues
    fem = FemSystem(points, tets, "Steel", 1)
    sigma = -((2 * math.pi * 20.0) ** 2)
    factor = Factor(fem, sigma)
    n = len(points)
    R = np.zeros((3 * n, 6))
    for a in range(3):
        R[a::3, a] = 1.0
        e = np.zeros(3); e[a] = 1.0
        R[:, 3 + a] = np.cross(e, points - points.mean(axis=0)).reshape(-1)
    MR = np.stack([fem.spmv("M", R[:, i]) for i in range(6)], axis=1)
    G = R.T @ MR
    for j in [6, 7, 8, 20, 57, 111, 180, 214]:
        x = r.basis[:, j].astype(np.float64)
        kx, mx = fem.spmv("K", x), fem.spmv("M", x)
        z = factor.solve(kx - lam[j] * mx)
        a = np.linalg.norm(z) / np.linalg.norm(x)
        c = np.linalg.solve(G, MR.T @ x)
        xd = x - R @ c
        kx, mx = fem.spmv("K", xd), fem.spmv("M", xd)
        z = factor.solve(kx - lam[j] * mx)
        b = np.linalg.norm(z) / np.linalg.norm(xd)
        print(rep, j, f"lam={lam[j]:.4e} raw={a:.3e} deflated={b:.3e} rigid_coeff={np.abs(c).max():.2e}", flush=True)
    del factor, fem
