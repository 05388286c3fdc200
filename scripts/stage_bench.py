#!/usr/bin/env python
"""Per-stage device timings against their rooflines: assembly, SpMV (K, M), factorisation, solve.
Usage: python scripts/stage_bench.py <kuhn edge cells> <order> [solves]"""
import json
import math
import sys

import numpy as np

sys.path.insert(0, ".")
from mesheditor_b200 import Factor, FemSystem, measure_fp64_rate  # noqa: E402
from mesheditor_b200 import workloads as wl  # noqa: E402

n, order = int(sys.argv[1]), int(sys.argv[2])
solves = int(sys.argv[3]) if len(sys.argv) > 3 else 3
points, tets = wl.kuhn_block(n, n, n, (0.3, 0.3, 0.3))
fem = FemSystem(points, tets, "Steel", order)
i = fem.info
T, V = len(tets), len(points)
c = 16 if order == 1 else 56
asm_bytes = c * T + 24 * V + 12 * (i["nnz_stiffness"] + i["nnz_mass"]) + 8 * (i["dofs"] + 1)
out = {"tets": T, "order": order, **i, "assemble_alg_bytes": asm_bytes, "assemble_GBs": asm_bytes / (i["assemble_kernel_ms"] * 1e-3) / 1e9}
x = np.random.default_rng(0).standard_normal(i["dofs"])
fem.spmv("K", x, 50)
nnz_full = 9 * i["node_blocks_full"]
out["spmv_K_ms"] = fem.last_spmv_ms
out["spmv_K_alg_bytes"] = 12 * nnz_full + 20 * i["dofs"] + 4
out["spmv_K_GBs"] = out["spmv_K_alg_bytes"] / (fem.last_spmv_ms * 1e-3) / 1e9
fem.spmv("M", x, 50)
out["spmv_M_ms"] = fem.last_spmv_ms
out["spmv_M_alg_bytes"] = 12 * i["node_blocks_full"] + 8 * i["node_count"] // 2 + 16 * i["dofs"]
out["spmv_M_GBs"] = out["spmv_M_alg_bytes"] / (fem.last_spmv_ms * 1e-3) / 1e9
sigma = -((2 * math.pi * 20.0) ** 2)
f = Factor(fem, sigma)
fi = f.info
out.update({"factor_ms": fi["factor_device_ms"], "factor_TFLOPs": fi["factor_flops"] / (fi["factor_device_ms"] * 1e-3) / 1e12, "factor_nnz": fi["factor_nonzeros"], "analyse_s": fi["analyse_seconds"], "levels": fi["levels"], "supernodes": fi["supernodes"]})
b = np.random.default_rng(1).standard_normal(i["dofs"])
for _ in range(solves):
    f.solve(b)
si = f.info
out["solve_ms"] = si["last_solve_device_ms"]
out["solve_alg_bytes"] = 16 * fi["factor_nonzeros"] + 16 * i["dofs"]
out["solve_GBs"] = out["solve_alg_bytes"] / (si["last_solve_device_ms"] * 1e-3) / 1e9
B = np.random.default_rng(2).standard_normal((i["dofs"], 8))
for _ in range(solves):
    X = f.solve(B)
out["solve8_ms"] = f.info["last_solve_device_ms"]
out["solve8_GBs"] = out["solve_alg_bytes"] / (out["solve8_ms"] * 1e-3) / 1e9
x1 = f.solve(B[:, 3].copy())
out["solve8_vs_single_relerr"] = float(np.linalg.norm(X[:, 3] - x1) / np.linalg.norm(x1))
out["dfma_TFLOPs"] = measure_fp64_rate(0, 0, 3) / 1e12
out["dmma_TFLOPs"] = measure_fp64_rate(0, 1, 3) / 1e12
print(json.dumps(out))
