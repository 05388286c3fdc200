#!/bin/bash
# Round 2, visit AA: panel sweeps with warp-owned strips (no barriers inside a run): parity tests, then the in-process A/B on the
# 1M-tet factor (macro blocks off / 4 / 8).
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_cholesky_gpu.py tests/test_modal_solve_gpu.py -m gpu -q -x) > gpurun_out/pytest_chol.log 2>&1; tail -5 gpurun_out/pytest_chol.log
timeout 900 python scripts/gpu_sweep_ab.py 55 1 4 8 1 4 8 2>&1 | tee gpurun_out/sweep_ab.txt
