#!/bin/bash
# Panel-sweep visit: Cholesky/solve tests, then stage rooflines (single and 8-wide solves) on the 1M-tet mesh.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_cholesky_gpu.py -x -q 2>&1 | tail -5
timeout 600 python scripts/stage_bench.py 55 1 3 > gpurun_out/stage_wide.json 2> gpurun_out/stage_wide.err; cat gpurun_out/stage_wide.json; tail -3 gpurun_out/stage_wide.err
