#!/bin/bash
# Round 2, visit AC: panel sweeps with a producer warp streaming slabs into a shared-memory ring: parity, A/B, timeline.
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_cholesky_gpu.py -m gpu -q -x) > gpurun_out/pytest_chol.log 2>&1; tail -5 gpurun_out/pytest_chol.log
timeout 600 python scripts/gpu_sweep_ab.py 55 1 8 1 8 2>&1 | tee gpurun_out/sweep_ab.txt
ME_SWEEP_TRACE=gpurun_out/sweep_trace_ring_macro8.bin timeout 600 python scripts/gpu_sweep_ab.py 55 8 2>&1 | tail -1
python scripts/sweep_trace.py gpurun_out/sweep_trace_ring_macro8.bin > gpurun_out/sweep_trace_ring_macro8.txt; head -7 gpurun_out/sweep_trace_ring_macro8.txt
gzip -f gpurun_out/sweep_trace_ring_macro8.bin
(time timeout 900 python -m pytest tests/test_modal_solve_gpu.py -m gpu -q -x) > gpurun_out/pytest_solve.log 2>&1; tail -3 gpurun_out/pytest_solve.log
