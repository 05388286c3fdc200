#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fem_gpu.py tests/test_cholesky_gpu.py tests/test_modal_solve_gpu.py -x -q > gpurun_out/pytest_analysis.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_analysis.log
tail -5 gpurun_out/pytest_analysis.log
timeout 1200 python scripts/bench_solve.py "$@" > gpurun_out/bench_solve.log 2>&1
cat gpurun_out/bench_solve.log | tail -20
