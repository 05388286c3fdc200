#!/bin/bash
# Block-Lanczos visit: solve tests, then the 1M-tet solve with both forms of the iteration.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_cholesky_gpu.py tests/test_modal_solve_gpu.py -x -q 2>&1 | tail -8
ME_LANCZOS=block timeout 600 python scripts/bench_solve.py c1 c3 2>&1 | tail -3 | cut -c1-900
ME_LANCZOS=single timeout 600 python scripts/bench_solve.py c3 2>&1 | tail -1 | cut -c1-900
