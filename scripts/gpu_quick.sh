#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cholesky_gpu.py tests/test_modal_solve_gpu.py -x -q 2>&1 | tail -3
python scripts/stage_bench.py 55 1 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('factor_ms','factor_TFLOPs','solve_ms','solve_GBs','spmv_K_GBs','spmv_M_GBs','assemble_GBs')})"
python scripts/bench_solve.py "$@" 2>&1 | tail -4
