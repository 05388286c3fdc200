#!/bin/bash
# Correctness of the solve path, stage rooflines at three sizes (critical path vs bandwidth), full ncu captures of the sweep and SpMV.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_cholesky_gpu.py tests/test_fem_gpu.py tests/test_modal_solve_gpu.py -x -q) > gpurun_out/pytest_solve.log 2>&1; tail -5 gpurun_out/pytest_solve.log
for n in 20 32 55; do
  timeout 300 python scripts/stage_bench.py $n 1 3 2> gpurun_out/stage_$n.err | tail -1 > gpurun_out/stage_$n.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/stage_$n.json"))
    print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('tets','levels','supernodes','factor_nnz','factor_ms','factor_TFLOPs','solve_ms','solve_GBs','spmv_K_ms','spmv_K_GBs','spmv_M_GBs','assemble_GBs')})
except Exception as e:
    print("stage $n failed", e); print(open("gpurun_out/stage_$n.err").read()[-2000:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:SweepKernel -s 2 -c 2 -o gpurun_out/sweep_full -f python scripts/stage_bench.py 55 1 2 > gpurun_out/sweep_ncu.log 2>&1; tail -2 gpurun_out/sweep_ncu.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:SpmvBsr3Kernel -s 5 -c 1 -o gpurun_out/spmv_full -f python scripts/stage_bench.py 55 1 1 > gpurun_out/spmv_ncu.log 2>&1; tail -2 gpurun_out/spmv_ncu.log | cut -c1-200
ls -la gpurun_out
