#!/bin/bash
# Round 2, visit AH: blocked FactorDiagKernel: factor / solve parity tests, factor time at 1M tets, the batch of 64 meshes.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_cholesky_gpu.py tests/test_modal_solve_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x) > gpurun_out/pytest_chol.log 2>&1; tail -4 gpurun_out/pytest_chol.log
timeout 600 python scripts/gpu_sweep_ab.py 55 1 2>&1 | tail -1
timeout 600 python bench.py --workload solve --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve.json 2> gpurun_out/bench_solve.err; tail -2 gpurun_out/bench_solve.err
timeout 600 python bench.py --workload batch --steps 1 --warmup 0 > gpurun_out/bench_batch.json 2> gpurun_out/bench_batch.err; tail -2 gpurun_out/bench_batch.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_solve.json'))
print('solve', d['value'], [round(x,3) for x in d['seconds_each']], {k:round(v,4) if isinstance(v,float) else v for k,v in d['profile'].items()})
print('factor', d['roofline_factor']['ms'], d['roofline_factor']['frac'], [(c['workload'][:24], round(c['value'],3)) for c in d.get('other_configs',[])])
b=json.load(open('gpurun_out/bench_batch.json')); print('batch', b['value'], b['seconds_per_batch'])
PY
