#!/bin/bash
# Round 2, visit O: QL rotation history applied on the device; solve schedules built behind the numeric factorisation.
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests/test_modal_solve_gpu.py tests/test_cholesky_gpu.py tests/test_reference_suite_gpu.py tests/test_fem_gpu.py -m gpu -q -x) > gpurun_out/pytest_solve.log 2>&1; tail -5 gpurun_out/pytest_solve.log
ME_PROFILE=1 timeout 600 python bench.py --workload solve --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve_o.json 2> gpurun_out/bench_solve_o.err; grep "block lanczos\] op\|\[solve\]" gpurun_out/bench_solve_o.err | tail -4
timeout 600 python bench.py --workload solve --steps 6 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve_o2.json 2> gpurun_out/bench_solve_o2.err
ME_HOST_RITZ=1 timeout 600 python bench.py --workload solve --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve_o3.json 2> gpurun_out/bench_solve_o3.err
python - <<'PY'
import json
for f in ('bench_solve_o2','bench_solve_o3'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['value'], {k:round(v,4) if isinstance(v,float) else v for k,v in d['profile'].items()})
PY
