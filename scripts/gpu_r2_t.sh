#!/bin/bash
# Round 2, visit T: whole GPU suite, then the solve with the basis accumulation of the host eigensolver on threads.
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
ME_PROFILE=1 timeout 600 python bench.py --workload solve --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "ritz m\|lanczos\] op" | tail -4
ME_HOST_THREADS=1 ME_PROFILE=1 timeout 600 python bench.py --workload solve --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "ritz m\|lanczos\] op" | tail -3
timeout 600 python bench.py --workload solve --steps 6 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve_t.json 2> gpurun_out/bench_solve_t.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_solve_t.json'))
print(d['value'], {k:round(v,4) if isinstance(v,float) else v for k,v in d['profile'].items()}, d['roofline']['frac'], d['roofline']['ms_per_launch'])
PY
nproc
