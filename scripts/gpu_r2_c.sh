#!/bin/bash
# Round 2, third visit: where the solve's time outside the operator goes (ME_PROFILE section timers, 6 timed solves), the torus
# tests, and the variance between solves.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_reference_suite_gpu.py -m gpu -q -x) > gpurun_out/pytest_torus.log 2>&1; tail -6 gpurun_out/pytest_torus.log
ME_PROFILE=1 timeout 600 python bench.py --workload solve --steps 6 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve_prof.json 2> gpurun_out/bench_solve_prof.err; grep -v "^$" gpurun_out/bench_solve_prof.err | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_solve_prof.json'))
print('solve', d['value'], [round(x,3) for x in d['seconds_each']])
PY
timeout 600 python bench.py --workload solve --steps 6 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve6.json 2> /dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_solve6.json'))
print('solve (no profile sync)', d['value'], [round(x,3) for x in d['seconds_each']], {k:round(v,4) if isinstance(v,float) else v for k,v in d['profile'].items()})
PY
