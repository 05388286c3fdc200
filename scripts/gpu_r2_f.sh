#!/bin/bash
# Round 2, visit F: panel sweeps in runs of slabs; Ritz vectors without the late allocation; tail timers.
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
ME_PROFILE=1 timeout 600 python bench.py --workload solve --steps 8 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve_prof.json 2> gpurun_out/bench_solve_prof.err; grep "tail\|extraction\|solve\]" gpurun_out/bench_solve_prof.err | tail -27
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_solve_prof.json'))
print('solve', d['value'], [round(x,3) for x in d['seconds_each']])
print('sweep', d['roofline']['ms_per_launch'], d['roofline']['frac'], 'op_solve', d['profile']['op_solve'])
PY
timeout 600 python bench.py --workload solve --steps 8 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve8.json 2> /dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_solve8.json'))
print('solve (no profile)', d['value'], [round(x,3) for x in d['seconds_each']])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_solve.csv python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_solve.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_solve.csv 2>/dev/null | head -8
