#!/bin/bash
# Round 2, visit E: panel sweeps with counter-based publication; where the solve's sporadic stalls sit (setup / teardown timers).
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
ME_PROFILE=1 timeout 600 python bench.py --workload solve --steps 8 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve_prof.json 2> gpurun_out/bench_solve_prof.err; grep -v "^$" gpurun_out/bench_solve_prof.err | grep -v "op 0" | tail -40
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_solve_prof.json'))
print('solve', d['value'], [round(x,3) for x in d['seconds_each']])
print('sweep', d['roofline']['ms_per_launch'], d['roofline']['frac'], 'op_solve', d['profile']['op_solve'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_solve.csv python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_solve.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_solve.csv 2>/dev/null | head -8
