#!/bin/bash
# Round 2, visit AE: ring version (single-ticket producer): A/B, the whole GPU suite, the solve bench.
mkdir -p gpurun_out
timeout 600 python scripts/gpu_sweep_ab.py 55 1 8 1 8 2>&1 | tee gpurun_out/sweep_ab.txt
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload solve --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve.json 2> gpurun_out/bench_solve.err; tail -2 gpurun_out/bench_solve.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_solve.json'))
print('solve', d['value'], [round(x,3) for x in d['seconds_each']], {k:round(v,4) if isinstance(v,float) else v for k,v in d['profile'].items()})
print('sweep', d['roofline']['ms_per_launch'], d['roofline']['frac'], [(c['workload'][:24], round(c['value'],3)) for c in d.get('other_configs',[])])
PY
