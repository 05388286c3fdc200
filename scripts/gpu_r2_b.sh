#!/bin/bash
# Round 2, second visit: jump-constant compensation in the resonator kernels (parity on configs[4]), the panel sweeps'
# release/acquire handshake, the one-wave Gram kernel.
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -s) > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
cat gpurun_out/parity_c5_*.json | tr -d '\n '; echo
(time timeout 600 python bench.py --workload solve --steps 3 --warmup 1 --no-cpu-baseline) > gpurun_out/bench_solve.json 2> gpurun_out/bench_solve.err; tail -3 gpurun_out/bench_solve.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_solve.json'))
print('solve', d['value'], d['seconds_each'], {k:round(v,4) if isinstance(v,float) else v for k,v in d['profile'].items()})
print('sweep', d['roofline']['ms_per_launch'], d['roofline']['frac'], 'single', d['roofline_single_sweep']['ms_per_launch'], 'factor', d['roofline_factor']['ms'])
PY
(time timeout 600 python bench.py --workload resonator --steps 5 --warmup 3 --no-cpu-baseline) > gpurun_out/bench_res.json 2> gpurun_out/bench_res.err; tail -3 gpurun_out/bench_res.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_res.json'))
print('resonator', d['ms_per_step'], d['run']['step_breakdown_ms_rank0'], d['parity']['slice'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_solve.csv python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_solve.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_solve.csv 2>/dev/null | head -16
