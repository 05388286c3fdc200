#!/bin/bash
# Round 2, visit H: plan arena (two async copies per span), pulse tails in the K-step form.
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
ME_BENCH_DEBUG=1 timeout 600 python bench.py --workload resonator --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_res.json 2> gpurun_out/bench_res.err; tail -2 gpurun_out/bench_res.err
ME_BENCH_DEBUG=1 timeout 600 python bench.py --workload resonator --voices 128 --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_res_v128.json 2> gpurun_out/bench_res_v128.err; tail -2 gpurun_out/bench_res_v128.err
python - <<'PY'
import json
for f in ('bench_res','bench_res_v128'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['ms_per_step'], d['run']['step_breakdown_ms_rank0'], 'e2e ms', d['e2e']['ms_per_step'], d.get('parity',{}).get('slice'))
PY
