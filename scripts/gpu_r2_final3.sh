#!/bin/bash
# Round 2, closing 1-GPU visit: the whole -m gpu suite, smoke, the default bench line and the solve bench on the final build.
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/f3_pytest_gpu.log 2>&1; tail -4 $O/f3_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/f3_smoke.log 2>&1; tail -2 $O/f3_smoke.log
(time timeout 900 python bench.py --steps 5 --warmup 3) > $O/f3_bench_n1.json 2> $O/f3_bench_n1.err; echo "bench exit $?"
timeout 600 python bench.py --workload solve --steps 6 --warmup 1 --no-cpu-baseline > $O/f3_bench_solve.json 2> $O/f3_bench_solve.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f3_bench_n1.json').read().strip().splitlines()[-1])
print('resonator', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity']['slice']['gpu_vs_exact'], d['parity']['full_config']['gpu_vs_reference'])
s=d['solve']; print('solve', s['value'], s['roofline']['ms_per_launch'], s['roofline']['frac'], s['roofline']['traffic'], 'batch', d['batch']['value'])
s=json.load(open('gpurun_out/f3_bench_solve.json')); print('solve x6', s['value'], [round(x,3) for x in s['seconds_each']], {k:(round(v,4) if isinstance(v,float) else v) for k,v in s['profile'].items()})
PY
