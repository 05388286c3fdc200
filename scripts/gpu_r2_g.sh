#!/bin/bash
# Round 2, visit G: block cache over the allocator (solve step-to-step stability), operator application issued ahead of the host
# eigensolve, the tensor window without host round trips between its kernels, a 128-voice rank's step on one GPU.
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload solve --steps 8 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve8.json 2> gpurun_out/bench_solve8.err; tail -2 gpurun_out/bench_solve8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_solve8.json'))
print('solve', d['value'], [round(x,3) for x in d['seconds_each']], {k:round(v,4) if isinstance(v,float) else v for k,v in d['profile'].items()})
print('sweep', d['roofline']['ms_per_launch'], d['roofline']['frac'])
PY
ME_PROFILE=1 timeout 600 python bench.py --workload solve --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "lanczos\] op\|setup" | tail -3
ME_BENCH_DEBUG=1 timeout 600 python bench.py --workload resonator --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_res.json 2> gpurun_out/bench_res.err; tail -3 gpurun_out/bench_res.err
ME_BENCH_DEBUG=1 timeout 600 python bench.py --workload resonator --voices 128 --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_res_v128.json 2> gpurun_out/bench_res_v128.err; tail -3 gpurun_out/bench_res_v128.err
python - <<'PY'
import json
for f in ('bench_res','bench_res_v128'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['ms_per_step'], d['run']['step_breakdown_ms_rank0'], 'segments', d['run']['time_segments'], 'e2e ms', d['e2e']['ms_per_step'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v128.csv python bench.py --workload resonator --voices 128 --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_v128.log 2>&1
grep -v "^==" gpurun_out/launches_v128.csv | python -c "
import csv,sys
rows=list(csv.DictReader(sys.stdin))
for r in rows[-14:]: print(r['ID'], r['Kernel Name'][:70], r['Grid Size'], r['Block Size'], r['Metric Value'], r['Metric Unit'])
"
