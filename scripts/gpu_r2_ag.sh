#!/bin/bash
# Round 2, visit AG (1 GPU): refresh of the solve-side evidence after the slab-ring sweeps: default bench line, solve bench, launch
# list of one solve, full capture of the two panel sweeps, one panel application's task timeline.
mkdir -p gpurun_out
O=gpurun_out
(time timeout 900 python bench.py --steps 5 --warmup 3) > $O/ag_bench_n1.json 2> $O/ag_bench_n1.err; echo "bench exit $?"; tail -3 $O/ag_bench_n1.err
timeout 600 python bench.py --workload solve --steps 6 --warmup 1 > $O/ag_bench_solve.json 2> $O/ag_bench_solve.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ag_launches_solve.csv python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
python scripts/summarize_launches.py $O/ag_launches_solve.csv 2>/dev/null | head -14
gzip -f $O/ag_launches_solve.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:WideSweepKernel --launch-skip 40 --launch-count 2 -o $O/ag_wide_sweep_full -f python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > $O/ag_ncu_wide.log 2>&1; tail -1 $O/ag_ncu_wide.log | cut -c1-120
ncu -i $O/ag_wide_sweep_full.ncu-rep --page raw --csv > $O/ag_wide_sweep_full_raw.csv 2>/dev/null
ME_SWEEP_TRACE=$O/ag_sweep_trace.bin timeout 600 python scripts/gpu_sweep_ab.py 55 1 2>&1 | tail -1
python scripts/sweep_trace.py $O/ag_sweep_trace.bin > $O/ag_sweep_timeline.txt; rm -f $O/ag_sweep_trace.bin
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ag_bench_n1.json').read().strip().splitlines()[-1])
print('resonator', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
s=d['solve']; print('solve', s['value'], s['roofline']['ms_per_launch'], s['roofline']['frac'], 'batch', d['batch']['value'])
PY
