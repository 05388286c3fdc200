#!/bin/bash
# Round 2, last 1-GPU visit: the whole -m gpu suite, smoke, both bench arms, the solve bench, full capture and task timeline of the
# panel sweeps in their final form. Outputs under gpurun_out/f2_*.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/f2_pytest_gpu.log 2>&1; tail -4 $O/f2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/f2_smoke.log 2>&1; tail -2 $O/f2_smoke.log
(time timeout 900 python bench.py --steps 5 --warmup 3) > $O/f2_bench_n1.json 2> $O/f2_bench_n1.err; echo "bench exit $?"; tail -4 $O/f2_bench_n1.err
(time timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > $O/f2_bench_reference_n1.json 2> $O/f2_bench_reference_n1.err; cut -c1-200 $O/f2_bench_reference_n1.json
timeout 600 python bench.py --workload solve --steps 6 --warmup 1 > $O/f2_bench_solve.json 2> $O/f2_bench_solve.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:WideSweepKernel --launch-skip 40 --launch-count 2 -o $O/f2_wide_sweep_full -f python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > $O/f2_ncu_wide.log 2>&1; tail -1 $O/f2_ncu_wide.log | cut -c1-120
ncu -i $O/f2_wide_sweep_full.ncu-rep --page raw --csv > $O/f2_wide_sweep_full_raw.csv 2>/dev/null
ME_SWEEP_TRACE=$O/f2_sweep_trace.bin timeout 600 python scripts/gpu_sweep_ab.py 55 4 2>&1 | tail -1
python scripts/sweep_trace.py $O/f2_sweep_trace.bin > $O/f2_sweep_timeline.txt; rm -f $O/f2_sweep_trace.bin
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f2_bench_n1.json').read().strip().splitlines()[-1])
print('resonator', d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity']['slice']['gpu_vs_exact'], d['parity']['full_config']['gpu_vs_reference'])
s=d['solve']; print('solve', s['value'], s['roofline']['ms_per_launch'], s['roofline']['frac'], 'batch', d['batch']['value'])
PY
