#!/bin/bash
# Round 2, visit S: depth of the block Lanczos basis.
mkdir -p gpurun_out
for e in 48 64 80 96; do
  ME_LANCZOS_EXTRA=$e ME_PROFILE=1 timeout 600 python bench.py --workload solve --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve_e$e.json 2> gpurun_out/bench_solve_e$e.err
  echo "extra $e: $(grep 'lanczos\] op' gpurun_out/bench_solve_e$e.err | tail -1)"; python -c "
import json; d=json.load(open('gpurun_out/bench_solve_e$e.json')); print('   value', round(d['value'],4), 'iterate', round(d['profile']['iterate'],4), 'ops', d['profile']['op_applications'], 'restarts', d['profile']['restarts'])"
done
