#!/bin/bash
# Round 2, visit Y: macro blocks in the panel sweeps (explicit inverses of runs of chain panels): parity tests, then the 1M-tet solve
# with 1 (off), 4 and 8 panels per macro block.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_cholesky_gpu.py tests/test_modal_solve_gpu.py -m gpu -q -x) > gpurun_out/pytest_chol.log 2>&1; tail -5 gpurun_out/pytest_chol.log
for g in 1 4 8; do
  ME_MACRO_PANELS=$g timeout 600 python bench.py --workload solve --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve_macro$g.json 2> gpurun_out/bench_solve_macro$g.err; tail -2 gpurun_out/bench_solve_macro$g.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_solve_macro$g.json'))
print('macro $g solve', d['value'], [round(x,3) for x in d['seconds_each']], {k:round(v,4) if isinstance(v,float) else v for k,v in d['profile'].items()})
print('sweep', d['roofline']['ms_per_launch'], d['roofline']['frac'], 'factor ms', d['roofline_factor']['ms'])
PY
done
