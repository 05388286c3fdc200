#!/bin/bash
# Round 2, visit I: pulse / walk pipeline over sub-windows (two streams), impacts admitted and planned per sub-window.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_resonator_gpu.py tests/test_resonator_tensor_gpu.py tests/test_c5_parity_gpu.py tests/test_tuning_gpu.py tests/test_pipeline_gpu.py tests/test_reference_shim_gpu.py -m gpu -q -x) > gpurun_out/pytest_res.log 2>&1; tail -6 gpurun_out/pytest_res.log
for t in 0 1 2 3 5; do
  ME_WALK_SUBWINDOW_TILES=$t ME_BENCH_DEBUG=1 timeout 300 python bench.py --workload resonator --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_res_sub$t.json 2> gpurun_out/bench_res_sub$t.err; tail -1 gpurun_out/bench_res_sub$t.err
  ME_WALK_SUBWINDOW_TILES=$t ME_BENCH_DEBUG=1 timeout 300 python bench.py --workload resonator --voices 128 --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_res_v128_sub$t.json 2> gpurun_out/bench_res_v128_sub$t.err; tail -1 gpurun_out/bench_res_v128_sub$t.err
done
ME_BENCH_DEBUG=1 timeout 600 python bench.py --workload resonator --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_res.json 2> gpurun_out/bench_res.err; tail -2 gpurun_out/bench_res.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_res.json'))
print(d['ms_per_step'], d['run']['step_breakdown_ms_rank0'], 'e2e ms', d['e2e']['ms_per_step'], d.get('parity',{}).get('slice'))
PY
