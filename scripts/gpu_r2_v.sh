#!/bin/bash
# Round 2, visit V: the mix kernel on the two-term FP16 split with scaled states.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_tensor_mix_gpu.py -m gpu -q -x) > gpurun_out/pytest_mix.log 2>&1; tail -12 gpurun_out/pytest_mix.log
(time timeout 900 python -m pytest tests/test_resonator_gpu.py tests/test_resonator_tensor_gpu.py tests/test_c5_parity_gpu.py tests/test_tuning_gpu.py tests/test_pipeline_gpu.py tests/test_reference_shim_gpu.py -m gpu -q) > gpurun_out/pytest_res.log 2>&1; tail -12 gpurun_out/pytest_res.log
ME_RENDER_TRACE=1 ME_BENCH_DEBUG=1 timeout 600 python bench.py --workload resonator --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_res.json 2> gpurun_out/bench_res.err; tail -18 gpurun_out/bench_res.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_res.json'))
print(d['ms_per_step'], d['run']['step_breakdown_ms_rank0'], 'e2e ms', d['e2e']['ms_per_step'], d.get('parity',{}).get('slice'))
PY
