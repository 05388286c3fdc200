#!/bin/bash
# Round 2, visit AI: timeline of one render of a 128-voice bank (one rank's share at 8 GPUs).
mkdir -p gpurun_out
ME_RENDER_TRACE=1 timeout 300 python bench.py --workload resonator --voices 128 --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2> gpurun_out/trace_v128.txt > gpurun_out/bench_v128.json
tail -16 gpurun_out/trace_v128.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v128.json'))
print(d['ms_per_step'], d['run']['step_breakdown_ms_rank0'], 'e2e', d['e2e']['ms_per_step'], d['gpu_launches'])
PY
