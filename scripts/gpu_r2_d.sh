#!/bin/bash
# Round 2, visit D: does the clock sampler (nvidia-smi -lms 50) cause the solve's step-to-step hiccups? And the reference's own
# ModalRenderTest built against the drop-in shim.
mkdir -p gpurun_out
./oracle/_ref/shim_modal_render_test > gpurun_out/shim_modal_render_test.log 2>&1; echo "shim test exit $?"; tail -6 gpurun_out/shim_modal_render_test.log
for ms in 0 50 500; do
ME_CLOCK_SAMPLE_MS=$ms timeout 600 python bench.py --workload solve --steps 8 --warmup 1 --no-cpu-baseline > gpurun_out/bench_solve_s$ms.json 2> /dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/bench_solve_s$ms.json'))
print('sampler $ms ms:', round(d['value'],3), [round(x,3) for x in d['seconds_each']])
PY
done
