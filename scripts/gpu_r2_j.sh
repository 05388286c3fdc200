#!/bin/bash
# Round 2, visit J: stream priorities (walk over pulses) and CTAs-per-SM limits of the two kernels that run side by side.
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" ME_BENCH_DEBUG=1 timeout 300 python bench.py --workload resonator --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name: $(tail -1 gpurun_out/bench_$name.err)"
}
(time timeout 900 python -m pytest tests/test_resonator_gpu.py tests/test_resonator_tensor_gpu.py tests/test_c5_parity_gpu.py tests/test_tuning_gpu.py tests/test_pipeline_gpu.py tests/test_reference_shim_gpu.py -m gpu -q -x) > gpurun_out/pytest_res.log 2>&1; tail -30 gpurun_out/pytest_res.log
run prio_sub3 ME_WALK_SUBWINDOW_TILES=3
run prio_sub2 ME_WALK_SUBWINDOW_TILES=2
run prio_sub3_p2 ME_WALK_SUBWINDOW_TILES=3 ME_PULSE_CTAS_PER_SM=2
run prio_sub3_p3 ME_WALK_SUBWINDOW_TILES=3 ME_PULSE_CTAS_PER_SM=3
run prio_sub3_p2_w1 ME_WALK_SUBWINDOW_TILES=3 ME_PULSE_CTAS_PER_SM=2 ME_WALK_CTAS_PER_SM=1
run prio_sub3_p3_w1 ME_WALK_SUBWINDOW_TILES=3 ME_PULSE_CTAS_PER_SM=3 ME_WALK_CTAS_PER_SM=1
run prio_sub3_w1 ME_WALK_SUBWINDOW_TILES=3 ME_WALK_CTAS_PER_SM=1
run prio_sub0_w1 ME_WALK_SUBWINDOW_TILES=0 ME_WALK_CTAS_PER_SM=1
run prio_sub1_p3_w1 ME_WALK_SUBWINDOW_TILES=1 ME_PULSE_CTAS_PER_SM=3 ME_WALK_CTAS_PER_SM=1
