#!/bin/bash
# Round 2, visit K: walk kernel on packed FP32 with stepped row pointers; (Im, Im, Re, Re) element order.
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" ME_BENCH_DEBUG=1 timeout 300 python bench.py --workload resonator --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name: $(tail -1 gpurun_out/bench_$name.err)"
}
(time timeout 900 python -m pytest tests/test_resonator_gpu.py tests/test_resonator_tensor_gpu.py tests/test_tensor_mix_gpu.py tests/test_c5_parity_gpu.py tests/test_tuning_gpu.py tests/test_pipeline_gpu.py tests/test_reference_shim_gpu.py -m gpu -q) > gpurun_out/pytest_res.log 2>&1; tail -30 gpurun_out/pytest_res.log
run k_sub0 ME_WALK_SUBWINDOW_TILES=0
run k_sub3 ME_WALK_SUBWINDOW_TILES=3
run k_sub2 ME_WALK_SUBWINDOW_TILES=2
run k_sub3_p3 ME_WALK_SUBWINDOW_TILES=3 ME_PULSE_CTAS_PER_SM=3
run k_sub3_p2 ME_WALK_SUBWINDOW_TILES=3 ME_PULSE_CTAS_PER_SM=2
ME_BENCH_DEBUG=1 timeout 300 python bench.py --workload resonator --voices 128 --steps 5 --warmup 3 --no-cpu-baseline --no-parity 2>&1 >/dev/null | tail -1
