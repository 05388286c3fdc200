#!/bin/bash
# Round 2, visit N: short first sub-windows; small banks keep one batch per window.
mkdir -p gpurun_out
run() { # name, args, env...
  local name=$1; local args=$2; shift; shift
  env "$@" ME_BENCH_DEBUG=1 timeout 300 python bench.py --workload resonator --steps 5 --warmup 3 --no-cpu-baseline --no-parity $args > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name: $(tail -1 gpurun_out/bench_$name.err)"
}
(time timeout 900 python -m pytest tests/test_resonator_gpu.py tests/test_resonator_tensor_gpu.py tests/test_tensor_mix_gpu.py tests/test_c5_parity_gpu.py tests/test_tuning_gpu.py tests/test_pipeline_gpu.py tests/test_reference_shim_gpu.py -m gpu -q) > gpurun_out/pytest_res.log 2>&1; tail -5 gpurun_out/pytest_res.log
run n_sub3 "" ME_WALK_SUBWINDOW_TILES=3
run n_sub2 "" ME_WALK_SUBWINDOW_TILES=2
run n_sub4 "" ME_WALK_SUBWINDOW_TILES=4
run n_v128 "--voices 128"
run n_v256 "--voices 256"
run n_v512 "--voices 512"
run n_v512_sub0 "--voices 512" ME_WALK_SUBWINDOW_TILES=0
ME_RENDER_TRACE=1 ME_WALK_SUBWINDOW_TILES=3 timeout 300 python bench.py --workload resonator --steps 1 --warmup 3 --no-cpu-baseline --no-parity 2>&1 >/dev/null | tail -16
