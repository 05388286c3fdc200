#!/bin/bash
# Round 2, visit R: small banks in two sub-windows.
mkdir -p gpurun_out
run() { # name, args, env...
  local name=$1; local args=$2; shift; shift
  env "$@" ME_BENCH_DEBUG=1 timeout 300 python bench.py --workload resonator --steps 5 --warmup 3 --no-cpu-baseline --no-parity $args > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name: $(tail -1 gpurun_out/bench_$name.err)"
}
(time timeout 900 python -m pytest tests/test_resonator_gpu.py tests/test_resonator_tensor_gpu.py tests/test_c5_parity_gpu.py tests/test_tuning_gpu.py -m gpu -q -x) > gpurun_out/pytest_res.log 2>&1; tail -3 gpurun_out/pytest_res.log
run r_v128 "--voices 128"
run r_v128_un "--voices 128" ME_SMALL_BANKS_UNPIPED=1
run r_v256 "--voices 256"
run r_v256_un "--voices 256" ME_SMALL_BANKS_UNPIPED=1
run r_v64 "--voices 64"
run r_v64_un "--voices 64" ME_SMALL_BANKS_UNPIPED=1
ME_RENDER_TRACE=1 timeout 300 python bench.py --workload resonator --voices 128 --steps 1 --warmup 3 --no-cpu-baseline --no-parity 2>&1 >/dev/null | tail -9
