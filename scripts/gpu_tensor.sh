#!/bin/bash
# GPU visit for the tensor-core form of the resonator bank: parity tests, then the bench.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_tensor_mix_gpu.py tests/test_resonator_tensor_gpu.py tests/test_resonator_gpu.py -x -q) > gpurun_out/pytest_tensor.log 2>&1; tail -30 gpurun_out/pytest_tensor.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_tensor.json 2> gpurun_out/bench_tensor.err; echo "exit $?"; cut -c1-3000 gpurun_out/bench_tensor.json; tail -5 gpurun_out/bench_tensor.err
