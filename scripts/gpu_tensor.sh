#!/bin/bash
# GPU visit for the tensor-core form of the resonator bank: parity tests, then the bench.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_tensor.log 2>&1; tail -5 gpurun_out/pytest_tensor.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_tensor.json 2> gpurun_out/bench_tensor.err; echo "exit $?"; cut -c1-3000 gpurun_out/bench_tensor.json; tail -5 gpurun_out/bench_tensor.err
