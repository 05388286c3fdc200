#!/bin/bash
# Quick GPU visit at HEAD: -m gpu suite, smoke, both bench workloads.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "exit $?"; cut -c1-400 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 900 python bench.py --workload solve --steps 3 --warmup 1 > gpurun_out/bench_solve.json 2> gpurun_out/bench_solve.err; echo "exit $?"
cat gpurun_out/bench_solve.json; tail -5 gpurun_out/bench_solve.err
