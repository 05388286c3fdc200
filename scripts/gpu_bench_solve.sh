#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --workload solve --steps 2 --warmup 1 > gpurun_out/bench_solve.json 2> gpurun_out/bench_solve.err; echo "exit $?" >> gpurun_out/bench_solve.err
cat gpurun_out/bench_solve.json; tail -5 gpurun_out/bench_solve.err
timeout 900 python bench.py --workload batch --steps 1 --warmup 0 > gpurun_out/bench_batch.json 2> gpurun_out/bench_batch.err; cat gpurun_out/bench_batch.json; tail -3 gpurun_out/bench_batch.err
