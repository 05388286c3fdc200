#!/bin/bash
# One GPU visit: the whole -m gpu suite, the solve bench line, stage rooflines and the ncu launch list of one solve.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --workload solve --steps 2 --warmup 1 > gpurun_out/bench_solve.json 2> gpurun_out/bench_solve.err; echo "exit $?" >> gpurun_out/bench_solve.err
cat gpurun_out/bench_solve.json; tail -5 gpurun_out/bench_solve.err
python scripts/stage_bench.py 55 1 3 > gpurun_out/stage_c3.json 2> gpurun_out/stage_c3.err; cat gpurun_out/stage_c3.json; tail -3 gpurun_out/stage_c3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_solve.csv python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/solve_ncu.log 2>&1
tail -2 gpurun_out/solve_ncu.log | cut -c1-300; wc -l gpurun_out/launches_solve.csv
