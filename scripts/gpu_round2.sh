#!/bin/bash
# One GPU visit at HEAD: -m gpu suite, both bench workloads + reference arm, stage rooflines, ncu launch lists.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "exit $?"; cut -c1-600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference_n1.json 2> gpurun_out/bench_reference_n1.err; echo "exit $?"; cut -c1-300 gpurun_out/bench_reference_n1.json
timeout 900 python bench.py --workload solve --steps 3 --warmup 1 > gpurun_out/bench_solve.json 2> gpurun_out/bench_solve.err; echo "exit $?"
cat gpurun_out/bench_solve.json; tail -5 gpurun_out/bench_solve.err
timeout 600 python scripts/stage_bench.py 55 1 3 > gpurun_out/stage_c3.json 2> gpurun_out/stage_c3.err; cat gpurun_out/stage_c3.json; tail -3 gpurun_out/stage_c3.err
timeout 600 python scripts/bench_solve.py c1 c2 > gpurun_out/bench_c1c2.json 2>&1; cat gpurun_out/bench_c1c2.json | cut -c1-700
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_solve.csv python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/solve_ncu.log 2>&1
tail -2 gpurun_out/solve_ncu.log | cut -c1-300; wc -l gpurun_out/launches_solve.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_resonator.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/res_ncu.log 2>&1
tail -2 gpurun_out/res_ncu.log | cut -c1-300; wc -l gpurun_out/launches_resonator.csv
