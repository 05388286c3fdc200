#!/bin/bash
# Round-end GPU visit at HEAD: -m gpu suite, smoke, bench lines (resonator + reference arm + solve), resonator launch list.
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "exit $?"; cut -c1-260 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference_n1.json 2> /dev/null; cut -c1-200 gpurun_out/bench_reference_n1.json
timeout 600 python bench.py --render-path loop --no-cpu-baseline > gpurun_out/bench_n1_loop.json 2> /dev/null; cut -c1-200 gpurun_out/bench_n1_loop.json
timeout 900 python bench.py --workload solve --steps 3 --warmup 1 > gpurun_out/bench_solve.json 2> gpurun_out/bench_solve.err; echo "exit $?"; cut -c1-200 gpurun_out/bench_solve.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_tensor.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/tensor_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_tensor.csv 2>/dev/null | head -12
