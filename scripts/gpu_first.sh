#!/bin/bash
# One GPU call: parity tests, bench (ours + reference), ncu launch list, one ncu --set full capture of the resonator kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ResonatorKernel -s 2 -c 2 -o gpurun_out/resonator_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err; cat gpurun_out/bench_ref.json
