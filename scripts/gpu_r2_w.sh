#!/bin/bash
# Round 2, visit W: full capture of the FP16-split mix kernel.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:TensorMixKernel --launch-skip 3 --launch-count 1 -o gpurun_out/mix_f16_full -f python bench.py --workload resonator --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_mix_f16.log 2>&1; tail -2 gpurun_out/ncu_mix_f16.log | cut -c1-200
