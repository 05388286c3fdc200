#!/bin/bash
# Full capture of ApplyRotationsKernel (Rayleigh-Ritz rotation history, order 328).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ApplyRotationsKernel --launch-skip 4 --launch-count 1 -o gpurun_out/rot_full -f python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_rot.log 2>&1; tail -1 gpurun_out/ncu_rot.log | cut -c1-100
ncu -i gpurun_out/rot_full.ncu-rep --page raw --csv > gpurun_out/rot_full_raw.csv 2>/dev/null
ncu -i gpurun_out/rot_full.ncu-rep --page source --csv --print-source sass > gpurun_out/rot_full_source.csv 2>/dev/null; wc -l gpurun_out/rot_full_source.csv
