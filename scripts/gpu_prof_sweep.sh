#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:SweepKernel -s 2 -c 2 -o gpurun_out/sweep_full python scripts/stage_bench.py 55 1 2 > gpurun_out/sweep_ncu.log 2>&1
tail -2 gpurun_out/sweep_ncu.log | cut -c1-200
ls -la gpurun_out/sweep_full.ncu-rep
