#!/bin/bash
# Warm path + reference-suite tests, block-Lanczos section timers, full ncu capture of the panel sweeps.
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests/test_modal_solve_gpu.py tests/test_reference_suite_gpu.py -x -q) > gpurun_out/pytest_new.log 2>&1; tail -30 gpurun_out/pytest_new.log
ME_PROFILE=1 timeout 600 python scripts/bench_solve.py c3 2>&1 | tail -3 | cut -c1-1200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:WideSweepKernel -s 4 -c 2 -o gpurun_out/widesweep_full python scripts/stage_bench.py 55 1 2 > gpurun_out/widesweep_ncu.log 2>&1
tail -2 gpurun_out/widesweep_ncu.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
