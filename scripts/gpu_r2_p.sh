#!/bin/bash
mkdir -p gpurun_out
ME_PROFILE=1 timeout 600 python bench.py --workload solve --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "ritz m\|lanczos\] op" | tail -9
ME_HOST_RITZ=1 ME_PROFILE=1 timeout 600 python bench.py --workload solve --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "ritz m\|lanczos\] op" | tail -9
