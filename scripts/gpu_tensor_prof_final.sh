#!/bin/bash
# Launch list of one bench step and one full capture of the tcgen05 mix kernel, for the committed (states-as-A) kernel.
mkdir -p gpurun_out
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_tensor.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/tensor_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_tensor.csv | head -12
timeout 40 ncu --set full --clock-control none --import-source on -k regex:TensorMixKernel -s 2 -c 1 -o gpurun_out/tensor_mix_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/tensor_full1.log 2>&1
ls -la gpurun_out/*.ncu-rep
