#!/bin/bash
# Round 2, first GPU visit: the -m gpu suite (new: configs[4] full-length parity, re-install guard, strike guard), smoke, the default
# bench line (resonator + parity + solve + batch), the reference arm on a short run, launch lists, and `ncu --set full` captures
# of the kernels this round has to move (WideSweepKernel, SyrkScatterKernel, PulseKernel, GramNarrowPartialKernel).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
(time timeout 1500 python -m pytest tests -m gpu -q -s) > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time timeout 900 python bench.py --steps 5 --warmup 3) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -5 gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json
(time timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/bench_reference_n1.json 2> gpurun_out/bench_reference_n1.err; cut -c1-300 gpurun_out/bench_reference_n1.json
# launch lists
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_resonator.csv python bench.py --workload resonator --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_resonator.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_resonator.csv 2>/dev/null | head -14
# full captures (one launch each, late in the run so that the kernels are warm)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:WideSweepKernel --launch-skip 40 --launch-count 2 -o gpurun_out/wide_sweep_full -f python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_wide.log 2>&1; tail -3 gpurun_out/ncu_wide.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:SyrkScatterKernel --launch-skip 20 --launch-count 1 -o gpurun_out/syrk_full -f python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_syrk.log 2>&1; tail -3 gpurun_out/ncu_syrk.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:GramNarrowPartialKernel --launch-skip 300 --launch-count 1 -o gpurun_out/gram_full -f python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_gram.log 2>&1; tail -3 gpurun_out/ncu_gram.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"PulseKernel|ResonatorKernel|TensorMixKernel" --launch-skip 9 --launch-count 3 -o gpurun_out/resonator_full -f python bench.py --workload resonator --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_res_full.log 2>&1; tail -3 gpurun_out/ncu_res_full.log
ls -la gpurun_out | head -40
