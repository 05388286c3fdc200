#!/bin/bash
# Round 2, final visit (1 GPU): the whole -m gpu suite, smoke, both bench arms, launch lists and full captures of the kernels
# the round moved. Outputs under gpurun_out/final_*.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/final_pytest_gpu.log 2>&1; tail -4 $O/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; tail -2 $O/final_smoke.log
(time timeout 900 python bench.py --steps 5 --warmup 3) > $O/final_bench_n1.json 2> $O/final_bench_n1.err; echo "bench exit $?"; tail -4 $O/final_bench_n1.err
(time timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > $O/final_bench_reference_n1.json 2> $O/final_bench_reference_n1.err; cut -c1-200 $O/final_bench_reference_n1.json
timeout 300 python bench.py --workload resonator --render-path loop --steps 3 --warmup 3 --no-cpu-baseline --no-parity > $O/final_bench_n1_loop.json 2>/dev/null
timeout 600 python bench.py --workload solve --steps 6 --warmup 1 --no-cpu-baseline > $O/final_bench_solve.json 2> $O/final_bench_solve.err
ME_RENDER_TRACE=1 timeout 300 python bench.py --workload resonator --steps 1 --warmup 3 --no-cpu-baseline --no-parity 2>&1 >/dev/null | tail -22 > $O/final_trace.txt
# launch lists
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches_resonator.csv python bench.py --workload resonator --steps 1 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/final_launches_solve.csv python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
python scripts/summarize_launches.py $O/final_launches_resonator.csv 2>/dev/null | head -12
python scripts/summarize_launches.py $O/final_launches_solve.csv 2>/dev/null | head -14
# full captures
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"PulseKernel|ResonatorKernel|TensorMixKernel" --launch-skip 30 --launch-count 4 -o $O/final_resonator_full -f python bench.py --workload resonator --steps 1 --warmup 3 --no-cpu-baseline --no-parity > $O/final_ncu_res.log 2>&1; tail -1 $O/final_ncu_res.log | cut -c1-120
ME_WALK_SUBWINDOW_TILES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ResonatorKernel" --launch-skip 3 --launch-count 1 -o $O/final_walk_full -f python bench.py --workload resonator --steps 1 --warmup 3 --no-cpu-baseline --no-parity > $O/final_ncu_walk.log 2>&1; tail -1 $O/final_ncu_walk.log | cut -c1-120
timeout 600 ncu --set full --clock-control none --import-source on -k regex:WideSweepKernel --launch-skip 40 --launch-count 2 -o $O/final_wide_sweep_full -f python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > $O/final_ncu_wide.log 2>&1; tail -1 $O/final_ncu_wide.log | cut -c1-120
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"SyrkScatterKernel|ApplyRotationsKernel" --launch-skip 20 --launch-count 1 -o $O/final_syrk_full -f python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > $O/final_ncu_syrk.log 2>&1; tail -1 $O/final_ncu_syrk.log | cut -c1-120
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ApplyRotationsKernel" --launch-count 1 -o $O/final_rotations_full -f python bench.py --workload solve --steps 1 --warmup 0 --no-cpu-baseline > $O/final_ncu_rot.log 2>&1; tail -1 $O/final_ncu_rot.log | cut -c1-120
ls -la $O | grep final_
