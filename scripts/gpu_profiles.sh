#!/bin/bash
# round-1 final ncu evidence (every command has already exited 0 without ncu in this round)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_quad -c 6 -o gpurun_out/prof_sweep_quad_final -f python scripts/ncu_sweep.py 10000 8,16,32 > gpurun_out/ncu_quad_final.log 2>&1; echo "ncu sweep rc=$?"; tail -2 gpurun_out/ncu_quad_final.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_final.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/launches_bench_final.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_scan_lines -c 2 -o gpurun_out/prof_scan_lines_final -f python scripts/steps_only.py 10000 3 > gpurun_out/ncu_lines_final.log 2>&1; echo "ncu lines rc=$?"; tail -2 gpurun_out/ncu_lines_final.log
