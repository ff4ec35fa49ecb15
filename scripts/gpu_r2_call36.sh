#!/bin/bash
mkdir -p gpurun_out
python scripts/ncu_sweep.py 10000 16,32 > gpurun_out/r2_ncu_dmma_split_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_dmma -c 4 -f -o gpurun_out/prof_sweep_dmma_split python scripts/ncu_sweep.py 10000 16,32 > gpurun_out/r2_ncu_dmma_split.log 2>&1
tail -3 gpurun_out/r2_ncu_dmma_split.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_launch_plain.json 2>/dev/null && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_final.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_bench_final.log 2>&1; wc -l gpurun_out/r2_launches_bench_final.csv
python bench.py --lines 64 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_launch_plain64.json 2>/dev/null && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_m64.csv python bench.py --lines 64 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_bench_m64.log 2>&1; wc -l gpurun_out/r2_launches_bench_m64.csv
