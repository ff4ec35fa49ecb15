#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "batched_multi or large_line_groups or launch_strategies or sweep_kernels or overlapped" > gpurun_out/r2_m64_tests.log 2>&1; tail -6 gpurun_out/r2_m64_tests.log
timeout 300 python bench.py --lines 64 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_10k_m64.json 2> gpurun_out/r2_bench_10k_m64.err; cut -c1-2300 gpurun_out/r2_bench_10k_m64.json; tail -3 gpurun_out/r2_bench_10k_m64.err
timeout 300 python bench.py --lines 16 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_10k_m16.json 2> gpurun_out/r2_bench_10k_m16.err; cut -c1-300 gpurun_out/r2_bench_10k_m16.json
EKF_SWEEP_SHAPE=10 python scripts/ncu_sweep.py 10000 16,32 > gpurun_out/r2_ncu_dmma_plain.log 2>&1 && \
EKF_SWEEP_SHAPE=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_dmma -c 4 -f -o gpurun_out/prof_sweep_dmma_r2 python scripts/ncu_sweep.py 10000 16,32 > gpurun_out/r2_ncu_dmma.log 2>&1
tail -4 gpurun_out/r2_ncu_dmma.log
python bench.py --workload mc --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_mc_plain2.json 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batch_scan -s 3 -c 1 -f -o gpurun_out/prof_batch_r2_v5 python bench.py --workload mc --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_mc_v5.log 2>&1
tail -2 gpurun_out/r2_ncu_mc_v5.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_plain3.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_launches.log 2>&1
tail -2 gpurun_out/r2_ncu_launches.log
