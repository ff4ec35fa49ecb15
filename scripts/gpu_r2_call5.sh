#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -s -k "batch or monte" > gpurun_out/r2_batch_tests.log 2>&1; tail -12 gpurun_out/r2_batch_tests.log
timeout 300 python bench.py --workload mc --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2_mc_v4.json 2> gpurun_out/r2_mc_v4.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batch_scan -s 3 -c 1 -f -o gpurun_out/prof_batch_r2_v4 \
   python bench.py --workload mc --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_mc_v4.log 2>&1
cat gpurun_out/r2_mc_v4.json; tail -3 gpurun_out/r2_mc_v4.err
tail -3 gpurun_out/r2_ncu_mc_v4.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_bench_extras_n1.json 2> gpurun_out/r2_bench_extras_n1.err; cat gpurun_out/r2_bench_extras_n1.json; tail -8 gpurun_out/r2_bench_extras_n1.err
