#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batch_scan -s 3 -c 1 -f -o gpurun_out/prof_batch_r2_v11 python bench.py --workload mc --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_mc_v11.log 2>&1; tail -2 gpurun_out/r2_ncu_mc_v11.log
