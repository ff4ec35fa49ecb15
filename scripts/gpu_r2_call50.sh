#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batch_scan -s 3 -c 1 -f -o gpurun_out/prof_batch_r2_final python bench.py --workload mc --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_mc_final.log 2>&1; tail -1 gpurun_out/r2_ncu_mc_final.log
python scripts/steps_only.py 10000 3 > gpurun_out/r2_steps_plain2.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_scan_lines2 -c 3 -f -o gpurun_out/prof_scan_lines2_10k python scripts/steps_only.py 10000 3 > gpurun_out/r2_ncu_lines2_10k.log 2>&1; tail -1 gpurun_out/r2_ncu_lines2_10k.log
python scripts/steps_only.py 1000 3 > gpurun_out/r2_steps_plain3.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_scan_lines2 -c 3 -f -o gpurun_out/prof_scan_lines2_1k python scripts/steps_only.py 1000 3 > gpurun_out/r2_ncu_lines2_1k.log 2>&1; tail -1 gpurun_out/r2_ncu_lines2_1k.log
