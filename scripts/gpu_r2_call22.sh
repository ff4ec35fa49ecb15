#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x -k "batch or monte" ) > gpurun_out/r2_mc_tests_v8.log 2>&1; tail -3 gpurun_out/r2_mc_tests_v8.log
timeout 300 python bench.py --workload mc --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_mc_v8.json 2> gpurun_out/r2_mc_v8.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_mc_v8.json').read().strip().split('\n')[-1]); print('mc value',d['value'],'e2e',d['e2e']['value'])"
EKF_LIB=slam_ros_b200/libekfcuda_mctiming.so timeout 200 python scripts/mc_timing.py 4096 > gpurun_out/r2_mc_timing4_4096.log 2>&1; tail -2 gpurun_out/r2_mc_timing4_4096.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batch_scan -s 3 -c 1 -f -o gpurun_out/prof_batch_r2_v8 python bench.py --workload mc --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_mc_v8.log 2>&1; tail -2 gpurun_out/r2_ncu_mc_v8.log
