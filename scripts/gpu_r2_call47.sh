#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x -k "batch or monte" ) > gpurun_out/r2_mc_tests_v15.log 2>&1; head -3 gpurun_out/r2_mc_tests_v15.log
for v in "" _mcfull; do
EKF_LIB=slam_ros_b200/libekfcuda$v.so timeout 300 python bench.py --workload mc --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_mc_v15$v.json 2> gpurun_out/r2_mc_v15$v.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_mc_v15$v.json').read().strip().split('\n')[-1]); print('$v mc value',d['value'],'e2e',d['e2e']['value'])"
done
EKF_LIB=slam_ros_b200/libekfcuda_mctiming.so timeout 200 python scripts/mc_timing.py 148 > gpurun_out/r2_mc_timing7_148.log 2>&1; tail -2 gpurun_out/r2_mc_timing7_148.log
