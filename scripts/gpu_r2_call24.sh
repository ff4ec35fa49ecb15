#!/bin/bash
mkdir -p gpurun_out
for v in "" _mcfold1; do
EKF_LIB=slam_ros_b200/libekfcuda$v.so timeout 300 python bench.py --workload mc --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_mc_v11$v.json 2> gpurun_out/r2_mc_v11$v.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_mc_v11$v.json').read().strip().split('\n')[-1]); print('$v mc value',d['value'],'e2e',d['e2e']['value'])"
done
EKF_LIB=slam_ros_b200/libekfcuda_mcfold1.so timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x -k "batch or monte" 2>&1 | tail -2
