#!/bin/bash
mkdir -p gpurun_out
for v in "" _mcs65 _mcs63 _mcs75 ""; do
EKF_LIB=slam_ros_b200/libekfcuda$v.so timeout 300 python bench.py --workload mc --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_mc_v16$v.json 2> gpurun_out/r2_mc_v16$v.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_mc_v16$v.json').read().strip().split('\n')[-1]); print('$v mc value',d['value'],'e2e',d['e2e']['value'])"
done
