#!/bin/bash
mkdir -p gpurun_out
for cfg in "8 8" "8 0" "16 8"; do set -- $cfg
EKF_SWEEP_SHAPE=$2 timeout 300 python bench.py --lines $1 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_m$1_shape$2.json 2> gpurun_out/r2_bench_m$1_shape$2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_m$1_shape$2.json').read().strip().split('\n')[-1]); print('10k m$1 shape $2 value',d['value'],'ms',d['ms_per_step'],'line ms',d['roofline']['line_stream_ms_per_step'],'sweep',d['roofline']['launch_ms'])"
done
EKF_LIB=slam_ros_b200/libekfcuda_timing.so timeout 120 python scripts/line_timing.py 10000 > gpurun_out/r2_line_timing_v6_10000.log 2>&1; tail -5 gpurun_out/r2_line_timing_v6_10000.log
EKF_FLAGS_NO_OVERLAP=1 true
