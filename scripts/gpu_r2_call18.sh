#!/bin/bash
mkdir -p gpurun_out
for lib in libekfcuda.so libekfcuda_mc3.so; do
EKF_LIB=slam_ros_b200/$lib timeout 300 python bench.py --workload mc --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_mc_$lib.json 2> gpurun_out/r2_mc_$lib.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_mc_$lib.json').read().strip().split('\n')[-1]); print('$lib mc value',d['value'],'e2e',d['e2e']['value'])"
done
timeout 300 python bench.py --workload 1k --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_1k_v5.json 2> gpurun_out/r2_bench_1k_v5.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1k_v5.json').read().strip().split('\n')[-1]); print('1k value',d['value'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'])"
