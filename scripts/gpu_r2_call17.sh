#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "not monte and not batch" > gpurun_out/r2_lineloop_tests3.log 2>&1; tail -3 gpurun_out/r2_lineloop_tests3.log
EKF_LIB=slam_ros_b200/libekfcuda_timing.so timeout 120 python scripts/line_timing.py 1000 > gpurun_out/r2_line_timing_v5_1000.log 2>&1; tail -3 gpurun_out/r2_line_timing_v5_1000.log
timeout 300 python bench.py --workload 1k --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_1k_v4.json 2> gpurun_out/r2_bench_1k_v4.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1k_v4.json').read().strip().split('\n')[-1]); print('1k value',d['value'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'])"
timeout 300 python bench.py --workload room --steps 1000 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_room.json 2> gpurun_out/r2_bench_room.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_room.json').read().strip().split('\n')[-1]); print('room value',d['value'])"
