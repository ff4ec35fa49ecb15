#!/bin/bash
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_full2.log 2>&1; tail -6 gpurun_out/r2_pytest_full2.log
EKF_LIB=slam_ros_b200/libekfcuda_timing.so timeout 120 python scripts/line_timing.py 1000 > gpurun_out/r2_line_timing_v4_1000.log 2>&1; tail -4 gpurun_out/r2_line_timing_v4_1000.log
timeout 300 python bench.py --workload 1k --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_1k_v3.json 2> gpurun_out/r2_bench_1k_v3.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1k_v3.json').read().strip().split('\n')[-1]); print('1k value',d['value'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'])"
EKF_LINE_LOOP=1 timeout 300 python bench.py --workload 1k --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_1k_v3_ll1.json 2> gpurun_out/r2_bench_1k_v3_ll1.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1k_v3_ll1.json').read().strip().split('\n')[-1]); print('1k LINE_LOOP=1 value',d['value'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'])"
timeout 300 python bench.py --workload mc --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_mc_v6.json 2> gpurun_out/r2_mc_v6.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_mc_v6.json').read().strip().split('\n')[-1]); print('mc value',d['value'],'e2e',d['e2e']['value'])"
timeout 300 python bench.py --lines 32 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_m32_v3.json 2> gpurun_out/r2_bench_m32_v3.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_m32_v3.json').read().strip().split('\n')[-1]); print('10k m32 value',d['value'],'line ms',d['roofline']['line_stream_ms_per_step'],'sweep',d['roofline']['launch_ms'])"
