#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "line_loop_forms or launch_strategies or thousand_steps" > gpurun_out/r2_lineloop_tests2.log 2>&1; tail -4 gpurun_out/r2_lineloop_tests2.log
for N in 1000 10000; do
  EKF_LIB=slam_ros_b200/libekfcuda_timing.so timeout 120 python scripts/line_timing.py $N > gpurun_out/r2_line_timing_v2_$N.log 2>&1; echo "== N=$N"; tail -7 gpurun_out/r2_line_timing_v2_$N.log
done
timeout 300 python bench.py --workload 1k --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_1k_v2.json 2> gpurun_out/r2_bench_1k_v2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1k_v2.json').read().strip().split('\n')[-1]); print('1k value',d['value'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'])"
timeout 300 python bench.py --lines 32 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_m32_v2.json 2> gpurun_out/r2_bench_m32_v2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_m32_v2.json').read().strip().split('\n')[-1]); print('10k m32 value',d['value'],'line ms',d['roofline']['line_stream_ms_per_step'],'sweep',d['roofline']['launch_ms'])"
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_10k_v2.json 2> gpurun_out/r2_bench_10k_v2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_10k_v2.json').read().strip().split('\n')[-1]); print('10k m8 value',d['value'],'line ms',d['roofline']['line_stream_ms_per_step'],'frac',d['roofline']['frac'])"
