#!/bin/bash
mkdir -p gpurun_out
for cfg in "32 2" "32 16" "8 2" "8 16"; do set -- $cfg
EKF_VERBOSE=1 EKF_L2_PERSIST=$2 timeout 300 python bench.py --lines $1 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_m$1_persist$2.json 2> gpurun_out/r2_bench_m$1_persist$2.err; grep -m1 "L2 window" gpurun_out/r2_bench_m$1_persist$2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_m$1_persist$2.json').read().strip().split('\n')[-1]); print('10k m$1 persist $2 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'],'sweep',d['roofline']['launch_ms'])"
done
