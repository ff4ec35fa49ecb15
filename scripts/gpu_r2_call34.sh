#!/bin/bash
mkdir -p gpurun_out
for cfg in "8 2" "8 3" "16 2" "64 2"; do set -- $cfg
EKF_SWEEP_STAGES=$2 timeout 300 python bench.py --lines $1 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_m$1_stages$2.json 2> gpurun_out/r2_bench_m$1_stages$2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_m$1_stages$2.json').read().strip().split('\n')[-1]); print('10k m$1 stages $2 value',d['value'],'ms',d['ms_per_step'],'line ms',d['roofline']['line_stream_ms_per_step'],'sweep',d['roofline']['launch_ms'])"
done
