#!/bin/bash
mkdir -p gpurun_out
for cfg in "32 40 256" "32 44 256" "8 40 256" "64 40 256" "32 20 512"; do set -- $cfg
EKF_LINE_THREADS=$3 EKF_LINE_SMS=$2 timeout 300 python bench.py --lines $1 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_m$1_lt$3_$2.json 2> gpurun_out/r2_bench_m$1_lt$3_$2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_m$1_lt$3_$2.json').read().strip().split('\n')[-1]); print('10k m$1 threads $3 sms $2 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'],'sweep',d['roofline']['launch_ms'])"
done
