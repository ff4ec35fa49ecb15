#!/bin/bash
mkdir -p gpurun_out
for cfg in "16 20" "32 20" "64 20" "64 28" "64 36" "32 28" "32 36" "16 28"; do set -- $cfg
EKF_LINE_SMS=$2 timeout 300 python bench.py --lines $1 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_m$1_sms$2.json 2> gpurun_out/r2_bench_m$1_sms$2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_m$1_sms$2.json').read().strip().split('\n')[-1]); print('10k m$1 line_sms $2 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'],'sweep',d['roofline']['launch_ms'])"
done
EKF_CHUNK_ABOVE=16 timeout 300 python bench.py --lines 32 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_m32_chunked.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_m32_chunked.json').read().strip().split('\n')[-1]); print('10k m32 chunked16 value',d['value'],'ms',d['ms_per_step'],'line ms',d['roofline']['line_stream_ms_per_step'],'sweep',d['roofline']['launch_ms'])"
