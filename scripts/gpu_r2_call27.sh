#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_line_groups or batched_multi or launch_strategies" ) > gpurun_out/r2_chunk_tests.log 2>&1; tail -15 gpurun_out/r2_chunk_tests.log
for cfg in "64 16" "64 0" "48 16" "32 0"; do set -- $cfg
EKF_CHUNK=$2 timeout 300 python bench.py --lines $1 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_m$1_chunk$2.json 2> gpurun_out/r2_bench_m$1_chunk$2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_m$1_chunk$2.json').read().strip().split('\n')[-1]); print('10k m$1 chunk$2 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'],'sweep',d['roofline']['launch_ms'])"
done
