#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "not monte and not batch" > gpurun_out/r2_lineloop_tests.log 2>&1; tail -12 gpurun_out/r2_lineloop_tests.log
for v in 1 0; do
EKF_LINE_LOOP=$v timeout 300 python bench.py --workload 1k --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_1k_ll$v.json 2> gpurun_out/r2_bench_1k_ll$v.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1k_ll$v.json').read().strip().split('\n')[-1]); print('1k LINE_LOOP=$v value',d['value'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'])"
EKF_LINE_LOOP=$v timeout 300 python bench.py --lines 32 --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_m32_ll$v.json 2> gpurun_out/r2_bench_m32_ll$v.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_m32_ll$v.json').read().strip().split('\n')[-1]); print('10k m32 LINE_LOOP=$v value',d['value'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'],'sweep',d['roofline']['launch_ms'])"
EKF_LINE_LOOP=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_10k_ll$v.json 2> gpurun_out/r2_bench_10k_ll$v.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_10k_ll$v.json').read().strip().split('\n')[-1]); print('10k m8 LINE_LOOP=$v value',d['value'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'],'frac',d['roofline']['frac'])"
done
