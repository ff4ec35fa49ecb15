#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "changing_line_counts or large_line_groups" ) > gpurun_out/r2_mix_tests.log 2>&1; tail -12 gpurun_out/r2_mix_tests.log
for m in 8 32; do
timeout 300 python bench.py --lines $m --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_m${m}_cls.json 2> gpurun_out/r2_bench_m${m}_cls.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_m${m}_cls.json').read().strip().split('\n')[-1]); print('10k m$m value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'],'sweep',d['roofline']['launch_ms'], 'frac', d['roofline']['frac'])"
done
