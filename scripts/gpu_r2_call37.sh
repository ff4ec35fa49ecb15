#!/bin/bash
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_full4.log 2>&1; tail -4 gpurun_out/r2_pytest_full4.log | head -2
for m in 8 16 32 64; do
timeout 300 python bench.py --lines $m --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_m${m}_snap.json 2> gpurun_out/r2_bench_m${m}_snap.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_m${m}_snap.json').read().strip().split('\n')[-1]); print('10k m$m value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'],'sweep',d['roofline']['launch_ms'])"
done
