#!/bin/bash
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_full3.log 2>&1; tail -6 gpurun_out/r2_pytest_full3.log
( time python bench.py ) > gpurun_out/r2_bench_default2.json 2> gpurun_out/r2_bench_default2.err; tail -3 gpurun_out/r2_bench_default2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_default2.json').read().strip().split('\n')[-1]); print('default value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac']); e=d['extra']; print(e['row_sharded_40k']['value'], e['row_sharded_40k']['parity']['passed_1e-9'], e['mc_4096x50']['strong_4096_total']['value'])"
