#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x -k "batch or monte" ) > gpurun_out/r2_mc_tests_v20.log 2>&1; head -3 gpurun_out/r2_mc_tests_v20.log
for i in 1 2; do timeout 300 python bench.py --workload mc --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_mc_v20.json 2> gpurun_out/r2_mc_v20.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_mc_v20.json').read().strip().split('\n')[-1]); print('mc value',d['value'],'e2e',d['e2e']['value'])"; done
