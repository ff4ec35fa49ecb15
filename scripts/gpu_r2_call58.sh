#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 300 python bench.py --workload mc --mc-per-gpu 512 --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r2_mc_512.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_mc_512.json').read().strip().split('\n')[-1]); print('mc 512 filters value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"; done
