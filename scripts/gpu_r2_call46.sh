#!/bin/bash
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_full6.log 2>&1; tail -5 gpurun_out/r2_pytest_full6.log | head -3
for fz in 1 0; do
EKF_FUSE_PREDICT=$fz timeout 300 python bench.py --workload 1k --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r2_1k_fuse$fz.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_1k_fuse$fz.json').read().strip().split('\n')[-1]); print('fuse $fz 1k value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'])"
done
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_10k_v8.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_10k_v8.json').read().strip().split('\n')[-1]); print('10k value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])"
timeout 300 python bench.py --workload room --no-cpu-baseline > gpurun_out/r2_room_v8.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_room_v8.json').read().strip().split('\n')[-1]); print('room value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
