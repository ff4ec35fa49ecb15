#!/bin/bash
mkdir -p gpurun_out
python bench.py --workload 1k --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_1k_plain.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_1k_plain.json').read().strip().split('\n')[-1]); print('1k value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'line ms',d['roofline']['line_stream_ms_per_step'],'sweep',d['roofline']['launch_ms'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_1k.csv python bench.py --workload 1k --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench_1k.log 2>&1; wc -l gpurun_out/r2_launches_bench_1k.csv
