#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
( time python bench.py ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -4 gpurun_out/r2_bench_default.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().split('\n')[-1]); print('default value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'line ms',d['roofline']['line_stream_ms_per_step']); print(json.dumps(d['cpu_baseline'])[:900]); e=d['extra']; print(e['row_sharded_40k']['value'], e['row_sharded_40k']['parity']['passed_1e-9'], e['mc_4096x50']['strong_4096_total']['value'])"
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; cut -c1-600 gpurun_out/r2_bench_ref.json; tail -3 gpurun_out/r2_bench_ref.err
timeout 300 python bench.py --workload 1k --steps 300 --warmup 5 > gpurun_out/r2_bench_1k_final.json 2> gpurun_out/r2_bench_1k_final.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1k_final.json').read().strip().split('\n')[-1]); print('1k value',d['value'],'e2e',d['e2e']['value']); print(json.dumps(d['cpu_baseline'])[:400])"
