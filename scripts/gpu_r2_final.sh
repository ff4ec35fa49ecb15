#!/bin/bash
# round 2, final single-GPU pass: the driver's own sequence (tests, smoke, bench, reference arm) + the side workloads + launch list
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_final_pytest.log 2>&1; tail -5 gpurun_out/r2_final_pytest.log | head -3
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -2 gpurun_out/r2_final_smoke.log
( time python bench.py ) > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -3 gpurun_out/r2_final_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().split('\n')[-1]); print('default value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'launches',d.get('gpu_launches')); e=d['extra']; print(e['row_sharded_40k']['value'], e['row_sharded_40k']['parity']['passed_1e-9'], e['mc_4096x50']['strong_4096_total']['value']); print(json.dumps(d['cpu_baseline'])[:300])"
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; cut -c1-300 gpurun_out/r2_final_ref.json
timeout 300 python bench.py --workload 1k --steps 300 --warmup 5 > gpurun_out/r2_final_1k.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_final_1k.json').read().strip().split('\n')[-1]); print('1k value',d['value'],'e2e',d['e2e']['value'])"
timeout 300 python bench.py --workload mc --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_final_mc.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_final_mc.json').read().strip().split('\n')[-1]); print('mc value',d['value'],'e2e',d['e2e']['value'])"
for m in 16 32 64; do timeout 300 python bench.py --lines $m --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_final_m$m.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_final_m$m.json').read().strip().split('\n')[-1]); print('10k m$m value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_final_ncu.log 2>&1; wc -l gpurun_out/r2_final_launches.csv
