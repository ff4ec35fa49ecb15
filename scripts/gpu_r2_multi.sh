#!/bin/bash
# round 2, multi-GPU call: the driver's SCALE invocation at N GPUs (default workload + `extra`), then the sharded tests
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2_bench_extras_n$N.json 2> gpurun_out/r2_bench_extras_n$N.err
tail -5 gpurun_out/r2_bench_extras_n$N.err
python - <<PY
import json
for ln in open("gpurun_out/r2_bench_extras_n$N.json"):
    ln = ln.strip()
    if ln.startswith("{"):
        d = json.loads(ln)
        print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
        print(json.dumps(d.get("extra"), indent=1)[:6000])
PY
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r2_sharded_tests_n$N.log 2>&1; tail -5 gpurun_out/r2_sharded_tests_n$N.log
fi
