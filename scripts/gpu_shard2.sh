#!/bin/bash
# 2-GPU check of the row-sharded mode: parity (NCCL and fused NVLink exchange), then the 40k bench with the fused exchange
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/shard2_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/shard2_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --workload 40k --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/shard2_bench_fused.json 2> gpurun_out/shard2_bench_fused.err
echo "bench rc=$?"; grep '"metric"' gpurun_out/shard2_bench_fused.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('40k x2 %s: %.3f ms/step value %.1f e2e %.1f sweep %.3f ms (%.0f GB/s) line %.3f' % (d['config']['exchange'], d['ms_per_step'], d['value'], d['e2e']['value'], r['launch_ms'], r['achieved'], r['line_stream_ms_per_step'] or 0))"
