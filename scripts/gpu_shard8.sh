#!/bin/bash
# 8-GPU: parity of the row-sharded filter (fused exchange, overlapped size), the 40k bench with the fused exchange, and the weak-scaling default
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
  tests/multi_gpu/sharded_check.py 3300 8 fused > gpurun_out/shard8_check.log 2>&1
echo "check rc=$?"; grep "sharded x" gpurun_out/shard8_check.log | head -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29550 \
    bench.py --gpus 8 --workload 40k --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/shard8_bench_fused.json 2> gpurun_out/shard8_bench_fused.err
echo "bench 40k rc=$?"
grep '"metric"' gpurun_out/shard8_bench_fused.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('40k x8 %s: %.3f ms/step value %.1f e2e %.1f sweep %.3f ms (%.0f GB/s) line %.3f' % (d['config']['exchange'], d['ms_per_step'], d['value'], d['e2e']['value'], r['launch_ms'], r['achieved'], r['line_stream_ms_per_step'] or 0))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/weak8_bench.json 2> gpurun_out/weak8_bench.err
echo "bench weak rc=$?"
grep '"metric"' gpurun_out/weak8_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('10k x8 weak: %.3f ms/step value %.1f e2e %.1f sweep %.3f ms (%.0f GB/s)' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['launch_ms'], r['achieved']))"
