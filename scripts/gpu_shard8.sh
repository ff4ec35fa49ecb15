#!/bin/bash
# 8-GPU check of the row-sharded mode: parity at 8 ranks (fused exchange, overlapped size), then the 40k bench in both exchange modes.
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
  tests/multi_gpu/sharded_check.py 3300 10 fused > gpurun_out/shard8_check.log 2>&1
echo "check rc=$?"; grep "sharded x" gpurun_out/shard8_check.log; tail -3 gpurun_out/shard8_check.log
for mode in 0 1; do
  EKF_SHARD_NCCL=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2955$mode \
    bench.py --gpus 8 --workload 40k --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/shard8_bench_nccl$mode.json 2> gpurun_out/shard8_bench_nccl$mode.err
  echo "bench nccl=$mode rc=$?"
  grep '"metric"' gpurun_out/shard8_bench_nccl$mode.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('40k x8 %s: %.3f ms/step value %.1f e2e %.1f sweep %.3f ms (%.0f GB/s) line %.3f' % (d['config']['exchange'], d['ms_per_step'], d['value'], d['e2e']['value'], r['launch_ms'], r['achieved'], r['line_stream_ms_per_step'] or 0))"
  tail -2 gpurun_out/shard8_bench_nccl$mode.err
done
