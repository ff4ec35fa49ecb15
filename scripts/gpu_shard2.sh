#!/bin/bash
# 2-GPU check of the row-sharded mode: parity (NCCL and fused NVLink exchange), then the 40k bench in both modes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/shard2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/shard2_tests.log
tail -15 gpurun_out/shard2_tests.log
for mode in 0 1; do
  EKF_SHARD_NCCL=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --workload 40k --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/shard2_bench_nccl$mode.json 2> gpurun_out/shard2_bench_nccl$mode.err
  echo "bench nccl=$mode rc=$?"
  tail -c 1500 gpurun_out/shard2_bench_nccl$mode.json
  tail -5 gpurun_out/shard2_bench_nccl$mode.err
done
