#!/bin/bash
# 2-GPU check of the row-sharded mode: parity with the NCCL exchange, the fused NVLink exchange, and switching between them
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/shard2_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/shard2_tests.log
