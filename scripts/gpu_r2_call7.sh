#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "sweep_kernels or sweep_ring" > gpurun_out/r2_dmma_tests.log 2>&1; tail -8 gpurun_out/r2_dmma_tests.log
for sh in 0 10; do
  EKF_SWEEP_SHAPE=$sh timeout 300 python scripts/probe_batched.py 10000 8,16,32 > gpurun_out/r2_probe_shape$sh.log 2>&1; cat gpurun_out/r2_probe_shape$sh.log
done
