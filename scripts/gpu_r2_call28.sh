#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sweep_kernels or sweep_ring or large_line_groups or batched_multi" ) > gpurun_out/r2_split_tests.log 2>&1; tail -6 gpurun_out/r2_split_tests.log
for sp in 1 0; do
EKF_DMMA_SPLIT=$sp timeout 300 python scripts/probe_batched.py > gpurun_out/r2_sweep_probe_split$sp.log 2>&1; echo split=$sp; grep "sweep m=" gpurun_out/r2_sweep_probe_split$sp.log
done
