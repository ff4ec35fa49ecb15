#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "large_line_groups or batched_multi or launch_strategies or sweep_probe or interleaved" > gpurun_out/batched_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/batched_tests.log
timeout 600 python scripts/probe_batched.py 10000 8,16,32,64 > gpurun_out/probe_batched_new.log 2>&1; echo "rc=$?"; cat gpurun_out/probe_batched_new.log
EKF_SWEEP_MAXC=8 timeout 600 python scripts/probe_batched.py 10000 16,32,64 > gpurun_out/probe_batched_old.log 2>&1; echo "rc=$?"; cat gpurun_out/probe_batched_old.log
