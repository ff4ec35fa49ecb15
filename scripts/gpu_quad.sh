#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/quad_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/quad_tests.log
timeout 600 python scripts/probe_batched.py 10000 8,16,32,64 > gpurun_out/probe_quad.log 2>&1; echo "rc=$?"; cat gpurun_out/probe_quad.log
EKF_SWEEP_MAXC=16 timeout 600 python scripts/probe_batched.py 10000 32,64 > gpurun_out/probe_quad_c16.log 2>&1; echo "rc=$?"; cat gpurun_out/probe_quad_c16.log
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quad_10k.json 2> gpurun_out/bench_quad_10k.err; echo "rc=$?"; cat gpurun_out/bench_quad_10k.json
