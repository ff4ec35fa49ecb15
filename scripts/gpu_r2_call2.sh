#!/bin/bash
# round 2, second GPU call: fp64 tensor-core probe, the rewritten Monte-Carlo kernel, the full-size parity fixtures
set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_probe scripts/dmma_probe.cu && timeout 120 /tmp/dmma_probe > gpurun_out/dmma_probe.log 2>&1; cat gpurun_out/dmma_probe.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "batch_filters" > gpurun_out/r2_batch_small.log 2>&1; tail -15 gpurun_out/r2_batch_small.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -s -k "not 40k" > gpurun_out/r2_fullsize.log 2>&1; tail -30 gpurun_out/r2_fullsize.log
timeout 300 python bench.py --workload mc --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2_mc_new.json 2> gpurun_out/r2_mc_new.err; cat gpurun_out/r2_mc_new.json; tail -3 gpurun_out/r2_mc_new.err
