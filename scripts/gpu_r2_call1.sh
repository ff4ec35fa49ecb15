#!/bin/bash
# round 2, first GPU call: baseline of the tests, the fp64 tensor-core probe, and the first ncu capture of k_batch_scan
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
./scripts/dmma_probe.bin > gpurun_out/dmma_probe.log 2>&1; tail -12 gpurun_out/dmma_probe.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_base.log 2>&1; tail -3 gpurun_out/r2_pytest_base.log
timeout 300 python bench.py --workload mc --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_mc_base.json 2> gpurun_out/r2_mc_base.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batch_scan -s 3 -c 2 -f -o gpurun_out/prof_batch_r2_base \
   python bench.py --workload mc --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_mc.log 2>&1
cat gpurun_out/r2_mc_base.json
tail -5 gpurun_out/r2_ncu_mc.log
