#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/probe_batched.py 10000 8,16,32 > gpurun_out/probe_quad2.log 2>&1; echo "rc=$?"; cat gpurun_out/probe_quad2.log
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quad2_10k.json 2> gpurun_out/bench_quad2_10k.err; echo "rc=$?"; cat gpurun_out/bench_quad2_10k.json
