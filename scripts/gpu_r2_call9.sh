#!/bin/bash
set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/fp64_latency scripts/fp64_latency.cu && /tmp/fp64_latency > gpurun_out/r2_fp64_latency.log 2>&1; cat gpurun_out/r2_fp64_latency.log
timeout 300 python bench.py --workload mc --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_mc_v5.json 2> gpurun_out/r2_mc_v5.err; cut -c1-1800 gpurun_out/r2_mc_v5.json
timeout 300 python bench.py --workload 1k --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_1k.json 2> gpurun_out/r2_bench_1k.err; cut -c1-2500 gpurun_out/r2_bench_1k.json
timeout 300 python bench.py --lines 32 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_10k_m32.json 2> gpurun_out/r2_bench_10k_m32.err; cut -c1-2500 gpurun_out/r2_bench_10k_m32.json
python scripts/sanitize_smoke.py > gpurun_out/r2_sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_smoke.py > gpurun_out/r2_sanitize_memcheck.log 2>&1
tail -15 gpurun_out/r2_sanitize_plain.log; tail -25 gpurun_out/r2_sanitize_memcheck.log
