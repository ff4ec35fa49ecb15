#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lines.py tests/test_gpu_cpp_dropin.py tests/test_abi.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --workload extract --steps 500 --warmup 5 > gpurun_out/bench_extract.json 2> gpurun_out/bench_extract.err; echo "rc=$?"; cat gpurun_out/bench_extract.json | cut -c1-2500
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_extract.csv python bench.py --workload extract --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; grep -c k_lx gpurun_out/launches_extract.csv
