#!/bin/bash
# round-end check on ONE GPU (what the driver runs): pytest -m gpu, smoke(), the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/validate1_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/validate1_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/validate1_bench.json 2> gpurun_out/validate1_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/validate1_bench.json
