#!/bin/bash
# full GPU test-suite (the sharded tests need the second GPU), smoke, and the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/validate_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/validate_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/validate_bench.json
