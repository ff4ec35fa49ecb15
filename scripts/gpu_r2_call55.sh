#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_cpp_dropin.py tests/test_gpu_lines.py tests/test_abi.py -q ) > gpurun_out/r2_f4_tests.log 2>&1; tail -8 gpurun_out/r2_f4_tests.log
