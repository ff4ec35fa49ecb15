#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_node_dropin.py tests/test_node_reference.py tests/test_gpu_dropin_ref_headers.py -q -x ) > gpurun_out/r2_node_tests.log 2>&1; tail -25 gpurun_out/r2_node_tests.log
