#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_lines.py tests/test_gpu_cpp_dropin.py tests/test_gpu_node.py -m gpu -q ) > gpurun_out/r2_lx_tests.log 2>&1; tail -8 gpurun_out/r2_lx_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
