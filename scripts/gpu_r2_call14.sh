#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/steps_only.py 1000 6 > gpurun_out/r2_steps_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_lines2 -s 3 -c 2 -f -o gpurun_out/prof_scan_lines2_1k python scripts/steps_only.py 1000 6 > gpurun_out/r2_ncu_lines2.log 2>&1
tail -3 gpurun_out/r2_ncu_lines2.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "predict or line_loop_forms or map_scenario" > gpurun_out/r2_predict_tests.log 2>&1; tail -3 gpurun_out/r2_predict_tests.log
