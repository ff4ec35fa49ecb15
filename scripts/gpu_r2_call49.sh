#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "device_resident" ) > gpurun_out/r2_devres_tests2.log 2>&1; tail -25 gpurun_out/r2_devres_tests2.log
