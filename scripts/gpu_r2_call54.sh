#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_lines.py -m gpu -q -x -k "world_segments" ) > gpurun_out/r2_world_seg.log 2>&1; tail -25 gpurun_out/r2_world_seg.log
