#!/bin/bash
set -x
mkdir -p gpurun_out
EKF_LIB=slam_ros_b200/libekfcuda_mctiming.so timeout 120 python scripts/mc_timing.py 4096 > gpurun_out/r2_mc_timing_4096.log 2>&1; cat gpurun_out/r2_mc_timing_4096.log
EKF_LIB=slam_ros_b200/libekfcuda_mctiming.so timeout 120 python scripts/mc_timing.py 148 > gpurun_out/r2_mc_timing_148.log 2>&1; cat gpurun_out/r2_mc_timing_148.log
