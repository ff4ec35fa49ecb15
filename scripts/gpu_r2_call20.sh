#!/bin/bash
mkdir -p gpurun_out
EKF_LIB=slam_ros_b200/libekfcuda_mctiming.so timeout 200 python scripts/mc_timing.py 148 > gpurun_out/r2_mc_timing2_148.log 2>&1; tail -8 gpurun_out/r2_mc_timing2_148.log
EKF_LIB=slam_ros_b200/libekfcuda_mctiming.so timeout 200 python scripts/mc_timing.py 4096 > gpurun_out/r2_mc_timing2_4096.log 2>&1; tail -4 gpurun_out/r2_mc_timing2_4096.log
