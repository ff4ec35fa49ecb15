#!/bin/bash
mkdir -p gpurun_out
for N in 1000 10000; do
  EKF_LIB=slam_ros_b200/libekfcuda_timing.so timeout 120 python scripts/line_timing.py $N > gpurun_out/r2_line_timing_v3_$N.log 2>&1; echo "== N=$N"; tail -7 gpurun_out/r2_line_timing_v3_$N.log
done
