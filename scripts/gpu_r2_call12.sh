#!/bin/bash
set -x
mkdir -p gpurun_out
for v in 1 0; do
  for N in 1000 10000; do
    EKF_LINE_LOOP=$v EKF_LIB=slam_ros_b200/libekfcuda_timing.so timeout 120 python scripts/line_timing.py $N > gpurun_out/r2_line_timing_ll${v}_$N.log 2>&1; echo "== LINE_LOOP=$v N=$N"; tail -7 gpurun_out/r2_line_timing_ll${v}_$N.log
  done
done
