#!/bin/bash
for c in 16 32 8; do echo "== MAXC=$c"; EKF_SWEEP_MAXC=$c timeout 120 python scripts/probe_sweep_m.py 2>&1 | tail -2; done
echo "== MAXC=16 batched 33 steps"; EKF_SWEEP_MAXC=16 timeout 120 python scripts/probe_batched.py 10000 64 2>&1 | tail -2
echo "== default"; timeout 200 python scripts/probe_batched.py 10000 8,16,32 2>&1 | grep -v "m=12\|m=24\|m=48"
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
