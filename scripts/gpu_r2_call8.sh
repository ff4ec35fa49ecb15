#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q --durations=12 ) > gpurun_out/r2_pytest_full.log 2>&1; tail -45 gpurun_out/r2_pytest_full.log
