#!/bin/bash
# full GPU test-suite (the sharded tests need the second GPU), smoke, the default bench line and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/validate_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/validate_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/validate_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/validate_ref.json 2> gpurun_out/validate_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/validate_ref.json
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_quad -c 6 -o gpurun_out/prof_sweep_quad_final -f python scripts/ncu_sweep.py 10000 8,16,32 > gpurun_out/ncu_quad_final.log 2>&1; echo "ncu sweep rc=$?"
