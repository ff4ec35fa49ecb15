#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "chunked_scan_through" ) > gpurun_out/r2_chunk_edge.log 2>&1; tail -25 gpurun_out/r2_chunk_edge.log
