#!/bin/bash
timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ragged_scans" 2>&1 | tail -15
