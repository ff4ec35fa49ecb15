#!/bin/bash
timeout 120 python scripts/probe_lines.py 2>&1 | tail -6
timeout 300 python -m pytest tests/test_gpu_lines.py -x -q 2>&1 | tail -15
