#!/bin/bash
timeout 200 python scripts/probe_batched.py 10000 16 2>&1 | grep -v "m=12\|m=24\|m=48"
