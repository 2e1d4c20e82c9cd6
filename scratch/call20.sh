#!/bin/bash
# GPU call 20: device callbacks + user DCS (tests), callback kernel timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_callbacks.py -m gpu -q -s > gpurun_out/c20_tests.log 2>&1
grep -n "Error\|^E  \|passed\|failed\|{" gpurun_out/c20_tests.log | cut -c1-600 | tail -40
