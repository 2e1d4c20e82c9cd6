#!/bin/bash
# GPU call 46: last validation of round 2 (inline exp build)
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > gpurun_out/c46_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c46_tests.log
timeout -s KILL 400 python bench.py > gpurun_out/c46_bench.json 2> gpurun_out/c46_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/c46_bench.json
