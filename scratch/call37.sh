#!/bin/bash
# GPU call 37: final validation of the round 2 build
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > gpurun_out/c37_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c37_tests.log
timeout -s KILL 400 python bench.py > gpurun_out/c37_bench.json 2> gpurun_out/c37_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/c37_bench.json
timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
