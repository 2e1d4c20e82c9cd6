#!/bin/bash
# GPU call 29: full GPU test suite + default bench (15-word slots build)
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c29_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c29_tests.log
timeout -s KILL 600 python bench.py > gpurun_out/c29_bench.json 2> gpurun_out/c29_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/c29_bench.json
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
