#!/bin/bash
# GPU call 42: final validation of the build with the loop's calls in line
mkdir -p gpurun_out
for w in straight straight_mixed rock_3km; do timeout -s KILL 60 python tests/_gpu_variant.py $w 1000000 2 2>&1 | tail -1; done
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > gpurun_out/c42_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c42_tests.log
timeout -s KILL 400 python bench.py > gpurun_out/c42_bench.json 2> gpurun_out/c42_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/c42_bench.json
timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
