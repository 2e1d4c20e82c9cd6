#!/bin/bash
# GPU call 10: full GPU test-suite + bench + parity report with the hot kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/c10_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; tail -c 1500 gpurun_out/c10_bench.json
timeout 900 python profiles/tools/parity_report.py gpurun_out/c10_parity.json 1 > gpurun_out/c10_parity.log 2>&1; tail -31 gpurun_out/c10_parity.log
