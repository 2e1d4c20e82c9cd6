#!/bin/bash
# GPU call 27: full GPU test suite, default bench (both arms), launch list of the bench under ncu
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c27_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c27_tests.log
timeout -s KILL 600 python bench.py > gpurun_out/c27_bench.json 2> gpurun_out/c27_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/c27_bench.json
timeout -s KILL 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c27_bench_ref.json 2> gpurun_out/c27_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/c27_bench_ref.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c27_launches.csv python bench.py --steps 2 --warmup 1 --tracks 2097152 > gpurun_out/c27_bench_under_ncu.log 2>&1; echo "ncu rc=$?"
