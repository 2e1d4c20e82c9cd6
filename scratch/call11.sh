#!/bin/bash
# GPU call 11: new bench.py (async loop, extras) + comm tests + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_comm.py tests/test_gpu_schedulers.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err; tail -c 3000 gpurun_out/c11_bench.json; tail -3 gpurun_out/c11_bench.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/c11_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/c11_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c11_launches.csv python bench.py --steps 2 --warmup 1 --tracks 2097152 --no-extra --cpu-sample 2000 > gpurun_out/c11_bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/c11_bench_under_ncu.log
