#!/bin/bash
# GPU call 1 (round 2): baseline vs FMA-contraction build: parity on all workloads + quick bench; FP64 micro-benchmarks
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt
./scratch/ubench > gpurun_out/c1_ubench.txt 2>&1
python profiles/tools/parity_report.py gpurun_out/c1_parity_base.json 1 > gpurun_out/c1_parity_base.log 2>&1
PUMAS_B200_LIBRARY=$PWD/pumas_b200/libpumas_b200_fmad.so python profiles/tools/parity_report.py gpurun_out/c1_parity_fmad.json 1 > gpurun_out/c1_parity_fmad.log 2>&1
python bench.py --steps 3 --warmup 3 --tracks 4194304 > gpurun_out/c1_bench_base.json 2> gpurun_out/c1_bench_base.err
PUMAS_B200_LIBRARY=$PWD/pumas_b200/libpumas_b200_fmad.so python bench.py --steps 3 --warmup 3 --tracks 4194304 > gpurun_out/c1_bench_fmad.json 2> gpurun_out/c1_bench_fmad.err
tail -3 gpurun_out/c1_parity_base.log gpurun_out/c1_parity_fmad.log
cat gpurun_out/c1_ubench.txt
