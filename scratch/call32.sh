#!/bin/bash
# 2 GPUs: sharded bench of the final build (NCCL all-reduce in the library), reference arm under torchrun, C example of the ranks
mkdir -p gpurun_out
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra > gpurun_out/c32_bench_n2.json 2> gpurun_out/c32_bench_n2.err
echo "n2 rc=$?"; cut -c1-300 gpurun_out/c32_bench_n2.json; tail -2 gpurun_out/c32_bench_n2.err
timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/c32_bench_n1.json 2>/dev/null; cut -c1-200 gpurun_out/c32_bench_n1.json
timeout -s KILL 200 python -m pytest tests/test_comm.py -m gpu -q 2>&1 | tail -2
