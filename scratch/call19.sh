#!/bin/bash
# GPU call 19 (8 GPUs): scaling check of bench.py
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/c19_bench_n$n.json 2> gpurun_out/c19_bench_n$n.err
echo "N=$n: $(grep -o '"value": [0-9.]*' gpurun_out/c19_bench_n$n.json | head -2 | tr '\n' ' ') lines=$(wc -l < gpurun_out/c19_bench_n$n.json)"
done
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-extra > gpurun_out/c19_bench_n1.json 2>/dev/null
echo "N=1: $(grep -o '"value": [0-9.]*' gpurun_out/c19_bench_n1.json | head -2 | tr '\n' ' ')"
timeout 300 python -m pytest tests/test_comm.py -m gpu -q 2>&1 | tail -2
