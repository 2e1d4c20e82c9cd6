#!/bin/bash
# GPU call 16 (2 GPUs): multi-rank paths
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_comm.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/c16_bench_n2.json 2> gpurun_out/c16_bench_n2.err
tail -c 1200 gpurun_out/c16_bench_n2.json; tail -5 gpurun_out/c16_bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 4 --warmup 3 --no-extra > gpurun_out/c16_bench_n1.json 2>/dev/null
python - <<'PY'
import json
a=json.load(open('gpurun_out/c16_bench_n1.json')); b=json.load(open('gpurun_out/c16_bench_n2.json'))
print("N=1", a['value'], a['e2e']['value'], "N=2", b['value'], b['e2e']['value'], "efficiency", b['value']/(2*a['value']))
PY
