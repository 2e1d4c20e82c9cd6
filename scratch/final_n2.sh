# Round-end check on 2 GPUs: the sharded bench (NCCL all-reduce of the tallies) and
# the single-GPU bench, outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/final_gpus.txt 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/final_bench_n2.json 2> gpurun_out/final_bench_n2.err
echo "n2 rc=$?"; tail -c 600 gpurun_out/final_bench_n2.json; tail -5 gpurun_out/final_bench_n2.err
timeout 200 python bench.py --steps 3 --warmup 3 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
echo "n1 rc=$?"; tail -c 600 gpurun_out/final_bench_n1.json; tail -5 gpurun_out/final_bench_n1.err
