#!/bin/bash
# GPU call 39: profiles of the final build: ncu --set full of the hot kernel (headline 600k), launch list of the bench
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:transport_hot -c 1 -o gpurun_out/c39_hot python tests/_gpu_prof.py headline 600000 1 0 > gpurun_out/c39_prof.log 2>&1
tail -1 gpurun_out/c39_prof.log
ncu -i gpurun_out/c39_hot.ncu-rep --page raw --csv > gpurun_out/c39_hot_raw.csv 2>/dev/null
ncu -i gpurun_out/c39_hot.ncu-rep --page source --csv > gpurun_out/c39_hot_source.csv 2>/dev/null
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c39_launches.csv python bench.py --steps 2 --warmup 1 --tracks 2097152 > gpurun_out/c39_bench_under_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/c39*
