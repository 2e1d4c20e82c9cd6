#!/bin/bash
# GPU call 25: ncu --set full of the hot kernel (fast math build), headline 600k tracks
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:transport_hot -c 1 -o gpurun_out/c25_hot python tests/_gpu_prof.py headline 600000 1 0 > gpurun_out/c25_prof.log 2>&1
tail -2 gpurun_out/c25_prof.log
ncu -i gpurun_out/c25_hot.ncu-rep --page raw --csv > gpurun_out/c25_hot_raw.csv 2>/dev/null
ls -la gpurun_out/c25*
