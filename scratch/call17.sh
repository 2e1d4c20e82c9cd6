#!/bin/bash
# GPU call 17: final captures for profiles/ (hot kernel ncu, launch list of bench.py, bench line)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err; head -c 400 gpurun_out/c17_bench.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c17_launches.csv python bench.py --steps 2 --warmup 3 --tracks 2097152 --no-extra --cpu-sample 2000 > gpurun_out/c17_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:transport_hot -c 1 -o gpurun_out/c17_hot python tests/_gpu_prof.py headline 600000 1 0 > gpurun_out/c17_prof.log 2>&1
ncu -i gpurun_out/c17_hot.ncu-rep --page raw --csv > gpurun_out/c17_hot_raw.csv 2>/dev/null
ncu -i gpurun_out/c17_hot.ncu-rep --page source --csv > gpurun_out/c17_hot_source.csv 2>/dev/null
gzip -f gpurun_out/c17_hot_source.csv
python profiles/tools/ncu_summary.py gpurun_out/c17_hot_raw.csv | head -45
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
