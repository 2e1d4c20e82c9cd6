#!/bin/bash
# GPU call 15: CSDA tile kernel (cp.async.bulk) vs one-thread-per-track
mkdir -p gpurun_out
echo "== A/B kernel 0 (tile) vs 1 (monolithic general code), bit for bit"
timeout 600 python tests/_gpu_kernels.py -k 0,1 csda_water:8388608 csda_limits_fwd:4000000 csda_limits_bwd:4000000 csda_spectrum_fwd:1000001 2>&1 | grep -v "^   first" | tail -12
echo "== parity"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "csda" 2>&1 | tail -3
echo "== ncu"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:csda_tile -c 1 -o gpurun_out/c15_csda python tests/_gpu_prof.py csda_water 8388608 2 0 > gpurun_out/c15_prof.log 2>&1
tail -2 gpurun_out/c15_prof.log
ncu -i gpurun_out/c15_csda.ncu-rep --page raw --csv > gpurun_out/c15_csda_raw.csv 2>/dev/null
python profiles/tools/ncu_summary.py gpurun_out/c15_csda_raw.csv | head -40
