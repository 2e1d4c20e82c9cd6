#!/bin/bash
# GPU call 5: hot kernel v4 (fast log / sincospi, unchecked draws, 32 bit slot indices, polling every period)
mkdir -p gpurun_out
L=$PWD/pumas_b200
echo "== A/B: hot (0) vs register resident (3)"
timeout 900 python tests/_gpu_kernels.py -k 0,3 headline:2000000 straight_mixed:1000000 2>&1 | grep -v "^   first" | tail -6
for d in 12 16; do
  echo "== service CTAs $d"
  PUMAS_B200_DEL_CTAS=$d timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
done
for v in w18 w12; do
echo "== variant $v"
PUMAS_B200_LIBRARY=$L/libpumas_b200_$v.so timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
done
echo "== stats"
PUMAS_B200_STATS=1 PUMAS_B200_LIBRARY=$L/libpumas_b200_stats.so timeout 300 python tests/_gpu_prof.py headline 2000000 1 0 2>&1 | tail -3
echo "== parity vs reference"
timeout 900 python profiles/tools/parity_report.py gpurun_out/c5_parity.json 1 headline,rock_3km,straight,straight_mixed,limits_all,limits_all_bwd,vertex_stop,vertex_stop_bwd 2>&1 | tail -8
echo "== ncu"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:transport_hot -c 1 -o gpurun_out/c5_prof python tests/_gpu_prof.py headline 600000 1 0 > gpurun_out/c5_prof.log 2>&1
tail -2 gpurun_out/c5_prof.log
ncu -i gpurun_out/c5_prof.ncu-rep --page raw --csv > gpurun_out/c5_prof_raw.csv 2>/dev/null
