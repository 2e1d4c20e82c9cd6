#!/bin/bash
# GPU call 4: hot kernel v3 (lean first grid cell)
mkdir -p gpurun_out
L=$PWD/pumas_b200
echo "== A/B: hot (0) vs register resident (3), bit for bit"
timeout 900 python tests/_gpu_kernels.py -k 0,3 headline:2000000 straight:1000000 straight_mixed:1000000 rock_3km:200000 2>&1 | grep -v "^   first" | tail -12
for d in 8 12 16; do
  echo "== service CTAs $d"
  PUMAS_B200_DEL_CTAS=$d timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
done
for pe in 5 6; do
  echo "== ehs period $pe, 12 service CTAs"
  PUMAS_B200_DEL_CTAS=12 PUMAS_B200_EHS_PERIOD=$pe timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
done
echo "== variant w12"
PUMAS_B200_LIBRARY=$L/libpumas_b200_w12.so timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
echo "== stats"
PUMAS_B200_STATS=1 PUMAS_B200_LIBRARY=$L/libpumas_b200_stats.so timeout 300 python tests/_gpu_prof.py headline 2000000 1 0 2>&1 | tail -3
echo "== ncu"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:transport_hot -c 1 -o gpurun_out/c4_prof python tests/_gpu_prof.py headline 600000 1 0 > gpurun_out/c4_prof.log 2>&1
tail -2 gpurun_out/c4_prof.log
ncu -i gpurun_out/c4_prof.ncu-rep --page raw --csv > gpurun_out/c4_prof_raw.csv 2>/dev/null
