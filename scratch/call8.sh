#!/bin/bash
# GPU call 8: round time vs warps per CTA (what limits the scaling with occupancy?)
mkdir -p gpurun_out
L=$PWD/pumas_b200
echo "== default"; timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
echo "== gmem16"; PUMAS_B200_LIBRARY=$L/libpumas_b200_gmem16.so timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
for w in 4 6 8 10 12 14; do
  echo "== warps $w"
  PUMAS_B200_STATS=1 PUMAS_B200_LIBRARY=$L/libpumas_b200_w$w.so timeout 300 python tests/_gpu_prof.py headline 1000000 1 0 2>&1 | tail -2
done
echo "== ncu w6"
PUMAS_B200_LIBRARY=$L/libpumas_b200_w6.so timeout 900 ncu --set full --import-source on --clock-control none -k regex:transport_hot -c 1 -o gpurun_out/c8_prof_w6 python tests/_gpu_prof.py headline 300000 1 0 > gpurun_out/c8_prof.log 2>&1
ncu -i gpurun_out/c8_prof_w6.ncu-rep --page raw --csv > gpurun_out/c8_prof_w6_raw.csv 2>/dev/null
echo "== ncu w14"
PUMAS_B200_LIBRARY=$L/libpumas_b200_w14.so timeout 900 ncu --set full --import-source on --clock-control none -k regex:transport_hot -c 1 -o gpurun_out/c8_prof_w14 python tests/_gpu_prof.py headline 600000 1 0 > gpurun_out/c8_prof.log 2>&1
ncu -i gpurun_out/c8_prof_w14.ncu-rep --page raw --csv > gpurun_out/c8_prof_w14_raw.csv 2>/dev/null
ls -la gpurun_out/c8*
