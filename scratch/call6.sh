#!/bin/bash
# GPU call 6: hot kernel v5 (tables staged to shared memory) + lockstep / warp count experiments
mkdir -p gpurun_out
L=$PWD/pumas_b200
echo "== A/B: hot (0) vs register resident (3)"
timeout 900 python tests/_gpu_kernels.py -k 0,1 headline:100000 straight_mixed:100000 2>&1 | grep -v "^   first" | tail -6
for v in "" _lock2 _lock4 _nolock _gmem16 _w8 _w12; do
  echo "== variant [$v]"
  PUMAS_B200_LIBRARY=$L/libpumas_b200$v.so timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
done
for d in 12 16 26; do
  echo "== service CTAs $d"
  PUMAS_B200_DEL_CTAS=$d timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
done
echo "== stats"
PUMAS_B200_STATS=1 PUMAS_B200_LIBRARY=$L/libpumas_b200_stats.so timeout 300 python tests/_gpu_prof.py headline 2000000 1 0 2>&1 | tail -3
echo "== ncu"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:transport_hot -c 1 -o gpurun_out/c6_prof python tests/_gpu_prof.py headline 600000 1 0 > gpurun_out/c6_prof.log 2>&1
tail -2 gpurun_out/c6_prof.log
ncu -i gpurun_out/c6_prof.ncu-rep --page raw --csv > gpurun_out/c6_prof_raw.csv 2>/dev/null
