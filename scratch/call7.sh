#!/bin/bash
# GPU call 7: hot kernel v6 (code diet: out-of-line div/sqrt/log/sincospi/locate/EHS)
mkdir -p gpurun_out
L=$PWD/pumas_b200
for v in "" _gmem16 _w12 _w10 _nolock; do
  echo "== variant [$v]"
  PUMAS_B200_LIBRARY=$L/libpumas_b200$v.so timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
done
for d in 12 16; do
  echo "== service CTAs $d"
  PUMAS_B200_DEL_CTAS=$d timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
done
echo "== stats"
PUMAS_B200_STATS=1 PUMAS_B200_LIBRARY=$L/libpumas_b200_stats.so timeout 300 python tests/_gpu_prof.py headline 2000000 1 0 2>&1 | tail -3
echo "== other workloads, hot vs resident"
timeout 900 python tests/_gpu_kernels.py -k 0,3 straight:1000000 straight_mixed:1000000 rock_3km:200000 2>&1 | grep -v "^   first\|bit for bit" | tail -6
echo "== ncu"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:transport_hot -c 1 -o gpurun_out/c7_prof python tests/_gpu_prof.py headline 600000 1 0 > gpurun_out/c7_prof.log 2>&1
tail -2 gpurun_out/c7_prof.log
ncu -i gpurun_out/c7_prof.ncu-rep --page raw --csv > gpurun_out/c7_prof_raw.csv 2>/dev/null
