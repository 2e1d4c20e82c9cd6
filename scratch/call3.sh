#!/bin/bash
# GPU call 3: hot kernel v2 (service CTAs through global mirrors, lockstep, lean EHS)
mkdir -p gpurun_out
L=$PWD/pumas_b200
echo "== A/B: hot (0) vs register resident (3) vs monolithic (1), bit for bit"
timeout 900 python tests/_gpu_kernels.py -k 0,3,1 headline:100000 straight:100000 straight_mixed:60000 rock_3km:20000 limits_all:40000 limits_all_bwd:40000 > gpurun_out/c3_ab.log 2>&1
grep -v "^   first" gpurun_out/c3_ab.log | tail -24
echo "== larger batch"
timeout 600 python tests/_gpu_kernels.py -k 0,3 headline:2000000 2>&1 | tail -3
for v in nolock w12; do
  echo "== variant $v"
  PUMAS_B200_LIBRARY=$L/libpumas_b200_$v.so timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
done
for d in 12 16 28 36; do
  echo "== service CTAs $d"
  PUMAS_B200_DEL_CTAS=$d timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
done
for pe in 3 5 6; do
  echo "== ehs period $pe"
  PUMAS_B200_EHS_PERIOD=$pe timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
done
echo "== stats"
PUMAS_B200_STATS=1 PUMAS_B200_LIBRARY=$L/libpumas_b200_stats.so timeout 300 python tests/_gpu_prof.py headline 2000000 1 0 2>&1 | tail -3
echo "== ncu"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:transport_hot -c 1 -o gpurun_out/c3_prof python tests/_gpu_prof.py headline 600000 1 0 > gpurun_out/c3_prof.log 2>&1
tail -2 gpurun_out/c3_prof.log
ncu -i gpurun_out/c3_prof.ncu-rep --page raw --csv > gpurun_out/c3_prof_raw.csv 2>/dev/null
ls -la gpurun_out/c3_prof.ncu-rep
