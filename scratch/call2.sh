#!/bin/bash
# GPU call 2: first run of the shared-memory resident kernel (hot_kernel.cuh)
mkdir -p gpurun_out
L=$PWD/pumas_b200
echo "== A/B default (14+4): hot (0) vs register resident (3) vs monolithic (1), bit for bit" 
timeout 600 python tests/_gpu_kernels.py -k 0,3,1 headline:100000 straight:100000 straight_mixed:60000 rock_3km:20000 limits_all:40000 limits_all_bwd:40000 > gpurun_out/c2_ab.log 2>&1
tail -30 gpurun_out/c2_ab.log
echo "== larger batch, hot vs resident"
timeout 600 python tests/_gpu_kernels.py -k 0,3 headline:2000000 >> gpurun_out/c2_ab.log 2>&1
tail -3 gpurun_out/c2_ab.log
for v in w14s2 w12s4 w16s4 w10s6; do
  echo "== variant $v"
  PUMAS_B200_LIBRARY=$L/libpumas_b200_$v.so timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1
done
echo "== stats"
PUMAS_B200_STATS=1 PUMAS_B200_LIBRARY=$L/libpumas_b200_w14s4stats.so timeout 300 python tests/_gpu_prof.py headline 2000000 1 0 2>&1 | tail -3
echo "== parity vs reference"
timeout 900 python profiles/tools/parity_report.py gpurun_out/c2_parity.json 1 headline,rock_3km,straight,straight_mixed,limits_all,limits_all_bwd,vertex_stop,vertex_stop_bwd,above_grid,tau_ice,models_kkb_fwd,models_kkb_bwd,models_abm_fwd,models_abm_bwd 2>&1 | tail -16
echo "== ncu"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:transport_hot -c 1 -o gpurun_out/c2_prof python tests/_gpu_prof.py headline 600000 1 0 > gpurun_out/c2_prof.log 2>&1
tail -2 gpurun_out/c2_prof.log
ncu -i gpurun_out/c2_prof.ncu-rep --page raw --csv > gpurun_out/c2_prof_raw.csv 2>/dev/null
ls -la gpurun_out/c2_prof.ncu-rep
