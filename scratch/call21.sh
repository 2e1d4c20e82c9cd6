#!/bin/bash
# GPU call 21: warps x slots x EHS period sweep of the hot kernel (headline, 1M tracks)
mkdir -p gpurun_out
run() { PUMAS_B200_LIBRARY=$PWD/pumas_b200/$1 PUMAS_B200_EHS_PERIOD=$2 PUMAS_B200_DEL_CTAS=$3 timeout 120 python tests/_gpu_variant.py headline 1000000 2 2>&1 | tail -1 | sed "s/^/period=$2 svc=$3 /"; }
run libpumas_b200.so 4 18
for v in w18s2; do for p in 4; do run libpumas_b200_$v.so $p 18; done; done
for v in w24s1 w28s1 w32s1; do for p in 1 2 3; do run libpumas_b200_$v.so $p 18; done; done
for v in w24s1 w32s1; do run libpumas_b200_$v.so 2 24; done
