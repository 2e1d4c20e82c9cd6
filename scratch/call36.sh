#!/bin/bash
run() { PUMAS_B200_LIBRARY=$PWD/pumas_b200/$1 PUMAS_B200_DEL_CTAS=$4 PUMAS_B200_DEL_THRESHOLD=$5 timeout -s KILL 40 python tests/_gpu_variant.py $2 $3 2 2>&1 | tail -1 | sed "s/^/svc=$4 thr=$5 /"; }
for s in 8 12 24 32; do run libpumas_b200.so headline 1000000 $s 0; done
for t in 64 200; do run libpumas_b200.so headline 1000000 18 $t; done
run libpumas_b200.so rock_3km 1000000 0 0
run libpumas_b200_before.so rock_3km 1000000 0 0
