#!/bin/bash
run() { PUMAS_B200_LIBRARY=$PWD/pumas_b200/$1 PUMAS_B200_DEL_CTAS=$3 timeout -s KILL 120 python tests/_gpu_variant.py $2 1000000 2 2>&1 | tail -1 | sed "s/^/svc=$3 /"; }
run libpumas_b200_before.so headline 0
for s in 0 6 8 10 12 14; do run libpumas_b200.so headline $s; done
run libpumas_b200_before.so rock_3km 0
for s in 0 8 12 16; do run libpumas_b200.so rock_3km $s; done
