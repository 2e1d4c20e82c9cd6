#!/bin/bash
run() { PUMAS_B200_LIBRARY=$PWD/pumas_b200/$1 PUMAS_B200_DEL_CTAS=$4 timeout -s KILL 200 python tests/_gpu_variant.py $2 $3 2 2>&1 | tail -1 | sed "s/^/del=$4 /"; }
run libpumas_b200.so geometry_detailed 1000000 0
run libpumas_b200_rs16.so geometry_detailed 1000000 0
for d in 6 12 16; do run libpumas_b200.so geometry_detailed 1000000 $d; done
