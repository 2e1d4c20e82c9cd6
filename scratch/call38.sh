#!/bin/bash
run() { PUMAS_B200_EHS_PERIOD=$1 PUMAS_B200_DEL_CTAS=$2 timeout -s KILL 40 python tests/_gpu_variant.py headline 2000000 2 2>&1 | tail -1 | sed "s/^/period=$1 svc=$2 /"; }
for p in 3 4 5 6; do run $p 0; done
for s in 15 16 20; do run 0 $s; done
