#!/bin/bash
run() { PUMAS_B200_LIBRARY=$PWD/pumas_b200/$1 PUMAS_B200_EHS_PERIOD=$2 timeout -s KILL 40 python tests/_gpu_variant.py headline 2000000 2 2>&1 | tail -1 | sed "s/^/period=$2 /"; }
run libpumas_b200_before.so 0
run libpumas_b200.so 0
run libpumas_b200.so 3
run libpumas_b200.so 5
