#!/bin/bash
# L1 capacity sensitivity: carveout percent -> 71: 164 KB, 85: 196 KB, 100: 228 KB; "" = driver default
run() { if [ -n "$2" ]; then export PUMAS_B200_CARVEOUT=$2; else unset PUMAS_B200_CARVEOUT; fi; PUMAS_B200_LIBRARY=$PWD/pumas_b200/$1 timeout -s KILL 100 python tests/_gpu_variant.py headline 1000000 2 2>&1 | tail -1 | sed "s/^/carveout=$2 /"; }
run libpumas_b200_w12.so ""
run libpumas_b200_w12.so 71
run libpumas_b200_w12.so 85
run libpumas_b200_w12.so 100
run libpumas_b200.so ""
run libpumas_b200.so 85
run libpumas_b200.so 100
