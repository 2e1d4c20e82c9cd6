#!/bin/bash
run() { PUMAS_B200_LIBRARY=$PWD/pumas_b200/$1 timeout -s KILL 40 python tests/_gpu_variant.py $2 $3 2 2>&1 | tail -1; }
for v in vs vse vsp vsep; do run libpumas_b200_$v.so headline 2000000; done
for v in vs vsep; do run libpumas_b200_$v.so straight_mixed 1000000; done
