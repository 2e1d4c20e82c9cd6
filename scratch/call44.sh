#!/bin/bash
run() { PUMAS_B200_LIBRARY=$PWD/pumas_b200/$1 timeout -s KILL 40 python tests/_gpu_variant.py $2 $3 2 2>&1 | tail -1; }
run libpumas_b200_before.so headline 2000000
run libpumas_b200_mathinl.so headline 2000000
run libpumas_b200_before.so rock_3km 1000000
run libpumas_b200_mathinl.so rock_3km 1000000
