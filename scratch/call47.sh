#!/bin/bash
run() { PUMAS_B200_LIBRARY=$PWD/pumas_b200/$1 timeout -s KILL 40 python tests/_gpu_variant.py headline 2000000 2 2>&1 | tail -1; }
run libpumas_b200.so; run libpumas_b200_pf.so
