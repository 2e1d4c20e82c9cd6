#!/bin/bash
run() { PUMAS_B200_LIBRARY=$PWD/pumas_b200/$1 timeout -s KILL 100 python tests/_gpu_variant.py $2 1000000 2 2>&1 | tail -1; }
for v in libpumas_b200.so libpumas_b200_ls2.so libpumas_b200_ls0.so; do run $v headline; done
