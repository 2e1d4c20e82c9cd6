#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/_gpu_mathtest.py 1e9
for w in headline straight straight_mixed rock_3km; do
PUMAS_B200_LIBRARY=$PWD/pumas_b200/libpumas_b200_mathcheck.so timeout 300 python tests/_gpu_variant.py $w 400000 1 2>&1 | tail -6
done
for w in headline straight straight_mixed rock_3km; do
for v in libpumas_b200_ieee.so libpumas_b200.so; do
PUMAS_B200_LIBRARY=$PWD/pumas_b200/$v timeout 300 python tests/_gpu_variant.py $w 1000000 2 2>&1 | tail -1
done; done
