#!/bin/bash
mkdir -p gpurun_out
for v in libpumas_b200.so libpumas_b200_g2.so libpumas_b200_g4.so; do
for w in headline straight; do
PUMAS_B200_LIBRARY=$PWD/pumas_b200/$v timeout 300 python tests/_gpu_variant.py $w 1000000 2 2>&1 | tail -1
done; done
