#!/bin/bash
run() { PUMAS_B200_LIBRARY=$PWD/pumas_b200/$1 timeout -s KILL 60 python tests/_gpu_variant.py $2 1000000 2 2>&1 | tail -1; }
run libpumas_b200_before.so geometry_fast; run libpumas_b200.so geometry_fast
run libpumas_b200_before.so mixed_loss_bwd; run libpumas_b200.so mixed_loss_bwd
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "geometry_fast or mixed_loss_bwd or limits_all_mixed_bwd" 2>&1 | tail -2
