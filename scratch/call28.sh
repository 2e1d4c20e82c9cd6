#!/bin/bash
# slots of 15 words (5 event-only words in the HBM mirror): A/B
run() { PUMAS_B200_LIBRARY=$PWD/pumas_b200/$1 timeout -s KILL 100 python tests/_gpu_variant.py $2 1000000 2 2>&1 | tail -1; }
for w in headline straight straight_mixed rock_3km; do run libpumas_b200_before.so $w; run libpumas_b200.so $w; done
