#!/bin/bash
# service CTA count with the prefill
run() { PUMAS_B200_DEL_CTAS=$1 timeout -s KILL 100 python tests/_gpu_variant.py $2 1000000 2 2>&1 | tail -1 | sed "s/^/svc=$1 /"; }
for s in 8 10 12 13; do run $s headline; done
for s in 12 16 20; do run $s straight; run $s straight_mixed; done
for s in 12 14 16 21; do run $s rock_3km; done
