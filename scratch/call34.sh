#!/bin/bash
for k in 0 3 2; do timeout -s KILL 200 python tests/_gpu_prof.py tau_ice 4000000 2 $k 2>&1 | tail -1; done
for s in 24 36 48 64; do PUMAS_B200_DEL_CTAS=$s timeout -s KILL 200 python tests/_gpu_prof.py tau_ice 4000000 2 0 2>&1 | tail -1 | sed "s/^/svc=$s /"; done
