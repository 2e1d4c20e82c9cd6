#!/bin/bash
# GPU call 9: tunables of the hot kernel (service CTAs, patience, EHS period)
mkdir -p gpurun_out
run() { timeout 300 python tests/_gpu_prof.py headline 2000000 2 0 2>&1 | tail -1 | sed 's/.*-> //'; }
echo "default: $(run)"
for d in 10 12 14 16 18; do for pt in 3 10 30; do
  echo "service_ctas=$d patience=$pt: $(PUMAS_B200_DEL_CTAS=$d PUMAS_B200_SERVICE_PATIENCE=$pt run)"
done; done
for pe in 3 5 6 8; do
  echo "ehs_period=$pe: $(PUMAS_B200_EHS_PERIOD=$pe run)"
done
for th in 8 16 32; do
  echo "del_threshold=$th ctas=14 patience=30: $(PUMAS_B200_DEL_THRESHOLD=$th PUMAS_B200_DEL_CTAS=14 PUMAS_B200_SERVICE_PATIENCE=30 run)"
done
