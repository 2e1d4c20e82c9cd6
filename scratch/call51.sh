#!/bin/bash
for w in headline_density straight_density csda_density; do timeout -s KILL 60 python tests/_gpu_variant.py $w 1000000 2 2>&1 | tail -1; done
timeout -s KILL 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
