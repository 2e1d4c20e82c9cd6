#!/bin/bash
timeout -s KILL 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout -s KILL 60 python tests/_gpu_variant.py headline 2000000 2 2>&1 | tail -1
