#!/bin/bash
timeout 900 python -m pytest tests/test_recorder.py tests/test_abi_symbols.py -m gpu -x -q 2>&1 | tail -12
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
