#!/bin/bash
# usage: build_variant.sh <name> [extra nvcc flags...]   -> scratch/variants/lib_<name>.so
name=$1; shift
cd /root/repo
PUMAS_B200_TARGET=/root/repo/scratch/variants/lib_$name.so PUMAS_B200_NVCC_EXTRA="$*" python -m pumas_b200.build --force > /tmp/build_$name.log 2>&1 && echo "built $name" || (echo "FAILED $name"; tail -5 /tmp/build_$name.log)
