#!/bin/bash
# usage: build_variants.sh name "extra nvcc flags" ...   -> pumas_b200/libpumas_b200_<name>.so (in parallel)
cd /root/repo
while [ $# -gt 1 ]; do
  name=$1; extra=$2; shift 2
  ( PUMAS_B200_TARGET=/root/repo/pumas_b200/libpumas_b200_$name.so PUMAS_B200_NVCC_EXTRA="$extra" python -m pumas_b200.build --force > /tmp/build_$name.log 2>&1; echo "built $name rc=$?" ) &
done
wait
