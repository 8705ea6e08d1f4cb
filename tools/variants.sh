#!/bin/bash
# Builds libnaiveb200 variants with different -D tuning macros into naivedynamics.jl_b200/variants/<name>.so
# usage: tools/variants.sh name "-DNB200_X=.. -DNB200_Y=.." [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/../naivedynamics.jl_b200/csrc"
mkdir -p ../variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  d=$(mktemp -d)
  for f in api atoms radix_sort lbvh_build traverse forces peer_exchange setup; do
    nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC $flags -Xptxas -v -c $f.cu -o $d/$f.o 2> $d/$f.log &
  done
  wait
  nvcc $ARCH -shared -o ../variants/$name.so $d/*.o -lcudart
  echo "$name: $(grep -A2 'traverse_kernelILb1' $d/traverse.log | grep -o 'Used [0-9]* registers' | head -1) $(grep -A2 'traverse_kernelILb1' $d/traverse.log | grep -o '[0-9]* bytes spill stores' | head -1)"
  rm -rf $d
done
