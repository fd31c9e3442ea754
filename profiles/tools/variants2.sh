#!/bin/bash
# usage: scratch/variants2.sh name1 "flags1" name2 "flags2" ...  -> builds profiles/tools/libs/<name>.so (all .cu), prints regs/spills of k_emit*
mkdir -p profiles/tools/libs
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  (cd magot_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags -Xptxas -v -o ../../profiles/tools/libs/$name.so *.cu > ../../profiles/tools/libs/$name.log 2>&1; grep -A2 "Compiling entry function '_Z[0-9]*k_emit" ../../profiles/tools/libs/$name.log | grep -E "registers|spill" | sed "s/^/$name: /") &
done
wait
