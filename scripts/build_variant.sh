#!/bin/bash
# usage: build_variant.sh <suffix> <extra nvcc flags...>   -> mcmc-symreg_b200/libbsr_b200_<suffix>.so (A/B experiments)
set -e
cd "$(dirname "$0")/.."
suf=$1; shift
mkdir -p build/obj_$suf
for f in mcmc-symreg_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c -o build/obj_$suf/$b.o $f &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o mcmc-symreg_b200/libbsr_b200_$suf.so build/obj_$suf/*.o
echo built mcmc-symreg_b200/libbsr_b200_$suf.so
