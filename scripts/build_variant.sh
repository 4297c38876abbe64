#!/bin/bash
# usage: build_variant.sh <suffix> <extra nvcc flags...>   -> mcmc-symreg_b200/libbsr_b200_<suffix>.so (A/B experiments; select it with BSR_LIB=...)
# Only the window translation unit is recompiled with the extra flags; the other objects come from build/obj (run build() first).
set -e
cd "$(dirname "$0")/.."
suf=$1; shift
mkdir -p build/obj_$suf
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c -o build/obj_$suf/bsr_tu_window.o mcmc-symreg_b200/csrc/bsr_tu_window.cu
objs=$(ls build/obj/*.o | grep -v bsr_tu_window.o | grep -v bsr_tu_eval_f32.o)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o mcmc-symreg_b200/libbsr_b200_$suf.so build/obj_$suf/bsr_tu_window.o $objs
echo built mcmc-symreg_b200/libbsr_b200_$suf.so
