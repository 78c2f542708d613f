#!/bin/bash
# builds a variant of libibk.so with extra -D flags for ibk_spread.cu / ibk_interp.cu into gpurun-visible scripts/variants/<name>.so
# usage: build_variant.sh <name> <file.cu> <flags...>   (the other objects are reused from ibamr_b200/_obj)
set -e
name=$1; src=$2; shift 2
mkdir -p scripts/variants
o=scripts/variants/${name}_$(basename $src .cu).o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr -ccbin /usr/bin/g++ "$@" -c ibamr_b200/csrc/$src -o $o
objs=$(ls ibamr_b200/_obj/*.o | grep -v "/$(basename $src .cu).o")
nvcc -shared -o scripts/variants/$name.so $objs $o -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -ldl
echo scripts/variants/$name.so
