#!/bin/bash
# build_variant.sh <name> <file.cu> [-D...]: build/pixie_cuda_<name>.so with one translation unit recompiled with extra flags
name=$1; src=$2; shift 2
base=$(basename $src .cu)
mkdir -p build/var
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC "$@" -I include -c -o build/var/${base}_$name.o $src || exit 1
objs=$(ls build/*.o | grep -v "/$base.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/pixie_cuda_$name.so $objs build/var/${base}_$name.o
