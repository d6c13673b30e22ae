#!/bin/bash
# scripts/build_variant.sh <name> <extra nvcc flags...>: libb200z with inflate_kernels.cu compiled under other macros, as
# archive_b200/variants/libb200z_<name>.so (B200Z_LIB=... selects it): several builds measured in one GPU call.
set -e
name=$1; shift
cd "$(dirname "$0")/../archive_b200/csrc"
make -s -j8
mkdir -p ../variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function "$@" -c inflate_kernels.cu -o /tmp/inflate_kernels_$name.o
objs=$(ls *.o | grep -v inflate_kernels.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libb200z_$name.so $objs /tmp/inflate_kernels_$name.o
echo built ../variants/libb200z_$name.so
