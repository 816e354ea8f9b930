#!/bin/bash
# Build an experimental variant of the library: tools/build_variant.sh <tag> <extra nvcc flags...>  ->  streamz_b200/lib_exp/libstreamz_b200_<tag>.so
# (only mlp.cu is recompiled with the extra flags; the other objects come from the regular build)
set -e
tag=$1; shift
cd "$(dirname "$0")/../streamz_b200/csrc"
make -j8 > /dev/null
mkdir -p _build/var ../lib_exp
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas "$@" -c mlp.cu -o _build/var/mlp_$tag.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib_exp/libstreamz_b200_$tag.so _build/capi.o _build/frontend.o _build/var/mlp_$tag.o _build/comm.o _build/formats.o _build/loops.o _build/train_small.o -ldl
echo built lib_exp/libstreamz_b200_$tag.so
