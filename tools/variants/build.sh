#!/bin/bash
# Builds variant_bench and one cubin per leaf-hash variant into tools/variants/out/ (git-ignored).
set -e
cd "$(dirname "$0")"
mkdir -p out
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o out/variant_bench variant_bench.cu -lcuda
v() { name=$1; shift; $NV -cubin -o out/$name.cubin leaf_kernel.cu "$@" & }
v base
v cvtmagic -DGL_CVT_MAGIC
v pfast -DGL_PARTIAL_FAST
v mdsint -DGL_MDS_INT
v mulx -DGL_MUL_EXPLICIT
v sqr3 -DGL_SQR_3WIDE
v redv1 -DGL_REDUCE_V1
v nosync -DVB_SYNC=0
v t128b4 -DVB_MINB=4
v t128b6 -DVB_MINB=6
v t256b2 -DVB_THREADS=256 -DVB_MINB=2
v t64b10 -DVB_THREADS=64 -DVB_MINB=10
wait
for x in "$@"; do :; done
ls out/*.cubin | wc -l
