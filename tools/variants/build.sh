#!/bin/bash
# Builds variant_bench and one cubin per leaf-hash variant into tools/variants/out/ (git-ignored).
set -e
cd "$(dirname "$0")"
mkdir -p out
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o out/variant_bench variant_bench.cu -lcuda
v() { name=$1; shift; $NV -cubin -o out/$name.cubin leaf_kernel.cu "$@" & }
v base
v redv1 -DGL_REDUCE_V1
v redv1_i2f -DGL_REDUCE_V1 -DGL_MDS_I2F
v i2f -DGL_MDS_I2F
v sqr3 -DGL_SQR_3WIDE
v sqr3_redv1 -DGL_SQR_3WIDE -DGL_REDUCE_V1
v rcconv -DGL_MDS_RC_CONVERT
v rcconv_redv1 -DGL_MDS_RC_CONVERT -DGL_REDUCE_V1
v mdsint -DGL_MDS_INT
v nosync -DVB_SYNC=0
v nosync_redv1 -DVB_SYNC=0 -DGL_REDUCE_V1
v pu2 -DGL_PARTIAL_UNROLL2
v pu2_redv1 -DGL_PARTIAL_UNROLL2 -DGL_REDUCE_V1
v t128b4 -DVB_MINB=4
v t128b4_redv1 -DVB_MINB=4 -DGL_REDUCE_V1
v t128b6_redv1 -DVB_MINB=6 -DGL_REDUCE_V1
v t128b3_redv1 -DVB_MINB=3 -DGL_REDUCE_V1
v t256b2_redv1 -DVB_THREADS=256 -DVB_MINB=2 -DGL_REDUCE_V1
v t256b3_redv1 -DVB_THREADS=256 -DVB_MINB=3 -DGL_REDUCE_V1
v t64b10_redv1 -DVB_THREADS=64 -DVB_MINB=10 -DGL_REDUCE_V1
v t64b8_redv1 -DVB_THREADS=64 -DVB_MINB=8 -DGL_REDUCE_V1
v t96b6_redv1 -DVB_THREADS=96 -DVB_MINB=6 -DGL_REDUCE_V1
v t128b4_nosync_redv1 -DVB_MINB=4 -DVB_SYNC=0 -DGL_REDUCE_V1
wait
for x in "$@"; do :; done
ls out/*.cubin | wc -l
