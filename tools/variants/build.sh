#!/bin/bash
# Builds variant_bench and one cubin per leaf-hash variant into tools/variants/out/ (git-ignored).
set -e
cd "$(dirname "$0")"
mkdir -p out
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o out/variant_bench variant_bench.cu -lcuda
v() { name=$1; shift; $NV -cubin -o out/$name.cubin leaf_kernel.cu "$@" & }
v base
v sboxint -DGL_SBOX_INT
v mulx -DGL_MUL_EXPLICIT
v rint -DGL_RENORM_INT
v f2i -DGL_CVT_F2I
v il -DGL_SBOX_INTLIMBS
v sqr3 -DGL_SQR_3WIDE
v mulx_sqr3 -DGL_MUL_EXPLICIT -DGL_SQR_3WIDE
v mulx_rint -DGL_MUL_EXPLICIT -DGL_RENORM_INT
v f2i_il -DGL_CVT_F2I -DGL_SBOX_INTLIMBS
v f2i_rint -DGL_CVT_F2I -DGL_RENORM_INT
v mulx_f2i_il -DGL_MUL_EXPLICIT -DGL_CVT_F2I -DGL_SBOX_INTLIMBS
v all1 -DGL_CVT_F2I -DGL_SBOX_INTLIMBS -DGL_MUL_EXPLICIT -DGL_RENORM_INT
v all2 -DGL_CVT_F2I -DGL_SBOX_INTLIMBS -DGL_MUL_EXPLICIT -DGL_RENORM_INT -DGL_SQR_3WIDE
v all3 -DGL_CVT_F2I -DGL_MUL_EXPLICIT -DGL_RENORM_INT -DGL_SBOX_INT
v all1_t128b4 -DGL_CVT_F2I -DGL_SBOX_INTLIMBS -DGL_MUL_EXPLICIT -DGL_RENORM_INT -DVB_MINB=4
v all2_t128b4 -DGL_CVT_F2I -DGL_SBOX_INTLIMBS -DGL_MUL_EXPLICIT -DGL_RENORM_INT -DGL_SQR_3WIDE -DVB_MINB=4
v all2_redv1 -DGL_CVT_F2I -DGL_SBOX_INTLIMBS -DGL_MUL_EXPLICIT -DGL_RENORM_INT -DGL_SQR_3WIDE -DGL_REDUCE_V1
v all1_mdslit -DGL_CVT_F2I -DGL_SBOX_INTLIMBS -DGL_MUL_EXPLICIT -DGL_RENORM_INT -DGL_MDS_LITERAL
v f2i_il_rint -DGL_CVT_F2I -DGL_SBOX_INTLIMBS -DGL_RENORM_INT
v cvtmagic -DGL_CVT_MAGIC
v pfast -DGL_PARTIAL_FAST
wait
for x in "$@"; do :; done
ls out/*.cubin | wc -l
