// Lazy 3-word signed Goldilocks values: v = w0 + w1*2^32 + e*2^64 (e small, signed). SASS-count lab.
#include "../../plonky2_b200/csrc/gl_field.cuh"
using namespace gl;
typedef uint64_t u64;
struct L3 { uint32_t w0, w1; int32_t e; };

__device__ __forceinline__ L3 l_from(u64 x) { return L3{lo32(x), hi32(x), 0}; }
__device__ __forceinline__ L3 ladd(L3 a, L3 b) {
    L3 r;
    asm("add.cc.u32 %0, %3, %6;\n\taddc.cc.u32 %1, %4, %7;\n\taddc.u32 %2, %5, %8;"
        : "=r"(r.w0), "=r"(r.w1), "=r"(r.e) : "r"(a.w0), "r"(a.w1), "r"(a.e), "r"(b.w0), "r"(b.w1), "r"(b.e));
    return r;
}
__device__ __forceinline__ L3 lsub(L3 a, L3 b) {
    L3 r;
    asm("sub.cc.u32 %0, %3, %6;\n\tsubc.cc.u32 %1, %4, %7;\n\tsubc.u32 %2, %5, %8;"
        : "=r"(r.w0), "=r"(r.w1), "=r"(r.e) : "r"(a.w0), "r"(a.w1), "r"(a.e), "r"(b.w0), "r"(b.w1), "r"(b.e));
    return r;
}
// v * 2^S (mod p), lazy in, lazy out. S = 32q + r, 0 < S < 96. |v| < 2^90 or so in, |out| < 2^67.
template <int S>
__device__ __forceinline__ L3 lshift(L3 a) {
    constexpr int q = S / 32, r = S % 32;
    uint32_t c0, c1, c2; int32_t c3;
    if (r == 0) { c0 = a.w0; c1 = a.w1; c2 = (uint32_t)a.e; c3 = a.e >> 31; }
    else {
        c0 = a.w0 << r;
        c1 = __funnelshift_l(a.w0, a.w1, r);
        c2 = __funnelshift_l(a.w1, (uint32_t)a.e, r);
        c3 = a.e >> (32 - r);
    }
    // value = c0 + c1 X + c2 X^2 + c3 X^3 (c3 signed), times X^q, with X^2 = X - 1, X^3 = -1:
    //  q=0: (c0 - c2 - c3) + (c1 + c2) X
    //  q=1: (-c1 - c2)     + (c0 + c1 - c3) X
    //  q=2: (-c0 - c1 + c3) + (c0 - c2 - c3) X
    // written as  P + Q*X - T  with P,Q unsigned words and T a 64-bit signed quantity (t0 + t1*X):
    L3 o;
    const int32_t s3 = c3 >> 31;
    if (q == 0) {
        // base = c0 + (c1 + c2) X ; T = c2 + c3
        uint32_t t0, t1;
        asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=r"(t0), "=r"(t1) : "r"(c2), "r"((uint32_t)c3), "r"((uint32_t)s3));
        uint32_t b1, be;
        asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=r"(b1), "=r"(be) : "r"(c1), "r"(c2));
        asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, %7, %8;"
            : "=r"(o.w0), "=r"(o.w1), "=r"(o.e) : "r"(c0), "r"(b1), "r"(t0), "r"(t1), "r"(be), "r"((int32_t)t1 >> 31));
    } else if (q == 1) {
        // (c0 + c1 - c3) X - (c1 + c2):  base = (0, c0 + c1 (carry -> e)) ; T = (c1 + c2) + c3 * X
        uint32_t b1, be;
        asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=r"(b1), "=r"(be) : "r"(c0), "r"(c1));
        uint32_t t0, t1; int32_t t2;
        asm("add.cc.u32 %0, %3, %4;\n\taddc.cc.u32 %1, %5, 0;\n\taddc.u32 %2, %6, 0;" : "=r"(t0), "=r"(t1), "=r"(t2) : "r"(c1), "r"(c2), "r"((uint32_t)c3), "r"(s3));
        asm("sub.cc.u32 %0, 0, %3;\n\tsubc.cc.u32 %1, %4, %5;\n\tsubc.u32 %2, %6, %7;"
            : "=r"(o.w0), "=r"(o.w1), "=r"(o.e) : "r"(t0), "r"(b1), "r"(t1), "r"(be), "r"(t2));
    } else {
        // (c3 - c0 - c1) + (c0 - c2 - c3) X = c0 X - [ (c0 + c1 - c3) + (c2 + c3) X ]
        uint32_t t0, t1; int32_t t2;
        // u = c0 + c1 - c3 (64-bit signed), w = c2 + c3 (64-bit signed)
        uint32_t u0, u1;
        asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=r"(u0), "=r"(u1) : "r"(c0), "r"(c1));
        asm("sub.cc.u32 %0, %0, %2;\n\tsubc.u32 %1, %1, %3;" : "+r"(u0), "+r"(u1) : "r"((uint32_t)c3), "r"((uint32_t)s3));
        uint32_t w0_, w1_;
        asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=r"(w0_), "=r"(w1_) : "r"(c2), "r"((uint32_t)c3), "r"((uint32_t)s3));
        // T = u + w*X : t0 = u0, t1 = u1 + w0, t2 = sext(u1) + w1 + carry
        t0 = u0;
        asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, %5;" : "=r"(t1), "=r"(t2) : "r"(u1), "r"(w0_), "r"((int32_t)u1 >> 31), "r"(w1_));
        asm("sub.cc.u32 %0, 0, %3;\n\tsubc.cc.u32 %1, %4, %5;\n\tsubc.u32 %2, 0, %6;"
            : "=r"(o.w0), "=r"(o.w1), "=r"(o.e) : "r"(t0), "r"(c0), "r"(t1), "r"(t2));
    }
    return o;
}
// lazy -> u64 (any representative < 2^64); requires |e| < 2^20
__device__ __forceinline__ u64 lnorm(L3 a) {
    // make e non-negative by adding K = 2^20 * p = 2^84 - 2^52 + 2^20 (3 words: 2^20, -2^20 (mod 2^32) with borrow..)
    // K = 2^20 + (2^32 - 2^20) * 2^32 + (2^20 - 1) * 2^64
    uint32_t w0, w1, e;
    asm("add.cc.u32 %0, %3, 0x00100000;\n\taddc.cc.u32 %1, %4, 0xfff00000;\n\taddc.u32 %2, %5, 0x000fffff;"
        : "=r"(w0), "=r"(w1), "=r"(e) : "r"(a.w0), "r"(a.w1), "r"(a.e));
    return reduce96(pack64(w0, w1), e);
}
template <int M>
__device__ __forceinline__ void dftL(L3* r) {
#pragma unroll
    for (int l = 0; l < M; l++) {
        const int half = 1 << (M - 1 - l);
#pragma unroll
        for (int b = 0; b < (1 << (M - 1)); b++) {
            const int j = b % half, blk = (b / half) * 2 * half;
            L3 u = r[blk + j], v = r[blk + j + half];
            r[blk + j] = ladd(u, v);
            L3 d = lsub(u, v);
            constexpr int dummy = 0; (void)dummy;
            const int k = (96 / half) * j;
            switch (k) {
#define CS(K) case K: d = lshift<K>(d); break;
                CS(6) CS(12) CS(18) CS(24) CS(30) CS(36) CS(42) CS(48) CS(54) CS(60) CS(66) CS(72) CS(78) CS(84) CS(90)
                CS(3) CS(9) CS(15) CS(21) CS(27) CS(33) CS(39) CS(45) CS(51) CS(57) CS(63) CS(69) CS(75) CS(81) CS(87) CS(93)
                default: break;
            }
            r[blk + j + half] = d;
        }
    }
}
__global__ void k_dft32L(const u64* in, u64* out) {
    L3 r[32];
    const size_t t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = l_from(in[t + (size_t)i * 4096]);
    dftL<5>(r);
#pragma unroll
    for (int i = 0; i < 32; i++) out[t + (size_t)i * 4096] = lnorm(r[i]);
}
template <int S> __global__ void k_lshift(const u64* in, u64* out) {
    L3 a = l_from(in[threadIdx.x]); a.e = (int32_t)in[threadIdx.x + 32];
    L3 o = lshift<S>(a);
    out[threadIdx.x] = pack64(o.w0, o.w1); out[threadIdx.x + 32] = o.e;
}
template __global__ void k_lshift<12>(const u64*, u64*);
template __global__ void k_lshift<48>(const u64*, u64*);
template __global__ void k_lshift<72>(const u64*, u64*);
__global__ void k_lnorm(const u64* in, u64* out) {
    L3 a = l_from(in[threadIdx.x]); a.e = (int32_t)in[threadIdx.x + 32];
    out[threadIdx.x] = lnorm(a);
}
