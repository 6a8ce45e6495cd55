#include "../../plonky2_b200/csrc/gl_field.cuh"
using namespace gl;
typedef uint64_t u64;
// reduce96 via one IMAD.WIDE: lo + hi*EPS
__device__ __forceinline__ u64 reduce96_w(u64 lo, uint32_t hi) {
    const unsigned __int128 t = (unsigned __int128)lo + (u64)hi * 0xFFFFFFFFull;
    const u64 r = (u64)t;
    const uint32_t c = (uint32_t)(t >> 64);
    // + c * EPS (cannot wrap)
    return r + (u64)(0u - c);
}
__global__ void k_r96(const u64* in, u64* out) { out[threadIdx.x] = reduce96(in[threadIdx.x], (uint32_t)in[threadIdx.x + 32]); }
__global__ void k_r96w(const u64* in, u64* out) { out[threadIdx.x] = reduce96_w(in[threadIdx.x], (uint32_t)in[threadIdx.x + 32]); }
__global__ void k_base(const u64* in, u64* out) { out[threadIdx.x] = in[threadIdx.x] ^ (uint32_t)in[threadIdx.x + 32]; }
