// Host run of the FP64-pipe formulation of Poseidon (the MDS layers of the full rounds and the FP64-resident
// partial rounds, gl_poseidon.cuh built with -DGL_FP64_ON_HOST) against the oracle: IEEE doubles and fma give
// the same bits on the CPU, so the exactness argument (all limbs are integers < 2^53) is checked without a GPU.
// Test infrastructure: built and run by tests/test_emu.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
static double g_max_limb = 0;
#define GL_F64_TRACK(x) do { const double a_ = std::fabs(x); if (a_ > g_max_limb) g_max_limb = a_; } while (0)
#include "../../plonky2_b200/csrc/gl_poseidon.cuh"
#include "../../oracle/gl_oracle.h"
#if !defined(GL_PARTIAL_F64)
#error "build with -DGL_FP64_ON_HOST"
#endif
using namespace gl;
static uint64_t rnd(uint64_t& st) {
    st += 0x9E3779B97F4A7C15ULL;
    uint64_t z = st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 200000;
    const uint64_t edge[] = {0, 1, 2, P - 1, P - 2, P, P + 1, 0xFFFFFFFFULL, 0x100000000ULL, 1ULL << 63,
                             P - (1ULL << 32), ~0ULL, ~0ULL - 1, 0xFFFFFFFF00000000ULL, 0xFFFFFFFEFFFFFFFFULL};
    const int ne = sizeof(edge) / sizeof(edge[0]);
    int bad = 0;
    uint64_t st = 7;
    for (int i = 0; i < iters; i++) {
        uint64_t s[12], r[12];
        for (int k = 0; k < 12; k++) {
            uint64_t v = rnd(st);
            if (i % 5 == 1) v = edge[v % ne];
            if (i % 5 == 2) v = ~0ULL - (v & 7);
            if (i % 5 == 3) v |= 0xFFFFFFF0FFFFFFF0ULL;  // both halves near 2^32: largest limbs
            s[k] = r[k] = v;
        }
        poseidon_permute(s);
        if (i & 1) glo_poseidon(r); else glo_poseidon_naive(r);
        for (int k = 0; k < 12; k++) bad += canon(s[k]) != r[k];
    }
    // the four known-answer vectors of the reference (poseidon_goldilocks.rs:466-487) through the FP64 form
    {
        uint64_t z[12] = {0}, r[12] = {0};
        poseidon_permute(z);
        glo_poseidon(r);
        for (int k = 0; k < 12; k++) bad += canon(z[k]) != r[k];
        bad += canon(z[0]) != 0x3c18a9786cb0b359ULL;  // first word of the all-zero KAT
    }
    // sponge through the same path
    for (uint32_t W = 0; W < 300; W += 7) {
        uint64_t in[300], o1[4], o2[4];
        for (auto& x : in) x = rnd(st);
        hash_or_noop_strided<true>(in, 1, W, o1);
        glo_hash_or_noop(in, W, o2);
        for (int k = 0; k < 4; k++) bad += o1[k] != o2[k];
    }
    const double lim = 9007199254740992.0;  // 2^53
    printf("largest FP64 limb magnitude seen: 2^%.2f (bound 2^51, exact below 2^53)\n", std::log2(g_max_limb));
    if (!(g_max_limb < lim / 4)) bad++;
    printf(bad ? "POSEIDON F64 EMU FAILED (%d)\n" : "POSEIDON F64 EMU OK\n", bad);
    return bad != 0;
}
