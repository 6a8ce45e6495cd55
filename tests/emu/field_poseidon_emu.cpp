// Host run of the device formulation of the field ops and Poseidon (GL_FORCE_32BIT_PATH) against the
// oracle. Test infrastructure: built and run by tests/test_emu.py.
#include <cstdio>
#include <cstdlib>
#include "../../plonky2_b200/csrc/gl_poseidon.cuh"
#include "../../oracle/gl_oracle.h"
using namespace gl;
static uint64_t rnd(uint64_t& st) {
    st += 0x9E3779B97F4A7C15ULL;
    uint64_t z = st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
int main() {
    const uint64_t edge[] = {0, 1, 2, P - 1, P - 2, P, P + 1, 0xFFFFFFFFULL, 0x100000000ULL, 1ULL << 63,
                             P - (1ULL << 32), ~0ULL, ~0ULL - 1, 0xFFFFFFFF00000000ULL, 0xFFFFFFFEFFFFFFFFULL};
    const int ne = sizeof(edge) / sizeof(edge[0]);
    int bad = 0;
    uint64_t st = 42;
    auto chk2 = [&](uint64_t a, uint64_t b) {
        bad += canon(add(a, b)) != glo_add(a, b);
        bad += canon(sub(a, b)) != glo_sub(a, b);
        bad += canon(mul(a, b)) != glo_mul(a, b);
        bad += canon(sqr(a)) != glo_mul(a, a);
        bad += canon(mul_add(a, b, a ^ b)) != glo_add(glo_mul(a, b), a ^ b);
        bad += neg(a) != glo_neg(a);
    };
    for (int i = 0; i < ne; i++)
        for (int j = 0; j < ne; j++) chk2(edge[i], edge[j]);
    for (int i = 0; i < 200000; i++) chk2(rnd(st), rnd(st));
    for (int i = 0; i < ne; i++)
        for (uint32_t k = 0; k < 96; k++) {
            uint64_t want = glo_mul(edge[i], glo_exp(2, k));
            bad += canon(mul_pow2(edge[i], k)) != want;
        }
    for (int i = 0; i < 20000; i++) {
        uint64_t a = rnd(st);
        uint32_t k = rnd(st) % 96;
        bad += canon(mul_pow2(a, k)) != glo_mul(a, glo_exp(2, k));
    }
    // reduce96 / 160-bit accumulator
    for (int i = 0; i < 20000; i++) {
        Acc160 acc = {0, 0, 0};
        uint64_t want = 0;
        for (int t = 0; t < 12; t++) {
            uint64_t x = (i & 1) ? ~0ULL - (rnd(st) & 3) : rnd(st), y = (i & 2) ? ~0ULL - (rnd(st) & 3) : rnd(st);
            acc_mul(acc, x, y);
            want = glo_add(want, glo_mul(x, y));
        }
        bad += canon(acc_reduce(acc)) != want;
    }
    // extension field
    for (int i = 0; i < 2000; i++) {
        uint64_t a[2] = {rnd(st), rnd(st)}, b[2] = {rnd(st), rnd(st)}, o[2];
        glo_ext2_mul(a, b, o);
        E2 r = e2_mul(E2{a[0], a[1]}, E2{b[0], b[1]});
        bad += canon(r.a) != o[0] || canon(r.b) != o[1];
        glo_ext2_inv(a, o);
        E2 q = e2_inv(E2{a[0], a[1]});
        bad += canon(q.a) != o[0] || canon(q.b) != o[1];
    }
    // Poseidon vs oracle (incl. non-canonical and extreme inputs)
    for (int i = 0; i < 3000; i++) {
        uint64_t s[12], r[12];
        for (int k = 0; k < 12; k++) {
            uint64_t v = rnd(st);
            if (i % 5 == 1) v = edge[v % ne];
            if (i % 5 == 2) v = ~0ULL - (v & 7);
            s[k] = r[k] = v;
        }
        poseidon_permute(s);
        glo_poseidon(r);
        for (int k = 0; k < 12; k++) bad += canon(s[k]) != r[k];
    }
    // sponge
    for (uint32_t W = 0; W < 40; W++) {
        uint64_t in[40], o1[4], o2[4];
        for (auto& x : in) x = rnd(st);
        hash_or_noop_strided<true>(in, 1, W, o1);
        glo_hash_or_noop(in, W, o2);
        for (int k = 0; k < 4; k++) bad += o1[k] != o2[k];
        hash_or_noop_strided<false>(in, 1, W, o1);
        glo_hash_no_pad(in, W, o2);
        for (int k = 0; k < 4; k++) bad += o1[k] != o2[k];
    }
    printf(bad ? "FIELD/POSEIDON EMU FAILED (%d)\n" : "FIELD/POSEIDON EMU OK\n", bad);
    return bad != 0;
}
