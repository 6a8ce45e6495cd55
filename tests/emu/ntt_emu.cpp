// CPU emulation of the CUDA NTT tile code (threads as loops, phases as barriers) checked against
// the oracle. Test infrastructure: built and run by tests/test_emu_ntt.py.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../plonky2_b200/csrc/gl_ntt.cuh"
#include "../../oracle/gl_oracle.h"
using namespace gl;

template <int LOG>
void run_tile_steps(uint64_t* s, const uint64_t* wt, int nthreads) {
    for (int i = 0; i < ntt_num_steps(LOG); i++)
        for (int tid = 0; tid < nthreads; tid++) tile_step<LOG>(s, wt, i, tid, nthreads);
}
template <int LOG>
void emu_passA(const PassA& pa, int nblocks) {
    const int NT = ntt_tile_threads(LOG);
    std::vector<uint64_t> s((size_t)(1 << LOG) * ntt_tile_TS(LOG));
    for (int blk = 0; blk < nblocks; blk++) {
        for (int tid = 0; tid < NT; tid++) passA_load<LOG>(pa, s.data(), blk, tid, NT);
        run_tile_steps<LOG>(s.data(), pa.wt, NT);
        for (int tid = 0; tid < NT; tid++) passA_store<LOG>(pa, s.data(), blk, tid, NT);
    }
}
template <int LOG, int MODE>
void emu_passB(const PassB& pb) {
    const int NT = ntt_tile_threads(LOG);
    std::vector<uint64_t> s((size_t)(1 << LOG) * ntt_tile_TS(LOG));
    int nblocks = passB_blocks<LOG>(pb, MODE);
    for (int blk = 0; blk < nblocks; blk++) {
        for (int tid = 0; tid < NT; tid++) passB_load<LOG, MODE>(pb, s.data(), blk, tid, NT);
        run_tile_steps<LOG>(s.data(), pb.wt, NT);
        for (int tid = 0; tid < NT; tid++) passB_store<LOG, MODE>(pb, s.data(), blk, tid, NT);
    }
}
#define DISPATCH(LOGV, CALL)                                                   \
    switch (LOGV) {                                                            \
        case 1: { constexpr int L = 1; CALL; } break;                          \
        case 2: { constexpr int L = 2; CALL; } break;                          \
        case 3: { constexpr int L = 3; CALL; } break;                          \
        case 4: { constexpr int L = 4; CALL; } break;                          \
        case 5: { constexpr int L = 5; CALL; } break;                          \
        case 6: { constexpr int L = 6; CALL; } break;                          \
        case 7: { constexpr int L = 7; CALL; } break;                          \
        case 8: { constexpr int L = 8; CALL; } break;                          \
        case 9: { constexpr int L = 9; CALL; } break;                          \
        case 10: { constexpr int L = 10; CALL; } break;                        \
        case 11: { constexpr int L = 11; CALL; } break;                        \
        case 12: { constexpr int L = 12; CALL; } break;                        \
        default: abort();                                                      \
    }

static std::vector<uint64_t> wt_table(int log) {
    std::vector<uint64_t> t((size_t)1 << log);
    for (size_t j = 0; j < t.size(); j++) t[j] = table_wt_entry(log, (uint32_t)j);
    return t;
}
static uint64_t rnd(uint64_t& st) {
    st += 0x9E3779B97F4A7C15ULL;
    uint64_t z = st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// mode: 0 forward natural, 1 inverse natural, 2 coset-LDE leaves (rate_bits r)
static int check(int log_n, int ncols, int mode, int rate_bits) {
    const size_t n = (size_t)1 << log_n;
    int a, b;
    ntt_split(log_n, a, b);
    uint64_t st = 1234 + log_n * 77 + mode;
    std::vector<uint64_t> in((size_t)ncols * n);
    for (auto& x : in) x = rnd(st);
    std::vector<uint64_t> wa = a ? wt_table(a) : std::vector<uint64_t>(), wb = wt_table(b);
    std::vector<uint64_t> twa;
    if (a) {
        twa.resize(n);
        for (size_t i = 0; i < n; i++) twa[i] = table_twa_entry(a, b, i);
    }
    const int ncos = mode == 2 ? (1 << rate_bits) : 1;
    const size_t N = n * ncos, W = ncols;
    std::vector<uint64_t> out(mode == 2 ? N * W : (size_t)ncols * n), tmp((size_t)ncols * n);
    for (int c = 0; c < ncos; c++) {
        std::vector<uint64_t> u, v;
        uint64_t shift = 1;
        if (mode == 2) {
            const uint64_t wN = root_of_unity(log_n + rate_bits);
            shift = mul(MULTIPLICATIVE_GROUP_GENERATOR, pow(wN, bitrev32(c, rate_bits)));
        }
        const uint64_t* src = in.data();
        std::vector<uint64_t> scaled;
        if (a) {
            PassA pa{};
            pa.in = in.data(); pa.out = tmp.data(); pa.in_stride = n; pa.out_stride = n;
            pa.twa = twa.data(); pa.wt = wa.data(); pa.log_c = b;
            pa.tiles_per_col = (1 << b) / ntt_tile_T(a);
            if (mode == 2) {
                u.resize((size_t)1 << a); v.resize((size_t)1 << b);
                uint64_t sC = pow(shift, (uint64_t)1 << b);
                for (size_t i = 0; i < u.size(); i++) u[i] = pow(sC, i);
                for (size_t i = 0; i < v.size(); i++) v[i] = pow(shift, i);
                pa.u = u.data(); pa.v = v.data();
            }
            DISPATCH(a, emu_passA<L>(pa, ncols * pa.tiles_per_col));
            src = tmp.data();
        } else if (mode == 2) {
            scaled = in;
            for (int col = 0; col < ncols; col++)
                for (size_t j = 0; j < n; j++) scaled[col * n + j] = mul(scaled[col * n + j], pow(shift, j));
            src = scaled.data();
        }
        PassB pb{};
        pb.in = src; pb.in_stride = n; pb.out = out.data(); pb.wt = wb.data(); pb.log_r = a;
        pb.ncols = ncols; pb.scale = 1;
        if (mode == 2) {
            pb.out_stride = W; pb.row0 = (size_t)c * n; pb.col0 = 0;
            DISPATCH(b, (emu_passB<L, PB_LEAVES>(pb)));
        } else {
            pb.out_stride = n;
            if (mode == 1) { pb.reverse = 1; pb.scale = inverse_2exp(log_n); }
            if (a) { DISPATCH(b, (emu_passB<L, PB_NATURAL>(pb))); }
            else { DISPATCH(b, (emu_passB<L, PB_NATURAL_COLS>(pb))); }
        }
    }
    // ---- oracle
    int bad = 0;
    if (mode != 2) {
        for (int col = 0; col < ncols; col++) {
            std::vector<uint64_t> ref(in.begin() + col * n, in.begin() + (col + 1) * n);
            if (mode == 0) glo_fft(ref.data(), log_n, 0); else glo_ifft(ref.data(), log_n);
            for (size_t i = 0; i < n; i++) bad += ref[i] != out[col * n + i];
        }
    } else {
        glo_commit* cm = glo_commit_new(in.data(), n, ncols, log_n, rate_bits, 0, nullptr, 1, 4);
        const uint64_t* leaves = glo_commit_leaves(cm);
        for (size_t i = 0; i < N * W; i++) bad += leaves[i] != out[i];
        glo_commit_free(cm);
    }
    printf("log_n=%d cols=%d mode=%d r=%d : %s (%d mismatches)\n", log_n, ncols, mode, rate_bits, bad ? "FAIL" : "ok", bad);
    return bad != 0;
}

int main(int argc, char** argv) {
    int maxlog = argc > 1 ? atoi(argv[1]) : 15;
    int fails = 0;
    for (int lg = 1; lg <= maxlog; lg++) {
        fails += check(lg, 3, 0, 0);
        fails += check(lg, 11, 1, 0);
        fails += check(lg, 10, 2, lg % 3 + 1);
    }
    fails += check(0 + 1, 1, 2, 3);
    printf(fails ? "EMU FAILED\n" : "EMU OK\n");
    return fails ? 1 : 0;
}
