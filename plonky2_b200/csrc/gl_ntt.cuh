// gl_ntt.cuh -- batched Goldilocks NTT as a four-step (six-step without the explicit transposes)
// decomposition n = R * C, R = 2^a (strided pass "A"), C = 2^b (contiguous pass "B").
//
// Replaces (reference, CPU): fft_classic / ifft_with_options      field/src/fft.rs:68-202
//                            coset_fft_with_options, lde          field/src/polynomial/mod.rs:199-201,280-293
//                            lde_values + transpose + bit-reverse plonky2/src/fri/oracle.rs:97-98,114-139
//
//   X[k1 + R*k2] = sum_{j2<C} w_n^{j2*k1} w_C^{j2*k2} ( sum_{j1<R} x[j1*C + j2] w_R^{j1*k1} )
//
// Pass A: a CTA owns a tile of T adjacent j2 and all R values of j1 (global accesses are T*8-byte
//         segments), runs T interleaved R-point DIF transforms in shared memory, multiplies by
//         w_n^{j2*k1} and writes Y[p][j2] with p = bitrev_a(k1) (the in-place DIF order).
// Pass B: a CTA owns T "lines" of C contiguous elements (rows p of one column, or the same row of
//         T adjacent columns), runs T interleaved C-point DIF transforms and stores either
//           - natural order  out[k1 + R*k2]            (NTT / iNTT API, optional index reversal), or
//           - leaf-major     leaves[row0 + p*C + q][c] (LDE; q = bitrev_b(k2), i.e. exactly the
//             bit-reversed row order the reference produces with transpose + reverse_index_bits).
// In-tile transforms are radix-16 register butterflies whose internal twiddles are powers of two
// (w_16 = 2^12) - shifts, no multiplies; only the inter-step twiddles are
// general 64x64 multiplies, read from a shared-memory table staged by a TMA bulk copy.
//
// Every per-thread phase below is a plain function of (tid, nthreads) so that tests/emu can run
// the same code on the CPU (threads as a loop, phases as barriers) to check indexing.
#pragma once
#include "gl_field.cuh"

namespace gl {

constexpr int NTT_MAX_LOG_TILE = 12;  // largest in-CTA transform

GL_HD constexpr int ntt_tile_T(int log) { return log >= 12 ? 4 : 8; }
GL_HD constexpr int ntt_tile_TS(int log) { return ntt_tile_T(log) + 1; }  // odd stride: conflict-free columns
GL_HD constexpr int ntt_tile_threads(int log) {
    int n = ((1 << log) * ntt_tile_T(log)) / 16;
    return n < 32 ? 32 : (n > 1024 ? 1024 : n);
}
// resident CTAs per SM the kernels are compiled for (register cap = 65536 / (threads * blocks))
GL_HD constexpr int ntt_tile_min_blocks(int log) {
    int b = 1024 / ntt_tile_threads(log);
    return b < 1 ? 1 : (b > 8 ? 8 : b);
}
GL_HD constexpr size_t ntt_tile_smem_bytes(int log) {
    // data tile + full-cycle twiddle table + mbarrier slot
    return ((size_t)(1 << log) * ntt_tile_TS(log) + (size_t)(1 << log)) * 8 + 16;
}

// 2^M-point DIF DFT in registers, natural in, bit-reversed out, w_{2^M} = 2^(192 / 2^M).
template <int M>
GL_HD void dft_regs(uint64_t* r) {
#pragma unroll
    for (int l = 0; l < M; l++) {
        const int half = 1 << (M - 1 - l);
#pragma unroll
        for (int blk = 0; blk < (1 << M); blk += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; j++) {
                uint64_t u = r[blk + j], v = r[blk + j + half];
                r[blk + j] = add(u, v);
                uint64_t d = sub(u, v);
                r[blk + j + half] = mul_pow2(d, (uint32_t)((96 / half) * j));
            }
        }
    }
}

// One radix-2^M DIF step over index bits [sbit-M+1 .. sbit] of T interleaved 2^LOG-point
// transforms stored as s[i*TS + t]. wt = full-cycle table w_{2^LOG}^j, j < 2^LOG.
template <int LOG, int M>
GL_HD void radix_step(uint64_t* s, const uint64_t* wt, int sbit, int tid, int nthreads) {
    constexpr int T = ntt_tile_T(LOG), TS = ntt_tile_TS(LOG);
    const int sh = sbit - M + 1;
    const int ngroups = (1 << (LOG - M)) * T;
    for (int g = tid; g < ngroups; g += nthreads) {
        const int t = g % T, gi = g / T;
        const int low = gi & ((1 << sh) - 1), high = gi >> sh;
        const int base = (high << (sbit + 1)) | low;
        uint64_t r[1 << M];
#pragma unroll
        for (int q = 0; q < (1 << M); q++) r[q] = s[(base | (q << sh)) * TS + t];
        dft_regs<M>(r);
        if (sh > 0) {
            // slot q holds frequency k = bitrev_M(q); twiddle w_{2^(sbit+1)}^{low*k}
            const int step = low << (LOG - sbit - 1);
#pragma unroll
            for (int q = 1; q < (1 << M); q++) {
                const int k = (int)bitrev32((uint32_t)q, M);
                r[q] = mul(r[q], wt[(step * k) & ((1 << LOG) - 1)]);
            }
        }
#pragma unroll
        for (int q = 0; q < (1 << M); q++) s[(base | (q << sh)) * TS + t] = r[q];
    }
}

// Number of radix steps and the M of step i for a 2^LOG transform: 4,4,...,rem.
GL_HD constexpr int ntt_num_steps(int log) { return (log + 3) / 4; }

// ---------------------------------------------------------------- pass descriptors
struct PassA {
    const uint64_t* in;    // column b at in + b*in_stride, natural order
    uint64_t* out;         // column b at out + b*out_stride, layout [p][j2]
    size_t in_stride, out_stride;
    const uint64_t* twa;   // n entries: w_n^{bitrev_a(p)*j2} at [p*C + j2]
    const uint64_t* u;     // optional coset scale (s^C)^{j1}, R entries (nullptr = none)
    const uint64_t* v;     // optional coset scale s^{j2}, C entries
    const uint64_t* wt;    // full-cycle table for 2^a
    int log_c;             // b
    int tiles_per_col;     // C / T
};

template <int LOG>
GL_HD void passA_load(const PassA& pa, uint64_t* s, int blk, int tid, int nthreads) {
    constexpr int T = ntt_tile_T(LOG), TS = ntt_tile_TS(LOG);
    const int col = blk / pa.tiles_per_col, tile = blk % pa.tiles_per_col;
    const size_t C = (size_t)1 << pa.log_c;
    const uint64_t* src = pa.in + (size_t)col * pa.in_stride + (size_t)tile * T;
    for (int e = tid; e < (1 << LOG) * T; e += nthreads) {
        const int j1 = e / T, tt = e % T;
        uint64_t x = src[(size_t)j1 * C + tt];
        if (pa.u) x = mul(x, pa.u[j1]);
        s[j1 * TS + tt] = x;
    }
}
template <int LOG>
GL_HD void passA_store(const PassA& pa, const uint64_t* s, int blk, int tid, int nthreads) {
    constexpr int T = ntt_tile_T(LOG), TS = ntt_tile_TS(LOG);
    const int col = blk / pa.tiles_per_col, tile = blk % pa.tiles_per_col;
    const size_t C = (size_t)1 << pa.log_c;
    uint64_t* dst = pa.out + (size_t)col * pa.out_stride + (size_t)tile * T;
    const uint64_t* tw = pa.twa + (size_t)tile * T;
    for (int e = tid; e < (1 << LOG) * T; e += nthreads) {
        const int p = e / T, tt = e % T;
        uint64_t y = mul(s[p * TS + tt], tw[(size_t)p * C + tt]);
        if (pa.v) y = mul(y, pa.v[tile * T + tt]);
        dst[(size_t)p * C + tt] = y;
    }
}

enum PassBMode { PB_NATURAL = 0, PB_NATURAL_COLS = 1, PB_LEAVES = 2 };

struct PassB {
    const uint64_t* in;   // column b at in + b*in_stride, layout [p][j2] (p = row of C elements)
    size_t in_stride;
    uint64_t* out;
    size_t out_stride;    // natural modes: column stride; leaves mode: leaf width W
    const uint64_t* wt;   // full-cycle table for 2^b
    const uint64_t* pre;  // single-pass only (log_r == 0): optional coset pre-scale s^j, n entries
    int log_r;            // a (0 for single-pass)
    int ncols;            // number of columns in this launch
    int reverse;          // natural modes: write to (n - k) mod n   (ifft index reversal, fft.rs:80-90)
    uint64_t scale;       // natural modes: multiply outputs (1 = none)  (n^-1 for the inverse)
    size_t row0;          // leaves mode: first leaf row of this coset
    int col0;             // leaves mode: first leaf column of this launch
};

// number of CTAs for a pass-B launch
template <int LOG>
GL_HD int passB_blocks(const PassB& pb, int mode) {
    constexpr int T = ntt_tile_T(LOG);
    const int R = 1 << pb.log_r;
    if (mode == PB_NATURAL) return pb.ncols * (R / T);
    const int ctiles = (pb.ncols + T - 1) / T;
    return ctiles * R;  // PB_NATURAL_COLS has R == 1
}

// line t of CTA blk -> (column, row p); returns false if the line is past the end
template <int LOG, int MODE>
GL_HD bool passB_line(const PassB& pb, int blk, int t, int& col, int& p, int& k1) {
    constexpr int T = ntt_tile_T(LOG);
    const int R = 1 << pb.log_r;
    if (MODE == PB_NATURAL) {
        const int per_col = R / T;
        col = blk / per_col;
        k1 = (blk % per_col) * T + t;
        p = (int)bitrev32((uint32_t)k1, pb.log_r);
        return true;
    } else {
        const int ctile = blk / R;
        p = blk % R;
        k1 = (int)bitrev32((uint32_t)p, pb.log_r);
        col = ctile * T + t;
        return col < pb.ncols;
    }
}

template <int LOG, int MODE>
GL_HD void passB_load(const PassB& pb, uint64_t* s, int blk, int tid, int nthreads) {
    constexpr int T = ntt_tile_T(LOG), TS = ntt_tile_TS(LOG);
    for (int e = tid; e < (1 << LOG) * T; e += nthreads) {
        const int t = e >> LOG, i = e & ((1 << LOG) - 1);
        int col, p, k1;
        uint64_t x = 0;
        if (passB_line<LOG, MODE>(pb, blk, t, col, p, k1))
            x = pb.in[(size_t)col * pb.in_stride + ((size_t)p << LOG) + i];
        if (pb.pre) x = mul(x, pb.pre[i]);
        s[i * TS + t] = x;
    }
}

template <int LOG, int MODE>
GL_HD void passB_store(const PassB& pb, const uint64_t* s, int blk, int tid, int nthreads) {
    constexpr int T = ntt_tile_T(LOG), TS = ntt_tile_TS(LOG);
    const size_t n = (size_t)1 << (LOG + pb.log_r);
    for (int e = tid; e < (1 << LOG) * T; e += nthreads) {
        int t, q;
        if (MODE == PB_NATURAL_COLS) {
            // consecutive threads -> consecutive output index k2
            t = e >> LOG;
            q = (int)bitrev32((uint32_t)(e & ((1 << LOG) - 1)), LOG);
        } else {
            t = e % T;
            q = e / T;
        }
        int col, p, k1;
        if (!passB_line<LOG, MODE>(pb, blk, t, col, p, k1)) continue;
        uint64_t y = s[q * TS + t];
        if (MODE == PB_LEAVES) {
            pb.out[(pb.row0 + ((size_t)p << LOG) + q) * pb.out_stride + pb.col0 + col] = canon(y);
        } else {
            const size_t k2 = bitrev32((uint32_t)q, LOG);
            size_t k = (size_t)k1 + (k2 << pb.log_r);
            if (pb.reverse) k = (n - k) & (n - 1);
            if (pb.scale != 1) y = mul(y, pb.scale);
            pb.out[(size_t)col * pb.out_stride + k] = canon(y);
        }
    }
}

// ---------------------------------------------------------------- all radix steps of a tile
// (callers put a barrier between consecutive calls: step index `i`)
template <int LOG>
GL_HD void tile_step(uint64_t* s, const uint64_t* wt, int i, int tid, int nthreads) {
    const int sbit = LOG - 1 - 4 * i;
    if (i < LOG / 4) {
        if constexpr (LOG >= 4) radix_step<LOG, 4>(s, wt, sbit, tid, nthreads);
    } else {
        constexpr int REM = LOG % 4;
        if constexpr (REM > 0) radix_step<LOG, REM>(s, wt, sbit, tid, nthreads);
    }
}

// ---------------------------------------------------------------- plan + table entries
// n = 2^log_n = R*C with R = 2^a (pass A, absent when a == 0) and C = 2^b (pass B).
GL_HD void ntt_split(int log_n, int& a, int& b, int force_b = 0) {
    if (log_n <= NTT_MAX_LOG_TILE) {
        a = 0;
        b = log_n;
    } else {
        b = (log_n + 1) / 2;
        // tuning override (gl_ctx_set_ntt_split): any b with 6 <= a, b <= 12
        if (force_b >= 6 && force_b <= NTT_MAX_LOG_TILE && log_n - force_b >= 6 && log_n - force_b <= NTT_MAX_LOG_TILE)
            b = force_b;
        a = log_n - b;
    }
}
// full-cycle in-tile table: w_{2^log}^j
GL_HD uint64_t table_wt_entry(int log, uint32_t j) { return pow(root_of_unity((uint32_t)log), j); }
// pass-A post-twiddle: w_n^{bitrev_a(p) * j2} at index p*C + j2
GL_HD uint64_t table_twa_entry(int a, int b, size_t idx) {
    const size_t p = idx >> b, j2 = idx & (((size_t)1 << b) - 1);
    const uint64_t k1 = bitrev32((uint32_t)p, (uint32_t)a);
    return pow(root_of_unity((uint32_t)(a + b)), k1 * j2);
}

}  // namespace gl
