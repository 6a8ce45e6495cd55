// C++ host-layer parity test (the reference's host code is compiled Rust; this is the compiled-language
// mirror of its interface, include/plonky2_b200.hpp). Reads like the reference's own tests:
// fft_and_ifft (field/src/fft.rs:215-249), merkle proofs (hash/merkle_tree.rs:269-311) and a prove/verify
// round trip. Everything goes through the C ABI; the oracle is the checker.
#include <cstdio>
#include <cstdlib>

#include "../../include/plonky2_b200.hpp"
#include "../../oracle/gl_oracle.h"

using namespace plonky2_b200;

static uint64_t rnd_state = 0x1234;
static F rnd() {
    rnd_state += 0x9E3779B97F4A7C15ULL;
    uint64_t z = rnd_state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z % ORDER;
}
#define REQUIRE(cond)                                                        \
    do {                                                                     \
        if (!(cond)) {                                                       \
            fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main() {
    Context ctx(0);
    // ---- fft_and_ifft (fft.rs:215-249): degree 200 padded to 256, coefficients i*1337 % 100
    {
        std::vector<F> coeffs(256, 0);
        for (int i = 0; i < 200; i++) coeffs[i] = (i * 1337) % 100;
        std::vector<F> naive(256);
        glo_naive_coset_eval(coeffs.data(), 8, 1, naive.data());
        auto points = fft_with_options(ctx, coeffs);
        REQUIRE(points == naive);
        REQUIRE(ifft_with_options(ctx, points) == coeffs);
        auto shifted = coset_fft(ctx, coeffs, 7);
        glo_naive_coset_eval(coeffs.data(), 8, 7, naive.data());
        REQUIRE(shifted == naive);
    }
    // ---- PolynomialBatch::from_values vs the oracle, Merkle proofs for a few leaves
    const uint32_t log_n = 9, r = 3, h = 2;
    const size_t n = size_t(1) << log_n, N = n << r;
    const size_t Bs[2] = {6, 4};
    std::vector<std::vector<std::vector<F>>> vals(2);
    std::vector<PolynomialBatch> batches;
    std::vector<glo_commit*> ocommits;
    for (int o = 0; o < 2; o++) {
        vals[o].assign(Bs[o], std::vector<F>(n));
        std::vector<F> flat;
        for (auto& col : vals[o]) {
            for (auto& x : col) x = rnd();
            flat.insert(flat.end(), col.begin(), col.end());
        }
        batches.push_back(PolynomialBatch::from_values(ctx, vals[o], r, false, h));
        ocommits.push_back(glo_commit_new(flat.data(), n, Bs[o], log_n, r, h, nullptr, 0, 4));
        const PolynomialBatch& b = batches.back();
        MerkleCap cap = b.cap();
        REQUIRE(std::memcmp(cap.hashes[0].elements, glo_commit_cap(ocommits[o]), 32 << h) == 0);
        auto polys = b.polynomials();
        for (size_t k = 0; k < Bs[o]; k++)
            REQUIRE(std::memcmp(polys[k].data(), glo_commit_coeffs(ocommits[o]) + k * n, n * 8) == 0);
        std::vector<F> lde(Bs[o]);
        glo_commit_get_lde_values(ocommits[o], 5, 2, lde.data());
        REQUIRE(b.get_lde_values(5, 2) == lde);
        std::vector<std::vector<F>> leaves;
        std::vector<MerkleProof> proofs;
        b.open({0, 77, N - 1}, leaves, proofs);
        const uint64_t idx[3] = {0, 77, N - 1};
        for (int q = 0; q < 3; q++) {
            REQUIRE(proofs[q].siblings.size() == log_n + r - h);
            REQUIRE(glo_merkle_verify(leaves[q].data(), leaves[q].size(), idx[q], proofs[q].siblings[0].elements,
                                      proofs[q].siblings.size(), cap.hashes[0].elements, h) == 1);
        }
    }
    // ---- shape errors mirror the reference's panics
    {
        bool threw = false;
        try {
            PolynomialBatch::from_values(ctx, {std::vector<F>(12, 1)}, 1, false, 0);  // "Not a power of two"
        } catch (const ShapeError&) { threw = true; }
        REQUIRE(threw);
        threw = false;
        try {
            PolynomialBatch::from_values(ctx, {std::vector<F>(8, 1)}, 1, false, 9);  // cap_height > log2(leaves)
        } catch (const ShapeError& e) { threw = std::string(e.what()).find("cap_height") != std::string::npos; }
        REQUIRE(threw);
        threw = false;
        try {
            PolynomialBatch::from_values(ctx, {std::vector<F>(8, 1), std::vector<F>(16, 1)}, 1, false, 0);
        } catch (const ShapeError&) { threw = true; }  // "Polynomial degrees inconsistent"
        REQUIRE(threw);
    }
    // ---- prove_openings: byte-identical FriProof, accepted by the restated verifier
    FriConfig cfg;
    cfg.rate_bits = r; cfg.cap_height = h; cfg.proof_of_work_bits = 7; cfg.num_query_rounds = 9;
    cfg.reduction_strategy.kind = FriReductionStrategy::Fixed;
    cfg.reduction_strategy.fixed = {3, 2};
    FriParams params = cfg.fri_params(log_n, false);
    REQUIRE(params.final_poly_len() == (n >> 5));
    Ext zeta{rnd(), rnd()};
    Ext gzeta = ext_mul(zeta, Ext{primitive_root_of_unity(log_n), 0});
    FriInstanceInfo inst;
    inst.oracles = {{Bs[0], false}, {Bs[1], false}};
    FriBatchInfo b0{zeta, {}}, b1{gzeta, {{1, 0}, {1, 2}}};
    for (uint32_t o = 0; o < 2; o++)
        for (uint32_t k = 0; k < Bs[o]; k++) b0.polynomials.push_back({o, k});
    inst.batches = {b0, b1};

    Challenger ch;
    glo_challenger* och = glo_challenger_new();
    for (int o = 0; o < 2; o++) {
        MerkleCap cap = batches[o].cap();
        ch.observe_cap(cap);
        auto flat = cap.flatten();
        glo_challenger_observe(och, flat.data(), flat.size());
    }
    glo_challenger* och_verify = glo_challenger_clone(och);
    FriProof proof = PolynomialBatch::prove_openings(inst, {&batches[0], &batches[1]}, ch, params);
    std::vector<uint8_t> bytes = proof.to_bytes();

    glo_fri_params op{};
    op.rate_bits = r; op.cap_height = h; op.proof_of_work_bits = 7; op.num_query_rounds = 9; op.num_reductions = 2;
    op.reduction_arity_bits[0] = 3; op.reduction_arity_bits[1] = 2;
    std::vector<std::vector<uint32_t>> oi(2), pi(2);
    glo_fri_batch ob[2];
    for (int b = 0; b < 2; b++) {
        for (auto& p : inst.batches[b].polynomials) { oi[b].push_back(p.oracle_index); pi[b].push_back(p.polynomial_index); }
        ob[b].point[0] = inst.batches[b].point.c0; ob[b].point[1] = inst.batches[b].point.c1;
        ob[b].num_polys = oi[b].size(); ob[b].oracle_index = oi[b].data(); ob[b].poly_index = pi[b].data();
    }
    uint8_t* obytes = nullptr;
    size_t olen = 0;
    REQUIRE(glo_prove_openings(ocommits.data(), 2, ob, 2, och, &op, &obytes, &olen, nullptr, nullptr, nullptr, nullptr) == 0);
    REQUIRE(olen == bytes.size() && std::memcmp(obytes, bytes.data(), olen) == 0);
    REQUIRE(ch.get_challenge() == glo_challenger_get_challenge(och));  // transcripts stay in sync
    // openings via eval_commitment (OpeningSet::new) feed the verifier
    std::vector<F> opened;
    for (int b = 0; b < 2; b++)
        for (auto& p : inst.batches[b].polynomials) {
            auto ev = batches[p.oracle_index].eval_commitment(inst.batches[b].point);
            F want[2], z[2] = {inst.batches[b].point.c0, inst.batches[b].point.c1};
            glo_eval_poly_base_at_ext(glo_commit_coeffs(ocommits[p.oracle_index]) + p.polynomial_index * n, n, z, want);
            REQUIRE(ev[p.polynomial_index].c0 == want[0] && ev[p.polynomial_index].c1 == want[1]);
            opened.push_back(want[0]); opened.push_back(want[1]);
        }
    const uint64_t* caps[2] = {glo_commit_cap(ocommits[0]), glo_commit_cap(ocommits[1])};
    const size_t widths[2] = {Bs[0], Bs[1]};
    REQUIRE(glo_verify_fri_proof(caps, Bs, widths, 2, ob, 2, opened.data(), log_n, och_verify, &op, bytes.data(), bytes.size()) == 0);
    glo_free(obytes);
    printf("CPP HOST PARITY OK (%zu proof bytes, %llu kernels launched)\n", bytes.size(), (unsigned long long)ctx.launch_count());
    return 0;
}
