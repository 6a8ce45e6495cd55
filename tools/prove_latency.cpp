// Wall-clock of one recursion-sized proof through the compiled host layer (include/plonky2_b200.hpp):
// the hot-path call sequence of BASELINE configs[3] (n = 2^14, standard_recursion_config): per proof
// from_values(135 wires) + from_values(20 Z/partial products) + from_coeffs(16 quotient chunks) +
// prove_openings over 4 oracles (84 constants/sigmas committed once), host transcript in the loop,
// host buffers in, proof bytes out.  Prints one JSON object.   usage: prove_latency [reps]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../include/plonky2_b200.hpp"
using namespace plonky2_b200;

static uint64_t st = 0x40;
static F rnd() {
    st += 0x9E3779B97F4A7C15ULL;
    uint64_t z = st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return (z ^ (z >> 31)) % ORDER;
}
int main(int argc, char** argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 7;
    const uint32_t log_n = 14, r = 3, h = 4;
    const size_t n = size_t(1) << log_n;
    const size_t Bs[4] = {84, 135, 20, 16};
    std::vector<std::vector<std::vector<F>>> data(4);
    for (int o = 0; o < 4; o++) {
        data[o].assign(Bs[o], std::vector<F>(n));
        for (auto& c : data[o])
            for (auto& x : c) x = rnd();
    }
    Context ctx(0);
    FriParams params = FriConfig::standard_recursion().fri_params(log_n, false);
    Ext zeta{rnd(), rnd()};
    Ext gz = ext_mul(zeta, Ext{primitive_root_of_unity(log_n), 0});
    FriInstanceInfo inst;
    FriBatchInfo b0{zeta, {}}, b1{gz, {{2, 0}, {2, 1}}};
    for (uint32_t o = 0; o < 4; o++)
        for (uint32_t k = 0; k < Bs[o]; k++) b0.polynomials.push_back({o, k});
    inst.batches = {b0, b1};
    PolynomialBatch constants = PolynomialBatch::from_values(ctx, data[0], r, false, h);
    size_t proof_len = 0;
    auto once = [&]() {
        Challenger ch;
        ch.observe_cap(constants.cap());
        PolynomialBatch wires = PolynomialBatch::from_values(ctx, data[1], r, false, h);
        ch.observe_cap(wires.cap());
        ch.get_n_challenges(4);
        PolynomialBatch zs = PolynomialBatch::from_values(ctx, data[2], r, false, h);
        ch.observe_cap(zs.cap());
        ch.get_n_challenges(2);
        PolynomialBatch quot = PolynomialBatch::from_coeffs(ctx, data[3], r, false, h);
        ch.observe_cap(quot.cap());
        ch.get_extension_challenge();
        FriProof p = PolynomialBatch::prove_openings(inst, {&constants, &wires, &zs, &quot}, ch, params);
        proof_len = p.to_bytes().size();
    };
    once();
    std::vector<double> ms;
    for (int i = 0; i < reps; i++) {
        auto t0 = std::chrono::steady_clock::now();
        once();
        ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    std::sort(ms.begin(), ms.end());
    printf("{\"cpp_host_ms_median\": %.3f, \"cpp_host_ms_min\": %.3f, \"proof_bytes\": %zu, \"reps\": %d}\n",
           ms[ms.size() / 2], ms[0], proof_len, reps);
    return 0;
}
