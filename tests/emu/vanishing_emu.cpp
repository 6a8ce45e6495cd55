// Host run of the device evaluator of plonky2's vanishing polynomial (plonky2_b200/csrc/gl_vanishing.cuh): the same
// vp_eval_point the kernel k_plonk_quotient calls per thread, with threads as a loop and host arrays in place of device
// memory. Test infrastructure: built as a shared library and driven by tests/test_plonk_quotient.py, which compares the
// result with the oracle's restatement of compute_quotient_polys.
#include <vector>
#include "../../plonky2_b200/csrc/gl_vanishing.cuh"
using namespace gl;

// quotient VALUES (before the coset iFFT): out = n_alphas columns of size = 2^(degree_bits + qd_bits) words.
// Parameters are set up the way gl_plonk_quotient (plonky2_b200.cu) does. Returns 1 if the program divided by zero.
extern "C" int emu_plonk_quotient_values(const uint64_t* const* lde, const size_t* lde_stride, uint32_t n_commits,
                                         uint32_t rate_bits, uint32_t degree_bits, uint32_t qd_bits,
                                         const gl_vp_instr* prog, uint32_t n_instr, const uint64_t* consts,
                                         const uint64_t* alphas, uint32_t n_alphas, uint32_t n_terms, uint64_t* out) {
    const uint32_t size_log = degree_bits + qd_bits;
    const size_t size = (size_t)1 << size_log;
    std::vector<uint64_t> apow((size_t)n_alphas * n_terms);
    for (uint32_t a = 0; a < n_alphas; a++) {
        uint64_t pw = 1;
        for (uint32_t t = 0; t < n_terms; t++, pw = mul(pw, alphas[a])) apow[(size_t)a * n_terms + t] = canon(pw);
    }
    const uint64_t ws = root_of_unity(size_log);
    const size_t tcnt = 4096 > (size >> 12) + 1 ? 4096 : (size >> 12) + 1;
    std::vector<uint64_t> xhi(tcnt), xlo(tcnt);
    {
        const uint64_t whi = gl::pow(ws, 4096);
        uint64_t a = 1, b = 1;
        for (size_t k = 0; k < tcnt; k++, a = mul(a, whi), b = mul(b, ws)) {
            xhi[k] = canon(a);
            xlo[k] = canon(b);
        }
    }
    VanishingParams p;
    for (uint32_t c = 0; c < GL_VP_MAX_COMMITS; c++) {
        p.lde[c] = c < n_commits ? lde[c] : nullptr;
        p.lde_stride[c] = c < n_commits ? lde_stride[c] : 0;
    }
    p.log_N = degree_bits + rate_bits;
    p.degree_bits = degree_bits;
    p.qd_bits = qd_bits;
    p.prog = prog;
    p.n_instr = n_instr;
    p.consts = consts;
    p.apow = apow.data();
    p.n_alphas = n_alphas;
    p.n_terms = n_terms;
    p.xhi = xhi.data();
    p.xlo = xlo.data();
    p.shift = MULTIPLICATIVE_GROUP_GENERATOR;
    p.n_field = canon((uint64_t)1 << degree_bits);
    uint64_t g_pow_n = MULTIPLICATIVE_GROUP_GENERATOR;
    for (uint32_t k = 0; k < degree_bits; k++) g_pow_n = sqr(g_pow_n);
    const uint64_t wq = root_of_unity(qd_bits);
    uint64_t xq = 1;
    for (uint32_t j = 0; j < GL_VP_MAX_QD; j++) p.zh[j] = p.zh_inv[j] = 0;
    for (uint32_t j = 0; j < (1u << qd_bits); j++, xq = mul(xq, wq)) {
        p.zh[j] = canon(sub(mul(g_pow_n, xq), 1));
        p.zh_inv[j] = canon(gl::inv(p.zh[j]));
    }
    p.out = out;
    p.flag = nullptr;
    int bad = 0;
    for (size_t j = 0; j < size; j++) {  // one "thread" per leaf row
        uint64_t regs[GL_VP_MAX_REGS];
        for (int k = 0; k < GL_VP_MAX_REGS; k++) regs[k] = 0xDEADBEEFDEADBEEFull;  // uninitialised on the device
        if (!vp_eval_point(p, j, regs)) bad = 1;
    }
    return bad;
}
