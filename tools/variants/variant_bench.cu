// Times every leaf-hash variant cubin given on the command line on the same input and checks the digests of
// the first leaves against the host formulation of the same header (gl_poseidon.cuh compiled for the CPU).
// Build: nvcc -O2 -std=c++17 -o variant_bench variant_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include "../../plonky2_b200/csrc/gl_poseidon.cuh"
#define CU(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); \
    fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, s_); exit(1); } } while (0)
int main(int argc, char** argv) {
    const size_t N = (size_t)1 << 19;
    const uint32_t W = argc > 1 ? (uint32_t)atoi(argv[1]) : 234;
    CU(cuInit(0));
    CUdevice dev; CU(cuDeviceGet(&dev, 0));
    CUcontext ctx; CU(cuDevicePrimaryCtxRetain(&ctx, dev)); CU(cuCtxSetCurrent(ctx));
    std::vector<uint64_t> h((size_t)N * W);
    uint64_t x = 0x9E3779B97F4A7C15ULL;
    for (auto& v : h) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = x; }  // arbitrary u64s (non-canonical too)
    CUdeviceptr d_in, d_out;
    CU(cuMemAlloc(&d_in, h.size() * 8)); CU(cuMemAlloc(&d_out, N * 32));
    CU(cuMemcpyHtoD(d_in, h.data(), h.size() * 8));
    const int NCHK = 512;
    std::vector<uint64_t> ref(NCHK * 4), got(NCHK * 4);
    for (int j = 0; j < NCHK; j++) gl::hash_or_noop_strided<true, false>(h.data() + (size_t)j, N, W, &ref[4 * j]);  // column-major input
    const gl::PoseidonTables& T = gl::host_poseidon_tables();
    CUevent e0, e1; CU(cuEventCreate(&e0, 0)); CU(cuEventCreate(&e1, 0));
    const double perms = (double)N * ((W + 7) / 8);
    for (int a = 2; a < argc; a++) {
        CUmodule mod; CU(cuModuleLoad(&mod, argv[a]));
        CUfunction f, ft; CU(cuModuleGetFunction(&f, mod, "k_leaf")); CU(cuModuleGetFunction(&ft, mod, "k_threads"));
        CUdeviceptr cp; size_t cs; CU(cuModuleGetGlobal(&cp, &cs, mod, "_ZN2gl5c_posE"));
        if (cs != sizeof(T)) { fprintf(stderr, "c_pos size mismatch\n"); return 1; }
        CU(cuMemcpyHtoD(cp, &T, sizeof(T)));
        CUdeviceptr d_t; CU(cuMemAlloc(&d_t, 4)); void* ta[] = {&d_t};
        CU(cuLaunchKernel(ft, 1, 1, 1, 1, 1, 1, 0, 0, ta, 0));
        int threads = 0; CU(cuMemcpyDtoH(&threads, d_t, 4)); CU(cuMemFree(d_t));
        int regs = 0, lmem = 0; cuFuncGetAttribute(&regs, CU_FUNC_ATTRIBUTE_NUM_REGS, f);
        cuFuncGetAttribute(&lmem, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, f);
        int occ = 0; cuOccupancyMaxActiveBlocksPerMultiprocessor(&occ, f, threads, 0);
        size_t n = N; uint32_t w = W; void* args[] = {&d_in, &n, &w, &d_out};
        unsigned grid = (unsigned)((N + threads - 1) / threads);
        CU(cuMemsetD8(d_out, 0, N * 32));
        float best = 1e30f;
        for (int it = 0; it < 4; it++) {
            CU(cuEventRecord(e0, 0));
            CU(cuLaunchKernel(f, grid, 1, 1, threads, 1, 1, 0, 0, args, 0));
            CU(cuEventRecord(e1, 0)); CU(cuEventSynchronize(e1));
            float ms; CU(cuEventElapsedTime(&ms, e0, e1));
            if (it && ms < best) best = ms;
        }
        CU(cuMemcpyDtoH(got.data(), d_out, NCHK * 32));
        bool ok = got == ref;
        printf("%-44s thr=%3d regs=%3d lmem=%3d cta/SM=%d  %8.3f ms  %8.1f Mperm/s  %s\n", argv[a], threads, regs, lmem, occ,
               best, perms / best / 1e3, ok ? "OK" : "MISMATCH");
        fflush(stdout);
        CU(cuModuleUnload(mod));
    }
    return 0;
}
