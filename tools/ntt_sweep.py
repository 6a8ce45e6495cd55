"""Sweep the two-pass split and the group size of the 2^20 NTT / LDE (run under gpurun)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from plonky2_b200 import _native as N
L = N.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = N.Context(0, stream=stream.cuda_stream)
log_n, cols = 20, 64
n = 1 << log_n
with torch.cuda.stream(stream):
    buf = torch.randint(0, 2**63 - 1, (cols, n), dtype=torch.int64, device=dev)
    for b in (0, 8, 9, 11, 12):
        for grp in (32, 64):
            ctx.set_ntt_group(grp)
            for _ in range(3):
                N.check(L.gl_ntt(ctx.h, C.c_void_p(buf.data_ptr()), log_n, cols, n, 0, 0, 1, N.MEM_DEVICE), ctx.h)
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(10):
                N.check(L.gl_ntt(ctx.h, C.c_void_p(buf.data_ptr()), log_n, cols, n, 0, 0, 1, N.MEM_DEVICE), ctx.h)
            e.record(stream); torch.cuda.synchronize()
            ms = a.elapsed_time(e) / 10
            print("b=%2d group=%2d : %.3f ms  (%.1f GB/s alg)" % (b, grp, ms, 16.0 * n * cols / ms / 1e6))
