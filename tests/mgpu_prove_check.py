"""Multi-GPU end-to-end check (run under torchrun, one rank per GPU):
sharded commitments -> NCCL all-gather of caps -> prove_openings with routed initial-tree openings.
Rank 0 compares caps and the proof bytes with the CPU oracle. Launched by tests/test_gpu_parity.py when the
box has >= 2 GPUs, or by hand:  gpurun --gpus 2 -- python -m torch.distributed.run --nproc-per-node 2 ... this file
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

import plonky2_b200 as pb
from conftest import synth
from plonky2_b200 import distributed as D


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pb.default_context(local)
    log_n, r, h = 10, 3, 4
    Bs = [7, 5, 3]
    vals = [synth(0x60 + i, (B, 1 << log_n)) for i, B in enumerate(Bs)]
    commits = [pb.PolynomialBatch.from_values(v, r, False, h, ctx=ctx, shard=(rank, world)) for v in vals]
    dev = torch.device("cuda", local)
    caps = [D.gather_cap(c.merkle_tree.cap, device=dev) for c in commits]
    ch = pb.Challenger()
    for cap in caps:
        ch.observe_cap(cap)
    zeta = (1234567, 7654321)
    gz = pb.field.ext_mul(zeta, (pb.field.primitive_root_of_unity(log_n), 0))
    allp = [pb.FriPolynomialInfo(o, i) for o, B in enumerate(Bs) for i in range(B)]
    inst = pb.FriInstanceInfo([pb.FriOracleInfo(B, False) for B in Bs],
                              [pb.FriBatchInfo(zeta, allp), pb.FriBatchInfo(gz, [pb.FriPolynomialInfo(2, 0)])])
    params = pb.FriParams(pb.FriConfig(r, h, 8, ("Fixed", [4, 2]), 12), False, log_n, [4, 2])
    proof = D.prove_openings_sharded(inst, commits, ch, params)
    ok = True
    # column-sharded iNTT whose stores are the coefficient all-gather + row-block sharded LDE/Merkle (PipelinedCommitter)
    import ctypes as C
    from plonky2_b200 import _native as N_
    Bc, lg = 70, 12  # two chunks of 64 columns, the second one partial
    vals_c = synth(0x77, (Bc, 1 << lg))
    # torch copies / NCCL and the library must share ONE stream: make the context on a torch stream
    tstream = torch.cuda.Stream(device=dev)
    ctx2 = pb.Context(local, stream=tstream.cuda_stream)
    with torch.cuda.stream(tstream):
        host = torch.from_numpy(vals_c.view(np.int64).copy()).pin_memory()
        caps_c, transports = [], []
        for transport in ("auto", "fused", "p2p", "nccl"):
            try:
                cm = D.PipelinedCommitter(ctx2, Bc, lg, 2, 3, rank, world, dev, transport=transport)
            except Exception as e:  # a transport this system lacks (no multicast / no peer mappings) is reported, not fatal
                transports.append("%s unavailable: %r" % (transport, e))
                continue
            transports.append(cm.transport + (" (%s)" % cm.transport_note if cm.transport_note else ""))
            for from_host in (True, False, True):  # back-to-back commitments reuse the matrix: exercises the hand-over barriers
                src = host if from_host else host.to(dev)
                hnd = cm.commit(src, from_host=from_host)
                lcap = np.empty(((1 << 3) // world, 4), dtype=np.uint64)
                N_.check(N_.lib().gl_commit_cap(hnd, N_.np_ptr(lcap), N_.MEM_HOST), ctx2.h)
                N_.lib().gl_commit_destroy(hnd)
                caps_c.append(lcap)
    torch.cuda.synchronize(dev)
    full_caps_c = [D.gather_cap(lc, device=dev) for lc in caps_c]
    if rank == 0:
        import oracle_lib

        ocommits = [oracle_lib.Commit(v, r, h) for v in vals]
        for cap, o in zip(caps, ocommits):
            ok &= bool(np.array_equal(cap.hashes, o.cap))
        och = oracle_lib.Challenger()
        for o in ocommits:
            och.observe_cap(o.cap)
        obatches = [(b.point, [(p.oracle_index, p.polynomial_index) for p in b.polynomials]) for b in inst.batches]
        oproof = oracle_lib.prove_openings(ocommits, obatches, och, oracle_lib.make_params(r, h, 8, 12, [4, 2]))
        ok &= proof.to_bytes() == oproof
        want = oracle_lib.Commit(vals_c, 2, 3).cap
        for fc in full_caps_c:
            ok &= bool(np.array_equal(fc.hashes, want))
        print("MGPU_PROVE_CHECK", "OK" if ok else "FAILED", "world", world, "transports", transports, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
