"""world_size-2 gloo test of the multi-GPU host logic (plonky2_b200/distributed.py) on CPU: each rank
owns one row block of the commitment (its leaves/cap come from the oracle here, since there is no GPU),
the ranks all-gather their cap entries, and every rank must end with the single-device cap."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import oracle_lib
    from conftest import synth
    from plonky2_b200 import distributed as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, log_n, r, h = 5, 6, 2, 3
        N = 1 << (log_n + r)
        vals = synth(0x55, (B, 1 << log_n))
        full = oracle_lib.Commit(vals, r, h, nthreads=1)
        lo, hi = D.shard_row_range(N, rank, world)
        clo, chi = D.shard_cap_range(h, rank, world)
        # this rank's shard: its own leaves reduced to its own cap entries
        _, local_cap = oracle_lib.merkle_build(full.leaves[lo:hi], h - int(np.log2(world)), nthreads=1)
        assert np.array_equal(local_cap, full.cap[clo:chi])
        cap = D.gather_cap(local_cap)
        ok = np.array_equal(cap.hashes, full.cap)
        owner = D.owner_of_leaf(N - 1, N, world)
        q.put((rank, bool(ok), owner))
    finally:
        dist.destroy_process_group()


def test_cap_all_gather_two_ranks():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res)
    assert all(r[2] == (1, 127) for r in res)


def test_shard_ranges():
    from plonky2_b200 import distributed as D

    assert D.shard_row_range(1 << 10, 3, 8) == (384, 512)
    assert D.shard_cap_range(4, 3, 8) == (6, 8)
    with pytest.raises(ValueError):
        D.shard_cap_range(2, 0, 8)
    assert D.owner_of_leaf(700, 1024, 4) == (2, 188)
    # single process: gather_cap is the identity
    cap = np.arange(16, dtype=np.uint64).reshape(4, 4)
    assert np.array_equal(D.gather_cap(cap).hashes, cap)


def test_open_sharded_routing_single_process():
    """Routing logic of open_sharded with a fake local batch (no GPU): owned indices are opened locally with
    the local index; an index owned by another shard makes the single-process call fail loudly."""
    from plonky2_b200 import distributed as D

    class FakeTree:
        def open_many(self, idx):
            idx = list(idx)
            return (np.array([[100 + i, 7] for i in idx], dtype=np.uint64).reshape(len(idx), 2),
                    np.zeros((len(idx), 3, 4), dtype=np.uint64))

    class FakeBatch:
        num_shards, shard_index, lde_size, leaf_width = 2, 1, 64, 2
        degree_log, rate_bits, cap_height = 5, 1, 3
        merkle_tree = FakeTree()

    lv, pt = D.open_sharded(FakeBatch(), [32, 63])   # both owned by shard 1 -> local 0 and 31
    assert lv[:, 0].tolist() == [100, 131] and pt.shape == (2, 3, 4)
    with pytest.raises(RuntimeError):
        D.open_sharded(FakeBatch(), [5])             # owned by shard 0, nobody serves it here


@pytest.mark.parametrize("num_polys,world", [(234, 8), (234, 4), (234, 2), (64, 8), (5, 4), (3, 8), (16, 1), (300, 2)])
def test_chunk_layout_tiles_the_coefficient_matrix(num_polys, world):
    """PipelinedCommitter's layout contract (column axis): chunk c = Wc consecutive global columns, rank r's
    sub-block = pc consecutive columns at c*Wc + r*pc, so that (a) the fused iNTT stores / the per-chunk all-gather of
    the ranks' (pc, n) blocks in rank order IS rows [c*Wc, (c+1)*Wc) of the coefficient matrix, (b) every column is
    transformed by exactly one rank, (c) gathered chunks are contiguous column ranges for gl_commit_add_columns."""
    from plonky2_b200 import distributed as D

    pc, wc, K = D.chunk_layout(num_polys, world)
    assert wc == pc * world and K * wc >= num_polys > (K - 1) * wc
    covered = []
    for c in range(K):
        for rank in range(world):
            b0, cnt = D.chunk_columns(num_polys, rank, world, c)
            assert 0 <= cnt <= pc and b0 == min(c * wc + rank * pc, num_polys) and b0 + cnt <= num_polys
            covered += list(range(b0, b0 + cnt))
    assert covered == list(range(num_polys))


def _pipeline_worker(rank, world, port, num_polys, n, out):
    """The committer's data movement with gloo standing in for NCCL: per chunk every rank contributes its (pc, n)
    block and the all-gather lands in rows [c*Wc, (c+1)*Wc) of the matrix."""
    import torch
    import torch.distributed as dist

    from plonky2_b200 import distributed as D

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        pc, wc, K = D.chunk_layout(num_polys, world)
        full = torch.arange(num_polys * n, dtype=torch.int64).reshape(num_polys, n) * 3 + 1  # "coefficients" of column b
        coeffs = torch.zeros((K * wc, n), dtype=torch.int64)
        for c in range(K):
            b0, cnt = D.chunk_columns(num_polys, rank, world, c)
            stage = torch.full((pc, n), -1, dtype=torch.int64)
            stage[:cnt] = full[b0:b0 + cnt]
            dist.all_gather_into_tensor(coeffs[c * wc:(c + 1) * wc], stage)
        out.put((rank, bool(torch.equal(coeffs[:num_polys], full))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_polys", [11, 70])
def test_pipelined_gather_layout_two_ranks_gloo(num_polys):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + num_polys
    procs = [ctx.Process(target=_pipeline_worker, args=(r, 2, port, num_polys, 8, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
