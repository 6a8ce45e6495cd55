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


@pytest.mark.parametrize("num_polys,world", [(234, 8), (234, 4), (64, 8), (5, 4), (3, 8), (16, 1)])
def test_column_slices_tile_the_all_gathered_coefficient_buffer(num_polys, world):
    """ColumnShardedCommitter's layout contract (column axis): equal-sized (padded) slices, in rank order,
    so that all_gather_into_tensor of the per-rank (per, n) buffers IS the (world*per, n) coefficient matrix whose
    first num_polys rows are the columns in order."""
    from plonky2_b200 import distributed as D

    per_all, covered = None, []
    for rank in range(world):
        b0, b1, per = D.column_slice(num_polys, rank, world)
        per_all = per if per_all is None else per_all
        assert per == per_all and 0 <= b0 <= b1 <= num_polys and b1 - b0 <= per
        # a rank's real columns sit at the start of its padded slot: global row rank*per + k <-> column b0 + k
        assert b0 == min(rank * per, num_polys)
        covered += list(range(b0, b1))
    assert covered == list(range(num_polys))
    assert per_all * world >= num_polys
