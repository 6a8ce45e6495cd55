"""Multi-GPU plumbing for the row-block sharded commitment: one process per GPU,
torch.distributed (NCCL on GPUs, gloo in CPU tests) for the single exchange the path needs -- an
all-gather of the shards' Merkle-cap entries (2^cap_height x 32 bytes in total).

No LDE data ever crosses NVLink: shard g of G evaluates every column on its own coset
(g_shift * w_N^{bitrev(g)}) <w_{N/G}>, hashes its leaves and reduces its own cap subtrees."""
import numpy as np

from .hash import MerkleCap


def shard_row_range(lde_size, shard_index, num_shards):
    """Leaf rows [begin, end) of the single-device tree that shard `shard_index` owns."""
    assert lde_size % num_shards == 0
    rows = lde_size // num_shards
    return shard_index * rows, (shard_index + 1) * rows


def shard_cap_range(cap_height, shard_index, num_shards):
    """Cap entries [begin, end) owned by the shard (whole cap subtrees: num_shards <= 2^cap_height)."""
    c = 1 << cap_height
    if num_shards > c:
        raise ValueError("num_shards=%d exceeds the cap size %d" % (num_shards, c))
    per = c // num_shards
    return shard_index * per, (shard_index + 1) * per


def owner_of_leaf(leaf_index, lde_size, num_shards):
    """(shard, local leaf index) holding leaf `leaf_index` of the single-device tree."""
    rows = lde_size // num_shards
    return leaf_index // rows, leaf_index % rows


def gather_cap(local_cap, group=None, device=None):
    """All-gather the shards' cap entries into the full MerkleCap (identical on every rank and equal to
    the single-device cap). `local_cap`: (C/G, 4) uint64 array or MerkleCap."""
    import torch
    import torch.distributed as dist

    hashes = local_cap.hashes if isinstance(local_cap, MerkleCap) else np.asarray(local_cap, dtype=np.uint64)
    hashes = np.ascontiguousarray(hashes, dtype=np.uint64).reshape(-1, 4)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return MerkleCap(hashes.copy())
    world = dist.get_world_size(group)
    t = torch.from_numpy(hashes.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * t.shape[0], 4), dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    full = out.cpu().numpy().view(np.uint64).reshape(-1, 4)
    return MerkleCap(full.copy())


def open_sharded(batch, leaf_indices, group=None):
    """MerkleTree::get + prove (merkle_tree.rs:226-237) for GLOBAL leaf indices of a row-block sharded
    PolynomialBatch: every rank opens the indices it owns on its own GPU (local index, local cap subtree --
    the sibling path is the same as in the single-device tree) and the ranks all-gather the results.
    Collective: every rank must call it with the same indices. Returns (leaves (q, W), paths (q, L, 4))."""
    import torch.distributed as dist

    idx = [int(i) for i in leaf_indices]
    G = batch.num_shards
    mine = [(k, owner_of_leaf(i, batch.lde_size, G)[1]) for k, i in enumerate(idx)
            if owner_of_leaf(i, batch.lde_size, G)[0] == batch.shard_index]
    lv, pt = batch.merkle_tree.open_many([loc for _, loc in mine])
    part = [(k, lv[j], pt[j]) for j, (k, _) in enumerate(mine)]
    if G == 1 or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        parts = [part]
    else:
        parts = [None] * dist.get_world_size(group)
        dist.all_gather_object(parts, part, group=group)
    layers = batch.degree_log + batch.rate_bits - batch.cap_height
    leaves = np.empty((len(idx), batch.leaf_width), dtype=np.uint64)
    paths = np.empty((len(idx), layers, 4), dtype=np.uint64)
    seen = 0
    for p in parts:
        for k, l, q in p:
            leaves[k], paths[k] = l, q
            seen += 1
    if seen != len(idx):
        raise RuntimeError("sharded opening: %d of %d indices were served" % (seen, len(idx)))
    return leaves, paths


def prove_openings_sharded(instance, oracles, challenger, fri_params, group=None):
    """prove_openings (oracle.rs:176-237) when the initial oracles are row-block sharded over the ranks.
    Coefficients are replicated (every rank ran the iNTT), so every rank runs the (small, single-column) FRI
    commit phase and the transcript redundantly and deterministically; only the initial-tree openings cross
    ranks. The returned FriProof is identical on every rank and byte-identical to the single-device proof.
    The caller must already have observed the FULL caps (gather_cap) in `challenger`."""
    from . import fri as F

    alpha = challenger.get_extension_challenge()
    state = F._begin(instance, oracles, alpha, fri_params)
    try:
        caps, final_coeffs = F.fri_committed_trees(state, challenger, fri_params)
        pow_witness = F.fri_proof_of_work(challenger, fri_params.config, state.ctx)
        n = fri_params.lde_size()

        class _Routed:  # quacks like PolynomialBatch for fri_prover_query_rounds
            def __init__(self, b):
                self.merkle_tree = self
                self._b = b

            def open_many(self, indices):
                return open_sharded(self._b, indices, group)

        rounds, _ = F.fri_prover_query_rounds([_Routed(o) for o in oracles], state, challenger, n, fri_params)
        return F.FriProof(caps, rounds, final_coeffs, pow_witness)
    finally:
        state.close()


def column_slice(num_polys, rank, world):
    """Columns [b0, b1) whose iNTT rank `rank` computes; slices are equal-sized (the last ones may be padded)."""
    per = (num_polys + world - 1) // world
    b0 = min(rank * per, num_polys)
    return b0, min(b0 + per, num_polys), per


class ColumnShardedCommitter:
    """from_values over G ranks with BOTH axes of SURVEY.md section 8e: each rank uploads and inverse-transforms
    only its own slice of the columns (the reference's rayon axis, oracle.rs:65-69), the ranks all-gather the
    coefficients over NVLink (NCCL; the optional second collective of section 8e), then every rank extends all
    columns on its own row block / coset and hashes its own leaves (gl_commit_create_sharded, is_coeffs=1).
    Compared with replicating the iNTT this cuts the per-rank H2D and iNTT work by G; the commitment is bit-identical.

    Buffers (device slice, gathered coefficients) are torch tensors allocated once and reused across calls."""

    def __init__(self, ctx, num_polys, log_n, rate_bits, cap_height, rank, world, device, group=None):
        import torch

        self.ctx, self.B, self.log_n, self.r, self.h = ctx, num_polys, log_n, rate_bits, cap_height
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.n = 1 << log_n
        self.b0, self.b1, self.per = column_slice(num_polys, rank, world)
        self.slice = torch.zeros((self.per, self.n), dtype=torch.int64, device=device)   # padded columns stay 0
        self.coeffs = torch.empty((self.per * world, self.n), dtype=torch.int64, device=device)

    def commit(self, values, from_host):
        """values: torch int64 tensor holding THIS RANK'S columns [b0, b1) x n (pinned host if from_host, else on
        the device). Returns the gl_commit handle (row-block shard `rank` of `world`)."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _native as N

        L = N.lib()
        cnt = self.b1 - self.b0
        if cnt:
            self.slice[:cnt].copy_(values[:cnt], non_blocking=True)   # H2D of 1/G of the trace (or D2D)
            N.check(L.gl_ntt(self.ctx.h, C.c_void_p(self.slice.data_ptr()), self.log_n, cnt, self.n, 1, 0, 1,
                             N.MEM_DEVICE), self.ctx.h)
        if self.world > 1:
            dist.all_gather_into_tensor(self.coeffs, self.slice, group=self.group)
            src = self.coeffs
        else:
            src = self.slice
        h = N.vp()
        N.check(L.gl_commit_create_sharded(self.ctx.h, C.c_void_p(src.data_ptr()), self.n, self.B, self.log_n, self.r,
                                           self.h, None, 1, N.MEM_DEVICE, self.rank, self.world, C.byref(h)), self.ctx.h)
        return h
