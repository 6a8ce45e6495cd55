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
