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
