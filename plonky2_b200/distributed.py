"""Multi-GPU plumbing for the row-block sharded commitment: one process per GPU,
torch.distributed (NCCL on GPUs, gloo in CPU tests) for the exchange the path needs -- an all-gather of the shards'
Merkle-cap entries (2^cap_height x 32 bytes in total) -- and PipelinedCommitter for the optional second axis
(column-sharded iNTT whose stores are the coefficient all-gather over NVLink).

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


def chunk_layout(num_polys, world, chunk_cols=64):
    """Column layout of the pipelined multi-GPU commitment: K chunks of Wc = pc*world consecutive columns; inside
    chunk c rank r transforms columns [c*Wc + r*pc, c*Wc + (r+1)*pc) (clipped to num_polys). Returns (pc, Wc, K)."""
    pc = max(1, chunk_cols // world)
    wc = pc * world
    return pc, wc, (num_polys + wc - 1) // wc


def chunk_columns(num_polys, rank, world, chunk, chunk_cols=64):
    """(first global column, count) of rank `rank`'s sub-block of chunk `chunk` (count may be 0 in the last chunk)."""
    pc, wc, _ = chunk_layout(num_polys, world, chunk_cols)
    b0 = min(chunk * wc + rank * pc, num_polys)
    return b0, min(b0 + pc, num_polys) - b0


class PipelinedCommitter:
    """from_values over G ranks with BOTH axes of SURVEY.md section 8e, in 64-column chunks:

      copy stream : H2D of my 64/G columns of each chunk (host input), issued ahead
      main stream : iNTT(0), iNTT(1), LDE(0), iNTT(2), LDE(1), ... , LDE(K-1), leaf hash, cap subtrees
                    iNTT(c)  = column-sharded inverse transform of MY columns of chunk c into my copy of the matrix
                    LDE(c)   = coset LDE of all 64 columns of chunk c on this rank's row block (gl_commit_add_columns)
      side stream : after iNTT(c): my coefficients -> EVERY rank's matrix over NVLink with 128-byte line stores to the
                    NVSwitch multicast address (gl_bcast; one store per peer mapping without multicast), then a
                    device-side barrier (symmetric-memory signal pads). LDE(c) waits for it; the transfer runs under
                    iNTT(c+1) / LDE(c-1) on a handful of SMs.

    The iNTT (the reference's rayon axis, oracle.rs:65-69) and the H2D are divided by G. Transports:
      "multimem" / "p2p"  as above (torch symmetric memory: CUDA IPC / fabric handles);
      "fused"             the iNTT's last pass stores straight to the multicast address (gl_ntt_bcast): no second
                          kernel, but its transposing stores are 64-byte segments and NVLink runs them at ~120 GB/s
                          (measured, profiles/r02) -- kept for comparison;
      "nccl"              ncclAllGather per chunk on the main stream (fallback when peer mappings are unavailable).
    The commitment is bit-identical to the single-device one.

    Stream contract (checked): `ctx` must have been created on the torch stream that is current when commit() is
    called. The returned handle BORROWS the committer's coefficient matrix: it is valid until the next commit()."""

    def __init__(self, ctx, num_polys, log_n, rate_bits, cap_height, rank, world, device, group=None,
                 transport="auto", chunk_cols=None, copy_ctas=32):
        import torch
        import torch.distributed as dist

        from . import _native as N

        self.ctx, self.B, self.log_n, self.r, self.h = ctx, num_polys, log_n, rate_bits, cap_height
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.n = 1 << log_n
        if chunk_cols is None:  # 64-column chunks, but at least ~4 chunks so that transfers have an LDE to hide under
            per = -(-num_polys // 4)
            chunk_cols = max(world, min(64, -(-per // world) * world))
        self.chunk_cols, self.copy_ctas = chunk_cols, copy_ctas
        self.pc, self.wc, self.K = chunk_layout(num_polys, world, chunk_cols)
        self.copy = torch.cuda.Stream(device=device)
        self.side = torch.cuda.Stream(device=device)
        self.ctx_side = N.Context(device.index if hasattr(device, "index") else int(device), stream=self.side.cuda_stream)
        self.stage = torch.empty((self.K, self.pc, self.n), dtype=torch.int64, device=device)
        self.h2d_events = [torch.cuda.Event() for _ in range(self.K)]
        self.intt_events = [torch.cuda.Event() for _ in range(self.K)]
        self.bcast_events = [torch.cuda.Event() for _ in range(self.K)]
        self.ready = torch.cuda.Event()
        self.timing = False          # set True to collect the transfer spans per commit()
        self._spans = []
        rows = self.K * self.wc
        self.symm, self.transport, self.transport_note = None, "nccl", ""
        if world > 1 and transport != "nccl":
            try:
                import torch.distributed._symmetric_memory as symm_mem

                g = group if group is not None else dist.group.WORLD
                self.coeffs = symm_mem.empty((rows, self.n), dtype=torch.int64, device=device)
                self.symm = symm_mem.rendezvous(self.coeffs, g)
                mc = int(getattr(self.symm, "multicast_ptr", 0) or 0)
                if transport in ("multimem", "fused") and not mc:
                    raise RuntimeError("no multicast mapping on this system")
                self.transport = transport if transport in ("fused", "p2p") else ("multimem" if mc else "p2p")
                ptrs = [int(p) for p in self.symm.buffer_ptrs]
                self.local = ptrs[rank]
                # multicast: one store reaches every rank (mine included); p2p: one store per peer mapping
                self.dests = [mc] if self.transport in ("multimem", "fused") else [p for i, p in enumerate(ptrs) if i != rank]
                if len(self.dests) > 8:
                    raise RuntimeError("more than 8 peers")
            except Exception as e:  # no peer mappings here: same loop over ncclAllGather
                if transport != "auto":
                    raise
                self.symm, self.transport = None, "nccl"
                self.transport_note = "symmetric memory unavailable: %r" % (e,)
        if self.symm is None:
            self.coeffs = torch.empty((rows, self.n), dtype=torch.int64, device=device)

    def _check_stream(self):
        import torch

        cur = torch.cuda.current_stream(self.device)
        if self.ctx.stream != cur.cuda_stream:
            raise RuntimeError("PipelinedCommitter: the context's stream (0x%x) is not the current torch stream (0x%x); "
                               "create the Context on the torch stream you call commit() under" % (self.ctx.stream, cur.cuda_stream))
        return cur

    def my_columns(self):
        """[(first global column, count)] per chunk: the columns this rank uploads and inverse-transforms."""
        return [chunk_columns(self.B, self.rank, self.world, c, self.chunk_cols) for c in range(self.K)]

    def commit(self, values, from_host):
        """values: torch int64 tensor of ALL columns (B x n) -- pinned host memory if from_host (only this rank's
        sub-blocks are read and uploaded), else on the device. Returns the gl_commit handle (row-block shard)."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _native as N

        L = N.lib()
        main = self._check_stream()
        n, nbytes = self.n, self.n * 8
        base = self.coeffs.data_ptr()
        mine = self.my_columns()
        if from_host:
            with torch.cuda.stream(self.copy):
                self.copy.wait_stream(main)           # the staging slots' previous readers are queued on `main`
                for c, (b0, cnt) in enumerate(mine):
                    if cnt:
                        self.stage[c, :cnt].copy_(values[b0:b0 + cnt], non_blocking=True)  # H2D of 1/G of the chunk
                    self.h2d_events[c].record(self.copy)
        if self.symm is not None:
            self.symm.barrier(channel=0)              # every rank's LDEs of the previous commitment have read the matrix
            self.ready.record(main)
        h = N.vp()
        N.check(L.gl_commit_begin(self.ctx.h, self.B, self.log_n, self.r, self.h, 0, self.rank, self.world, N.vp(base),
                                  C.byref(h)), self.ctx.h)

        def transform(c):
            b0, cnt = mine[c]
            if from_host:
                main.wait_event(self.h2d_events[c])
            src = self.stage[c].data_ptr() if from_host else (values[b0:b0 + cnt].data_ptr() if cnt else 0)
            off = b0 * nbytes
            if self.transport == "fused":
                if cnt:
                    outs = (N.vp * 1)(N.vp(self.dests[0] + off))
                    N.check(L.gl_ntt_bcast(self.ctx.h, N.vp(src), n, self.log_n, cnt, 1, outs, 1, n), self.ctx.h)
                self.symm.barrier(channel=1)
            elif self.symm is not None:
                if cnt:  # out of place into MY copy of the matrix
                    outs = (N.vp * 1)(N.vp(self.local + off))
                    N.check(L.gl_ntt_bcast(self.ctx.h, N.vp(src), n, self.log_n, cnt, 1, outs, 1, n), self.ctx.h)
                self.intt_events[c].record(main)
                with torch.cuda.stream(self.side):
                    if c == 0:
                        self.side.wait_event(self.ready)
                    self.side.wait_event(self.intt_events[c])
                    if self.timing:
                        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        t0.record(self.side)
                    if cnt:
                        outs = (N.vp * len(self.dests))(*[N.vp(d + off) for d in self.dests])
                        N.check(L.gl_bcast(self.ctx_side.h, N.vp(self.local + off), cnt * n, outs, len(self.dests),
                                           self.copy_ctas), self.ctx_side.h)
                    self.symm.barrier(channel=1)
                    if self.timing:
                        t1.record(self.side)
                        self._spans.append((t0, t1))
                    self.bcast_events[c].record(self.side)
            else:
                if cnt:
                    if not from_host:
                        self.stage[c, :cnt].copy_(values[b0:b0 + cnt], non_blocking=True)
                    N.check(L.gl_ntt(self.ctx.h, N.vp(self.stage[c].data_ptr()), self.log_n, cnt, n, 1, 0, 1, N.MEM_DEVICE),
                            self.ctx.h)
                dist.all_gather_into_tensor(self.coeffs[c * self.wc:(c + 1) * self.wc], self.stage[c], group=self.group)

        def extend(c):
            if self.symm is not None and self.transport != "fused":
                main.wait_event(self.bcast_events[c])
            c0 = c * self.wc
            N.check(L.gl_commit_add_columns(h, c0, min(self.wc, self.B - c0), N.vp(base + c0 * nbytes), n,
                                            N.COLS_COEFFS_CANONICAL, N.MEM_DEVICE), self.ctx.h)

        try:
            for c in range(self.K):
                transform(c)
                if c:
                    extend(c - 1)
            extend(self.K - 1)
            N.check(L.gl_commit_finish(h, None, N.MEM_DEVICE), self.ctx.h)
        except Exception:
            L.gl_commit_destroy(h)
            raise
        return h

    def transfer_ms(self, reset=True):
        """(accumulated ms, commits) of the side-stream [NVLink copy + barrier] spans since the last reset; they run
        under the main stream's transforms except for the last chunk's."""
        import torch

        torch.cuda.synchronize(self.device)
        ms = sum(a.elapsed_time(b) for a, b in self._spans)
        cnt = len(self._spans) // max(1, self.K)
        if reset:
            self._spans = []
        return ms, cnt


ColumnShardedCommitter = PipelinedCommitter  # round-1 name
