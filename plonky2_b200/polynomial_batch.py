"""PolynomialBatch mirroring plonky2/src/fri/oracle.rs:30-237: a batch of polynomials committed with a
Poseidon Merkle tree over its coset LDE. All bulk data stays on the GPU behind a gl_commit handle; the
reference's public fields (`polynomials`, `merkle_tree.{leaves,digests,cap}`) are read back on demand."""
import ctypes as C

import numpy as np

from . import _native as N
from .field import log2_strict
from .hash import MerkleCap, MerkleProof

SALT_SIZE = 4  # oracle.rs:26


def random_field_elements(count):
    """Uniform canonical field elements from the OS CSPRNG (F::rand, field/src/goldilocks_field.rs:61-64):
    64-bit draws, rejecting values >= p -- vectorised (a 2^23 x 4 salt takes ~0.3 s, not minutes)."""
    import os

    from .field import ORDER

    out = np.empty(count, dtype=np.uint64)
    filled = 0
    while filled < count:
        need = count - filled
        cand = np.frombuffer(os.urandom(8 * (need + need // 1024 + 16)), dtype="<u8")
        cand = cand[cand < np.uint64(ORDER)][:need]
        out[filled:filled + len(cand)] = cand
        filled += len(cand)
    return out


class _DeviceMerkleTree:
    """View of PolynomialBatch.merkle_tree (merkle_tree.rs:46-62) living on the device."""

    def __init__(self, batch):
        self._b = batch

    @property
    def cap(self):
        """The Merkle cap (for a sharded batch: this shard's cap entries; see distributed.gather_cap)."""
        b = self._b
        out = np.empty(((1 << b.cap_height) // b.num_shards, 4), dtype=np.uint64)
        N.check(N.lib().gl_commit_cap(b.h, N.np_ptr(out), N.MEM_HOST), b.ctx.h)
        return MerkleCap(out)

    @property
    def leaves(self):
        return self.get_rows(0, self._b.local_rows)

    def get_rows(self, begin, count):
        b = self._b
        out = np.empty((count, b.leaf_width), dtype=np.uint64)
        if count:
            N.check(N.lib().gl_commit_leaves(b.h, begin, count, N.np_ptr(out), N.MEM_HOST), b.ctx.h)
        return out

    @property
    def digests(self):
        b = self._b
        out = np.empty((2 * (b.local_rows - (1 << b.cap_height) // b.num_shards), 4), dtype=np.uint64)
        if out.size:
            N.check(N.lib().gl_commit_digests(b.h, N.np_ptr(out), N.MEM_HOST), b.ctx.h)
        return out

    def get(self, i):
        return self.get_rows(i, 1)[0]

    def open_many(self, indices):
        b = self._b
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        layers = b.degree_log + b.rate_bits - b.cap_height  # local rows and local cap shrink together
        leaves = np.empty((len(idx), b.leaf_width), dtype=np.uint64)
        paths = np.empty((len(idx), layers, 4), dtype=np.uint64)
        if len(idx):
            N.check(N.lib().gl_commit_open(b.h, N.np_ptr(idx), len(idx), N.np_ptr(leaves),
                                           N.np_ptr(paths) if paths.size else None), b.ctx.h)
        return leaves, paths

    def prove(self, leaf_index):
        return MerkleProof(self.open_many([leaf_index])[1][0])


class PolynomialBatch:
    """PolynomialBatch<F, PoseidonGoldilocksConfig, 2> (oracle.rs:30-37)."""

    def __init__(self, handle, ctx, num_polys, degree_log, rate_bits, cap_height, blinding, shard=(0, 1)):
        self.h, self.ctx = handle, ctx
        self.shard_index, self.num_shards = shard
        self.num_polys, self.degree_log, self.rate_bits = num_polys, degree_log, rate_bits
        self.cap_height, self.blinding = cap_height, blinding
        self.leaf_width = num_polys + (SALT_SIZE if blinding else 0)
        self.lde_size = 1 << (degree_log + rate_bits)
        self.local_rows = self.lde_size // self.num_shards  # leaf rows [shard*local_rows, (shard+1)*local_rows)
        self.merkle_tree = _DeviceMerkleTree(self)

    @classmethod
    def _create(cls, cols, rate_bits, blinding, cap_height, is_coeffs, salt, ctx, shard=(0, 1)):
        ctx = ctx or N.default_context()
        cols = np.ascontiguousarray(cols, dtype=np.uint64)
        if cols.ndim != 2 or cols.shape[0] == 0:
            raise N.ShapeError("expected a non-empty (num_polys, degree) array")
        B, n = cols.shape
        log_n = log2_strict(n)
        sp = None
        if blinding:
            if salt is None:
                # the reference draws the salt from OsRng (oracle.rs:133-137); same source here
                salt = random_field_elements(SALT_SIZE * (n << rate_bits)).reshape(SALT_SIZE, -1)
            salt = np.ascontiguousarray(salt, dtype=np.uint64)
            if salt.shape != (SALT_SIZE, n << rate_bits):
                raise N.ShapeError("salt must be (4, n << rate_bits)")
            sp = N.np_ptr(salt)
        h = N.vp()
        N.check(N.lib().gl_commit_create_sharded(ctx.h, N.np_ptr(cols), n, B, log_n, rate_bits, cap_height, sp,
                                                 int(is_coeffs), N.MEM_HOST, int(shard[0]), int(shard[1]),
                                                 C.byref(h)), ctx.h)
        return cls(h, ctx, B, log_n, rate_bits, cap_height, bool(blinding), (int(shard[0]), int(shard[1])))

    @classmethod
    def from_values(cls, values, rate_bits, blinding, cap_height, timing=None, fft_root_table=None, *,
                    salt=None, ctx=None, shard=(0, 1)):
        """from_values (oracle.rs:57-79). `timing`/`fft_root_table` are accepted for signature parity.
        shard=(g, G): build only leaf rows [g*N/G, (g+1)*N/G) on this device (multi-GPU row-block sharding)."""
        return cls._create(values, rate_bits, blinding, cap_height, False, salt, ctx, shard)

    @classmethod
    def from_coeffs(cls, polynomials, rate_bits, blinding, cap_height, timing=None, fft_root_table=None, *,
                    salt=None, ctx=None, shard=(0, 1)):
        """from_coeffs (oracle.rs:82-112)."""
        return cls._create(polynomials, rate_bits, blinding, cap_height, True, salt, ctx, shard)

    @property
    def polynomials(self):
        out = np.empty((self.num_polys, 1 << self.degree_log), dtype=np.uint64)
        N.check(N.lib().gl_commit_coeffs(self.h, N.np_ptr(out), N.MEM_HOST), self.ctx.h)
        return out

    def get_lde_values(self, index, step):
        """get_lde_values (oracle.rs:142-147)."""
        out = np.empty(self.num_polys, dtype=np.uint64)
        N.check(N.lib().gl_commit_get_lde_values(self.h, index, step, N.np_ptr(out)), self.ctx.h)
        return out

    def eval_commitment(self, z):
        """eval_commitment of OpeningSet::new (plonk/proof.rs:313-351): every polynomial at z in F_{p^2};
        returns (num_polys, 2)."""
        pt = np.array([int(z[0]), int(z[1])], dtype=np.uint64)
        out = np.empty((self.num_polys, 2), dtype=np.uint64)
        N.check(N.lib().gl_commit_eval_ext(self.h, N.np_ptr(pt), N.np_ptr(out)), self.ctx.h)
        return out

    @staticmethod
    def prove_openings(instance, oracles, challenger, fri_params, final_poly_coeff_len=None,
                       max_num_query_steps=None, timing=None):
        """prove_openings (oracle.rs:176-237)."""
        from .fri import prove_openings

        return prove_openings(instance, oracles, challenger, fri_params, final_poly_coeff_len,
                              max_num_query_steps)

    def close(self):
        if getattr(self, "h", None):
            N.lib().gl_commit_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
