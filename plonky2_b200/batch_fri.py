"""Batch FRI (plonky2/src/batch_fri/oracle.rs:30-183, batch_fri/prover.rs:30-258): polynomial commitments over
polynomials of several degrees -- one BatchMerkleTree over the degree groups' LDE rows -- and one FRI proof for all of
them, the lower-degree instances being mixed into the folded codeword when it reaches their length. SURVEY 8(f) row 4.

Device path: every degree group is a PolynomialBatch (coefficients and LDE stay on the device); the batch tree's first
stage IS the tallest group's own Merkle tree built to the height of the next group, later stages hash `previous cap ||
group rows`; every instance's composed polynomial comes from gl_fri_begin, the rounds are gl_fri_commit_round /
gl_fri_fold plus gl_fri_mix at the mixing points."""
import ctypes as C

import numpy as np

from . import _native as N
from . import fri as F
from .field import log2_strict
from .hash import MerkleCap, MerkleProof, MerkleTree
from .polynomial_batch import PolynomialBatch


class BatchFriOracle:
    """BatchFriOracle<F, C, D> (batch_fri/oracle.rs:30-40)."""

    def __init__(self, groups, degree_bits, rate_bits, cap_height, group_of_poly, ctx):
        self.groups, self.degree_bits, self.rate_bits, self.cap_height = groups, degree_bits, rate_bits, cap_height
        self.group_of_poly, self.ctx = group_of_poly, ctx   # polynomial index -> (group, index inside the group)
        self.blinding = False
        heights = [d + rate_bits for d in degree_bits]
        self.leaf_heights = heights
        self.stages = []
        cap = None
        for k, g in enumerate(groups):
            nxt = heights[k + 1] if k + 1 < len(groups) else cap_height
            if k == 0:
                self.stages.append(g)       # the tallest group's own tree, built with cap height = next stage's height
                cap = g.merkle_tree.cap.hashes
            else:
                rows = g.merkle_tree.get_rows(0, 1 << heights[k])
                t = MerkleTree(np.ascontiguousarray(np.concatenate([cap, rows], axis=1)), nxt, ctx)
                self.stages.append(t)
                cap = t.cap.hashes
        self.cap = MerkleCap(cap)

    @classmethod
    def from_values(cls, values, rate_bits, blinding, cap_height, ctx=None):
        """from_values (oracle.rs:45-68): values = list of 1-D arrays, lengths non-increasing powers of two."""
        return cls._build(values, rate_bits, blinding, cap_height, False, ctx)

    @classmethod
    def from_coeffs(cls, polynomials, rate_bits, blinding, cap_height, ctx=None):
        """from_coeffs (oracle.rs:71-131)."""
        return cls._build(polynomials, rate_bits, blinding, cap_height, True, ctx)

    @classmethod
    def _build(cls, polys, rate_bits, blinding, cap_height, is_coeffs, ctx):
        if blinding:
            raise NotImplementedError("blinding batch oracles are not supported")
        ctx = ctx or N.default_context()
        polys = [np.ascontiguousarray(p, dtype=np.uint64).reshape(-1) for p in polys]
        bits = [log2_strict(len(p)) for p in polys]
        if any(a < b for a, b in zip(bits, bits[1:])):
            raise N.ShapeError("polynomials must be sorted by degree, largest first")   # oracle.rs:83
        groups, degree_bits, group_of_poly = [], [], []
        start = 0
        for i, d in enumerate(bits):
            if i == len(bits) - 1 or d > bits[i + 1]:
                cols = np.stack(polys[start:i + 1])
                nxt_bits = bits[i + 1] if i + 1 < len(bits) else None
                # stage 0 is built straight to the next group's height; later groups only need their LDE rows
                h = (nxt_bits + rate_bits if nxt_bits is not None else cap_height) if not groups else 0
                make = PolynomialBatch.from_coeffs if is_coeffs else PolynomialBatch.from_values
                groups.append(make(cols, rate_bits, False, h, ctx=ctx))
                group_of_poly += [(len(groups) - 1, j) for j in range(i + 1 - start)]
                degree_bits.append(d)
                start = i + 1
        if cap_height > degree_bits[-1] + rate_bits:
            raise N.ShapeError("cap_height=%d should be at most last_leaves_cap_height=%d" % (cap_height, degree_bits[-1] + rate_bits))
        return cls(groups, degree_bits, rate_bits, cap_height, group_of_poly, ctx)

    @property
    def polynomials(self):
        return [self.groups[g].polynomials[j] for g, j in self.group_of_poly]

    def values(self, leaf_index):
        """BatchMerkleTree::values (batch_merkle_tree.rs:154-164)."""
        h0 = self.leaf_heights[0]
        return [g.merkle_tree.get_rows(leaf_index >> (h0 - hk), 1)[0] for g, hk in zip(self.groups, self.leaf_heights)]

    def open_batch(self, leaf_index):
        """BatchMerkleTree::open_batch (batch_merkle_tree.rs:131-152)."""
        h0 = self.leaf_heights[0]
        sib = []
        for k, (st, hk) in enumerate(zip(self.stages, self.leaf_heights)):
            idx = leaf_index >> (h0 - hk)
            sib.append((st.merkle_tree if k == 0 else st).open_many([idx])[1][0])
        return MerkleProof(np.concatenate(sib) if sib else np.zeros((0, 4), dtype=np.uint64))

    def get_lde_values(self, degree_bits_index, index, step, slice_start, slice_len):
        """get_lde_values (oracle.rs:186-199)."""
        return self.groups[degree_bits_index].get_lde_values(index, step)[slice_start:slice_start + slice_len]

    def close(self):
        for g in self.groups:
            g.close()
        for t in self.stages[1:]:
            t.close()


def batch_prove_openings(degree_bits, instances, oracles, challenger, fri_params):
    """BatchFriOracle::prove_openings + batch_fri_proof (oracle.rs:124-183, prover.rs:30-147). instances[i] opens the
    polynomials of degree 2^degree_bits[i]; polynomial indices are indices into each oracle's full polynomial list."""
    assert len(degree_bits) == len(instances)
    L = N.lib()
    ctx = oracles[0].ctx
    alpha = challenger.get_extension_challenge()
    states = []
    try:
        for db, inst in zip(degree_bits, instances):
            # the polynomials of this instance live in the degree-`db` group of their oracle
            group_batches, handles, index_of = [], [], {}
            for o_idx, o in enumerate(oracles):
                if db in o.degree_bits:
                    index_of[o_idx] = len(handles)
                    handles.append(o.groups[o.degree_bits.index(db)])
            for b in inst.batches:
                polys = []
                for p in b.polynomials:
                    g, j = oracles[p.oracle_index].group_of_poly[p.polynomial_index]
                    assert oracles[p.oracle_index].degree_bits[g] == db, "polynomial of another degree in this instance"
                    polys.append(F.FriPolynomialInfo(index_of[p.oracle_index], j))
                group_batches.append(F.FriBatchInfo(b.point, polys))
            sub = F.FriInstanceInfo([F.FriOracleInfo(h.num_polys, False) for h in handles], group_batches)
            params_i = F.FriParams(fri_params.config, fri_params.hiding, db, fri_params.reduction_arity_bits)
            states.append(F._begin(sub, handles, alpha, params_i))
        # batch_fri_committed_trees (prover.rs:88-147)
        main = states[0]
        cap_words = 4 << fri_params.config.cap_height
        caps, nxt = [], 1
        log_cur = degree_bits[0] + fri_params.config.rate_bits
        for arity_bits in fri_params.reduction_arity_bits:
            cap = np.empty(cap_words, dtype=np.uint64)
            N.check(L.gl_fri_commit_round(main.h, arity_bits, N.np_ptr(cap)), ctx.h)
            cap = MerkleCap(cap)
            challenger.observe_cap(cap)
            caps.append(cap)
            beta = challenger.get_extension_challenge()
            b = np.array(beta, dtype=np.uint64)
            N.check(L.gl_fri_fold(main.h, N.np_ptr(b)), ctx.h)
            log_cur -= arity_bits
            if nxt < len(states) and log_cur == degree_bits[nxt] + fri_params.config.rate_bits:
                N.check(L.gl_fri_mix(main.h, states[nxt].h, N.np_ptr(b)), ctx.h)
                nxt += 1
        assert nxt == len(states), "reduction_arity_bits must pass through every instance's LDE size (prover.rs:44-57)"
        n_final = 1 << (log_cur - fri_params.config.rate_bits)
        buf = np.empty(2 * max(n_final, 1), dtype=np.uint64)
        ln = C.c_size_t()
        N.check(L.gl_fri_final_poly(main.h, N.np_ptr(buf), buf.size, C.byref(ln)), ctx.h)
        final = buf[:2 * ln.value].reshape(-1, 2).copy()
        challenger.observe_extension_elements([(int(c[0]), int(c[1])) for c in final])
        pow_witness = F.fri_proof_of_work(challenger, fri_params.config, ctx)
        # batch_fri_prover_query_rounds (prover.rs:149-215)
        n = 1 << (degree_bits[0] + fri_params.config.rate_bits)
        nq = fri_params.config.num_query_rounds
        x_indices = [c % n for c in challenger.get_n_challenges(nq)]
        rounds = []
        for x in x_indices:
            init = F.FriInitialTreeProof([(np.concatenate(o.values(x)), o.open_batch(x).siblings) for o in oracles])
            steps, xi, lc = [], x, degree_bits[0] + fri_params.config.rate_bits
            for r, arity_bits in enumerate(fri_params.reduction_arity_bits):
                xi >>= arity_bits
                layers = lc - arity_bits - fri_params.config.cap_height
                leaf = np.empty((1, 2 << arity_bits), dtype=np.uint64)
                path = np.empty((1, layers, 4), dtype=np.uint64)
                idx = np.array([xi], dtype=np.uint64)
                N.check(L.gl_fri_open(main.h, r, N.np_ptr(idx), 1, N.np_ptr(leaf), N.np_ptr(path) if path.size else None), ctx.h)
                steps.append(F.FriQueryStep(leaf[0].reshape(-1, 2), path[0]))
                lc -= arity_bits
            rounds.append(F.FriQueryRound(init, steps))
        return F.FriProof(caps, rounds, final, pow_witness)
    finally:
        for st in states:
            st.close()
