"""Small pass over every kernel family (NTT passes incl. partial tiles, LDE, hashing, Merkle, openings, FRI,
PoW, eval, partial products) meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import plonky2_b200 as pb
from plonky2_b200.prover import wires_permutation_partial_products_and_zs
from conftest import synth

for log_n in (1, 5, 9, 13):
    x = synth(log_n, (3, 1 << log_n))
    y = pb.fft(x); assert np.array_equal(pb.ifft(y), x)
    assert np.array_equal(pb.coset_ifft(pb.coset_fft(x, 5), 5), x)
for (B, log_n, r, h) in [(9, 6, 3, 2), (11, 13, 1, 4), (3, 2, 2, 1)]:
    v = synth(B, (B, 1 << log_n))
    c = pb.PolynomialBatch.from_values(v, r, True, h, salt=synth(99, (4, 1 << (log_n + r))))
    c.merkle_tree.open_many([0, (1 << (log_n + r)) - 1]); c.eval_commitment((3, 4)); c.get_lde_values(1, 1)
    s = pb.PolynomialBatch.from_values(v, r, False, h, shard=(1, 2)); s.close()
    c.close()
t = pb.MerkleTree(synth(7, (64, 7)), 3); t.open_many([5, 63]); t.close()
pb.PoseidonHash.hash_many(synth(8, (33, 19))); pb.PoseidonHash.two_to_one_many(synth(9, (17, 8)))
log_n, Bs = 8, [4, 3]
cm = [pb.PolynomialBatch.from_values(synth(20 + i, (B, 1 << log_n)), 3, False, 2) for i, B in enumerate(Bs)]
inst = pb.FriInstanceInfo([pb.FriOracleInfo(B, False) for B in Bs],
                          [pb.FriBatchInfo((11, 12), [pb.FriPolynomialInfo(o, i) for o, B in enumerate(Bs) for i in range(B)]),
                           pb.FriBatchInfo((13, 14), [pb.FriPolynomialInfo(1, 0)])])
ch = pb.Challenger(); [ch.observe_cap(c.merkle_tree.cap) for c in cm]
params = pb.FriParams(pb.FriConfig(3, 2, 6, ("Fixed", [2, 3]), 5), False, log_n, [2, 3])
pr = pb.prove_openings(inst, cm, ch, params); assert len(pr.to_bytes()) > 0
wires_permutation_partial_products_and_zs(synth(30, (10, 64)), synth(31, (10, 64)), synth(32, (10,)), 5, 6, 4)
print("SANITIZE SMOKE DONE")
