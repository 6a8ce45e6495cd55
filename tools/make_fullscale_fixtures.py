"""Full-scale golden fixtures from the CPU oracle (run HERE, on the CPU; minutes and ~40 GB of RAM per config).

For the configurations bench.py times -- BASELINE.json configs[1] (234 x 2^20, rate 1/8, cap 4, seed 0x02) and
configs[4] (64 x 2^24, rate 1/2, cap 4, seed 0x05) -- run the oracle's PolynomialBatch::from_values once on the
SURVEY 8(d) splitmix64 input and keep what a GPU box can check without repeating the run:
  * the Merkle cap (16 hashes),
  * for K sampled leaf indices: the leaf digest hash_or_noop(row), the first 8 words of the row, and for the first
    4 of them the full row and the Merkle siblings,
  * a 64-bit checksum of the coefficient matrix (sum over columns of sum_k coeffs[k] * (k+1) mod 2^64).
tests/test_gpu_fullscale.py compares the CUDA path with these; bench.py asserts the cap.
    python tools/make_fullscale_fixtures.py cfg2|cfg5|small
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle_lib
from conftest import synth

CONFIGS = {  # name: (seed, columns, log_n, rate_bits, cap_height)
    "cfg2": (0x02, 234, 20, 3, 4),
    "cfg5": (0x05, 64, 24, 1, 4),
    "small": (0x02, 234, 12, 3, 4),  # same generator, CPU-test sized (checks the tool and the test plumbing)
}


def sample_indices(seed, N, k):
    v = synth(0x5A17 + seed, (k,), canonical=False)
    idx = [int(x) % N for x in v]
    idx[0], idx[1] = 0, N - 1
    return idx


def coeff_checksum(coeffs):
    with np.errstate(over="ignore"):
        w = np.arange(1, coeffs.shape[1] + 1, dtype=np.uint64)
        return int(np.bitwise_xor.reduce((coeffs * w).sum(axis=1, dtype=np.uint64) * np.arange(1, coeffs.shape[0] + 1, dtype=np.uint64)))


def main(name):
    seed, B, log_n, r, h = CONFIGS[name]
    N = 1 << (log_n + r)
    t0 = time.perf_counter()
    vals = synth(seed, (B, 1 << log_n))
    t1 = time.perf_counter()
    c = oracle_lib.Commit(vals, r, h)
    t2 = time.perf_counter()
    idx = sample_indices(seed, N, 64)
    rows = c.leaf_rows(idx)
    digests = oracle_lib.hash_many(rows, 1)
    co = np.ctypeslib.as_array(oracle_lib.lib().glo_commit_coeffs(c.h), shape=(B << log_n,)).reshape(B, 1 << log_n)
    fx = {
        "config": {"seed": seed, "columns": B, "log_n": log_n, "rate_bits": r, "cap_height": h,
                   "generator": "tests/conftest.py synth(seed, (columns, 2^log_n)) = splitmix64 counter (SURVEY 8d)"},
        "made_by": "tools/make_fullscale_fixtures.py %s (CPU oracle, %d threads, synth %.1f s, commit %.1f s)" % (
            name, oracle_lib.nproc(), t1 - t0, t2 - t1),
        "cap": [[int(x) for x in hsh] for hsh in c.cap],
        "leaf_indices": idx,
        "leaf_digests": [[int(x) for x in d] for d in digests],
        "leaf_head": [[int(x) for x in row[:8]] for row in rows],
        "full_rows": [[int(x) for x in row] for row in rows[:4]],
        "siblings": [[[int(x) for x in s] for s in c.prove(i)] for i in idx[:4]],
        "coeff_checksum": coeff_checksum(co),
    }
    out = os.path.join(ROOT, "tests", "golden", "fullscale_%s.json" % name)
    json.dump(fx, open(out, "w"))
    print("wrote %s (%d bytes): cap0=%s, synth %.1f s, oracle commit %.1f s" % (
        out, os.path.getsize(out), fx["cap"][0], t1 - t0, t2 - t1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "small")
