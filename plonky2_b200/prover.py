"""Pieces of plonky2/src/plonk/prover.rs adjacent to the hot path ("next" rows of SURVEY.md section 8f) that
run on the GPU: the Z / partial-product columns of the permutation argument."""
import numpy as np

from . import _native as N
from .field import log2_strict


def wires_permutation_partial_products_and_zs(wires, sigmas, k_is, beta, gamma, degree, ctx=None):
    """wires_permutation_partial_products_and_zs (prover.rs:387-449): wires, sigmas are (num_routed, n) arrays of
    column values; returns (num_partial_products + 1, n): the partial-product columns, then Z (the reference's
    return order; the prover moves Z to the front before committing, prover.rs:227-232)."""
    ctx = ctx or N.default_context()
    wires = np.ascontiguousarray(wires, dtype=np.uint64)
    sigmas = np.ascontiguousarray(sigmas, dtype=np.uint64)
    k_is = np.ascontiguousarray(k_is, dtype=np.uint64)
    if wires.shape != sigmas.shape or wires.ndim != 2 or len(k_is) != wires.shape[0]:
        raise N.ShapeError("wires, sigmas must be (num_routed, n) and k_is (num_routed,)")
    R, n = wires.shape
    log_n = log2_strict(n)
    out = np.empty(((R + degree - 1) // degree, n), dtype=np.uint64)
    rc = N.lib().gl_partial_products_and_zs(ctx.h, N.np_ptr(wires), N.np_ptr(sigmas), N.np_ptr(k_is), log_n, R,
                                            int(beta), int(gamma), int(degree), N.np_ptr(out), N.MEM_HOST)
    N.check(rc, ctx.h)   # GL_ERR_DIV_ZERO -> ZeroDivisionError ("Tried to invert zero"), others keep their own type
    return out
