"""NTT entry points mirroring field/src/fft.rs:53-91 and field/src/polynomial/mod.rs:58-88,199-201,280-293.

Arrays are numpy uint64, shape (n,) for one polynomial or (batch, n) for a batch of columns; every
call runs the CUDA four-step NTT (gl_ntt) -- natural order in, natural order out, canonical output."""
import numpy as np

from . import _native as N
from .field import coset_shift, log2_strict


def _run(a, inverse, zero_factor, shift, ctx):
    ctx = ctx or N.default_context()
    a = np.array(a, dtype=np.uint64, order="C", copy=True)
    one = a.ndim == 1
    m = a.reshape(1, -1) if one else a
    if m.ndim != 2:
        raise N.ShapeError("expected (n,) or (batch, n)")
    batch, n = m.shape
    log_n = log2_strict(n)
    N.check(N.lib().gl_ntt(ctx.h, N.np_ptr(m), log_n, batch, n, int(inverse), int(zero_factor or 0),
                           int(shift), N.MEM_HOST), ctx.h)
    return m.reshape(-1) if one else m


def fft_with_options(coeffs, zero_factor=None, root_table=None, ctx=None):
    """fft_with_options (fft.rs:53-61). `root_table` is accepted for signature parity and ignored:
    twiddles are cached on the device per context."""
    return _run(coeffs, False, zero_factor, 1, ctx)


def fft(coeffs, ctx=None):
    return fft_with_options(coeffs, None, None, ctx)


def ifft_with_options(values, zero_factor=None, root_table=None, ctx=None):
    """ifft_with_options (fft.rs:68-91)."""
    return _run(values, True, zero_factor, 1, ctx)


def ifft(values, ctx=None):
    return ifft_with_options(values, None, None, ctx)


def coset_fft_with_options(coeffs, shift, zero_factor=None, root_table=None, ctx=None):
    """PolynomialCoeffs::coset_fft_with_options (polynomial/mod.rs:280-293)."""
    return _run(coeffs, False, zero_factor, shift, ctx)


def coset_fft(coeffs, shift, ctx=None):
    return coset_fft_with_options(coeffs, shift, None, None, ctx)


def coset_ifft(values, shift, ctx=None):
    """PolynomialValues::coset_ifft (polynomial/mod.rs:63-73)."""
    return _run(values, True, None, shift, ctx)


def lde(coeffs, rate_bits):
    """PolynomialCoeffs::lde (polynomial/mod.rs:199-201): zero-pad to n << rate_bits."""
    coeffs = np.asarray(coeffs, dtype=np.uint64)
    pad = list(coeffs.shape)
    pad[-1] = coeffs.shape[-1] * ((1 << rate_bits) - 1)
    return np.concatenate([coeffs, np.zeros(pad, dtype=np.uint64)], axis=-1)


def lde_onto_coset(values, rate_bits, ctx=None):
    """PolynomialValues::lde_onto_coset (polynomial/mod.rs:85-88)."""
    return coset_fft_with_options(lde(ifft(values, ctx), rate_bits), coset_shift(), rate_bits, None, ctx)
