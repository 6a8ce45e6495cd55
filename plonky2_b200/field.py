"""Goldilocks scalars on the host (Python ints) for the sequential, latency-bound parts of the
protocol (transcript, parameters). Bulk arithmetic lives in the CUDA library.

Reference: field/src/goldilocks_field.rs:13-25,80,87,198; field/src/types.rs:226-272,429-443;
field/src/goldilocks_extensions.rs:14-27; field/src/extension/quadratic.rs:86-100,180-193."""

ORDER = 0xFFFFFFFF00000001
EPSILON = 0xFFFFFFFF
MULTIPLICATIVE_GROUP_GENERATOR = 14293326489335486720
POWER_OF_TWO_GENERATOR = 7277203076849721926
TWO_ADICITY = 32
NEG_ONE = ORDER - 1
D = 2  # extension degree of PoseidonGoldilocksConfig
W = 7  # X^2 = 7


def to_canonical_u64(x):
    return int(x) % ORDER


def coset_shift():
    return MULTIPLICATIVE_GROUP_GENERATOR


def primitive_root_of_unity(n_log):
    assert n_log <= TWO_ADICITY
    return pow(POWER_OF_TWO_GENERATOR, 1 << (TWO_ADICITY - n_log), ORDER)


def inverse_2exp(k):
    if k <= TWO_ADICITY:
        return ORDER - ((ORDER - 1) >> k)
    return pow(pow(2, k, ORDER), ORDER - 2, ORDER)


def inverse(x):
    x %= ORDER
    if x == 0:
        raise ZeroDivisionError("Tried to invert zero")
    return pow(x, ORDER - 2, ORDER)


def ext_mul(a, b):
    return ((a[0] * b[0] + W * a[1] * b[1]) % ORDER, (a[0] * b[1] + a[1] * b[0]) % ORDER)


def ext_add(a, b):
    return ((a[0] + b[0]) % ORDER, (a[1] + b[1]) % ORDER)


def ext_sub(a, b):
    return ((a[0] - b[0]) % ORDER, (a[1] - b[1]) % ORDER)


def ext_inverse(a):
    norm = (a[0] * a[0] - W * a[1] * a[1]) % ORDER
    ni = inverse(norm)
    return (a[0] * ni % ORDER, (-a[1]) * ni % ORDER)


def ext_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = ext_mul(r, a)
        a = ext_mul(a, a)
        e >>= 1
    return r


def log2_strict(n):
    lg = int(n).bit_length() - 1
    if n <= 0 or (1 << lg) != n:
        raise ValueError("Not a power of two: %d" % n)  # util/src/lib.rs:28
    return lg


def reverse_bits(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r
