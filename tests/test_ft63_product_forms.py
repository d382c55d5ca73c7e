"""The Ft63 Montgomery product of lcpc_field.cuh in Python integers: the digit steps spelled as three-operand 64-bit sums
(ft63::redc_cols<true>, ft63::to_canon) against the plain digit form they replaced (redc_digit twice) and against the
definition a*b*2^-64 mod p -- on random operands and on the ones a random GPU test never meets (a digit word equal to 0 or
0xffffffff, operands 0 / 1 / p-1, high words at their bounds).  Every step is reduced mod 2^64 exactly where the device
code wraps."""
import random

P = 0x46D0760000000001
P_HI = 0x46D07600
Q = (1 << 32) - P_HI
M64 = (1 << 64) - 1
M32 = (1 << 32) - 1
RINV = pow(1 << 64, -1, P)


def lo32(x):
    return x & M32


def hi32(x):
    return (x >> 32) & M32


def pack(lo, hi):
    return (lo & M32) | ((hi & M32) << 32)


def fix(v):
    """v in [-p, p) as a two's-complement 64-bit value -> [0, p)."""
    m = (v >> 63) & 1
    lo = lo32(v) + m
    hi = (hi32(v) + m * P_HI + (lo >> 32)) & M32
    return pack(lo, hi)


def redc_digit(w, acc):
    u = (acc + w * Q) & M64
    return pack(lo32(u), hi32(u) - w)


def redc_cols_plain(p00, col1, p11):
    u = redc_digit(lo32(p00), (col1 + P + hi32(p00)) & M64)
    v = redc_digit(lo32(u), (p11 + hi32(u)) & M64)
    return fix(v)


def redc_cols_sum3(p00, col1, p11):
    w = lo32(p00)
    u = ((col1 + w * Q) + (P + (1 << 32)) + pack(hi32(p00), ~w)) & M64
    w2 = lo32(u)
    v = (p11 + pack(hi32(u), -w2) + w2 * Q) & M64
    return fix(v)


def mul(a, b, cols):
    a0, a1, b0, b1 = lo32(a), hi32(a), lo32(b), hi32(b)
    return cols(a0 * b0, (a0 * b1 + a1 * b0) & M64, a1 * b1)


def to_canon_sum3(a):
    w = lo32(a)
    u = (w * Q + pack(hi32(a), ~w) + (P + (1 << 32))) & M64
    w2 = lo32(u)
    return fix((pack(hi32(u), -w2) + w2 * Q) & M64)


def _operands():
    rng = random.Random(63)
    edge = [0, 1, 2, P - 1, P - 2, M32, 1 << 32, (1 << 32) + 1, P_HI << 32, (P_HI << 32) - 1, (P_HI - 1) << 32 | M32,
            0x46D0760000000000, 0x00000000FFFFFFFF, 0x0000000100000000, 0x46D075FFFFFFFFFF]
    ops = [(a, b) for a in edge for b in edge]
    ops += [(rng.randrange(P), rng.randrange(P)) for _ in range(20000)]
    # operands whose low partial product has an all-zero or all-one low word (the digit w = 0 / 0xffffffff)
    for _ in range(2000):
        b = rng.randrange(P) | 1
        b0inv = pow(lo32(b), -1, 1 << 32)
        for target in (0, M32, 1):
            a0 = (target * b0inv) & M32
            ops.append((pack(a0, rng.randrange(P_HI)), b))
    return ops


def test_three_operand_digit_steps_equal_the_plain_form_and_the_definition():
    for a, b in _operands():
        want = a * b * RINV % P
        got = mul(a, b, redc_cols_sum3)
        assert got == want, (hex(a), hex(b))
        assert mul(a, b, redc_cols_plain) == want, (hex(a), hex(b))


def test_to_canon_form():
    rng = random.Random(64)
    vals = [0, 1, P - 1, M32, 1 << 32, P_HI << 32, 0x46D0760000000000] + [rng.randrange(P) for _ in range(20000)]
    vals += [pack(0, rng.randrange(P_HI)) for _ in range(200)] + [pack(M32, rng.randrange(P_HI)) for _ in range(200)]
    for a in vals:
        assert to_canon_sum3(a) == a * RINV % P, hex(a)
        assert redc_cols_plain(lo32(a), hi32(a), 0) == a * RINV % P, hex(a)
