"""Word-level model of the running accumulators of lcpc_mont32.cuh (m32::SplitAcc / split_mac / split_sum, m32::KaraAcc /
kara_mac / kara_sum) in Python: the same pairs, the same carry chains, the same sink indices, every word reduced mod 2^32
where the device code wraps.  Checked against plain integer arithmetic on random operands and on operands made of
all-one and all-zero words (every chain carries out, every Karatsuba half-sum overflows), for every word count the
library instantiates (N = 2, 4, 6, 8; Karatsuba halves H = 2, 3, 4) and for long sums.  The device code itself is checked
bit for bit against the oracle by the -m gpu tests; this pins the index arithmetic those tests can only sample."""
import random

M32 = (1 << 32) - 1


class SplitAcc:
    def __init__(self, n):
        self.n = n
        self.X = [0] * n          # X[i] = words 2i, 2i+1
        self.Y = [0] * n          # Y[i] = words 2i+1, 2i+2
        self.Z = [0] * (n + 2)    # carries of weight 2^(32(n+k))


def mad_chain(pairs, first, np_, x, y, y0, sink, k):
    """pairs[first .. first+np_) += x * (y[y0], y[y0+2], ...) as one carry chain; the carry out goes to sink[k]."""
    carry = 0
    for t in range(np_):
        v = pairs[first + t] + x * y[y0 + 2 * t] + carry
        pairs[first + t] = v & ((1 << 64) - 1)
        carry = v >> 64
    assert carry in (0, 1)
    sink[k] = (sink[k] + carry) & M32


def split_mac(s, a, b):
    n = s.n
    for i in range(n):
        j0 = i & 1
        np_ = (n - j0 + 1) // 2
        end = i + j0 + 2 * np_
        mad_chain(s.X, (i + j0) // 2, np_, b[i], a, j0, s.Z, end - n)
        j0 = 1 - (i & 1)
        np_ = (n - j0 + 1) // 2
        end = i + j0 + 2 * np_
        mad_chain(s.Y, (i + j0 - 1) // 2, np_, b[i], a, j0, s.Z, end - n)


def words_to_int(w):
    return sum(v << (32 * i) for i, v in enumerate(w))


def split_sum(s):
    """2n + 2 words, exactly as the device code adds them (no carry may leave the top word)."""
    n = s.n
    S = [0] * (2 * n + 2)
    Yw = [0] * (2 * n + 2)
    for i in range(n):
        S[2 * i], S[2 * i + 1] = s.X[i] & M32, s.X[i] >> 32
        Yw[2 * i + 1], Yw[2 * i + 2] = s.Y[i] & M32, s.Y[i] >> 32
    total = words_to_int(S) + words_to_int(Yw) + (words_to_int(s.Z) << (32 * n))
    assert total < 1 << (32 * (2 * n + 2))
    return total


def kara_terms(a, b, h):
    aL, aH, bL, bH = a[:h], a[h:], b[:h], b[h:]
    sa, sb = words_to_int(aL) + words_to_int(aH), words_to_int(bL) + words_to_int(bH)
    ca, cb = sa >> (32 * h), sb >> (32 * h)
    sa_w = [(sa >> (32 * i)) & M32 for i in range(h)]
    sb_w = [(sb >> (32 * i)) & M32 for i in range(h)]
    return aL, aH, bL, bH, sa_w, sb_w, ca, cb


def kara_dot(terms, n):
    h = n // 2
    p0, p2, pm = SplitAcc(h), SplitAcc(h), SplitAcc(h)
    Ca = Cb = Cc = 0
    for a, b in terms:
        aL, aH, bL, bH, sa, sb, ca, cb = kara_terms(a, b, h)
        split_mac(p0, aL, bL)
        split_mac(p2, aH, bH)
        split_mac(pm, sa, sb)
        Ca += words_to_int(sb) if ca else 0
        Cb += words_to_int(sa) if cb else 0
        Cc += ca & cb
    assert Ca < 1 << (32 * (h + 1)) and Cb < 1 << (32 * (h + 1))
    W = 1 << (32 * h)
    S0, S2 = split_sum(p0), split_sum(p2)
    Sm = split_sum(pm) + (Ca + Cb) * W + Cc * W * W
    assert Sm < 1 << (32 * (2 * h + 2))          # what kara_sum holds in 2h + 2 words
    mid = Sm - S0 - S2
    assert mid >= 0
    return S0 + mid * W + S2 * W * W


def _operands(n, rng, k):
    ones, zero = [M32] * n, [0] * n
    half = [M32] * (n // 2) + [0] * (n - n // 2)
    pool = [ones, zero, half, half[::-1], [1] + [0] * (n - 1), [0] * (n - 1) + [M32]]
    out = [(a, b) for a in pool for b in pool]
    out += [([rng.randrange(1 << 32) for _ in range(n)], [rng.randrange(1 << 32) for _ in range(n)]) for _ in range(k)]
    return out


def test_split_accumulator_equals_the_integer_sum():
    rng = random.Random(1)
    for n in (2, 4, 6, 8):
        terms = _operands(n, rng, 300)
        s = SplitAcc(n)
        want = 0
        for a, b in terms:
            split_mac(s, a, b)
            want += words_to_int(a) * words_to_int(b)
        assert split_sum(s) == want, n
        # a long sum of the worst term: every chain carries out every time, the sinks count them
        s, want = SplitAcc(n), 0
        for _ in range(5000):
            split_mac(s, [M32] * n, [M32] * n)
            want += ((1 << (32 * n)) - 1) ** 2
        assert split_sum(s) == want, n


def test_karatsuba_accumulator_equals_the_integer_sum():
    rng = random.Random(2)
    for n in (4, 6, 8):
        terms = _operands(n, rng, 300)
        assert kara_dot(terms, n) == sum(words_to_int(a) * words_to_int(b) for a, b in terms), n
        worst = [([M32] * n, [M32] * n)] * 3000
        assert kara_dot(worst, n) == 3000 * ((1 << (32 * n)) - 1) ** 2, n
