# Find operands for which the ORIGINAL dual-accumulator Montgomery product (chain-end carry added into a data word)
# loses a carry with 32-bit words -- deterministic regression vectors for the multi-limb fields.
import random, json
W = 32; MASK = (1 << W) - 1
FIELDS = {1: 0x6e754097ba20e0bf7f2bd90000000001,
          2: 0x453708aa3fbc8dda936888270ceecbcdd246820000000001,
          3: 0x663c799b6e4d2900fda9df04b9575969ef73c79086595f3002a4f20000000001}
words = lambda v, n: [(v >> (W * i)) & MASK for i in range(n)]
val = lambda A: sum(w << (W * i) for i, w in enumerate(A))
def row_mad_old(N, I, X, Y, x, yw):
    lost = False
    for acc, j0 in ((X, I & 1), (Y, 1 - (I & 1))):
        carry = 0; j = j0; first = True
        while j < N:
            prod = x * yw[j]; lo, hi = prod & MASK, prod >> W
            t = acc[I + j] + lo + (0 if first else carry); acc[I + j] = t & MASK; carry = t >> W
            t = acc[I + j + 1] + hi + carry; acc[I + j + 1] = t & MASK; carry = t >> W
            first = False; j += 2
        end = I + j0 + ((N - j0 + 1) // 2) * 2
        t = acc[end] + carry; lost |= bool(t >> W); acc[end] = t & MASK
    return lost
def product(N, a, b):
    X = [0] * (2 * N + 2); Y = [0] * (2 * N + 2)
    aw, bw = words(a, N), words(b, N)
    for I in range(N): assert not row_mad_old(N, I, X, Y, bw[I], aw)
    return X, Y
def old_mul_loses(N, a, b, p):
    X, Y = product(N, a, b); pw = words(p, N); c = 0; lost = False
    for I in range(N):
        m = (-(X[I] + Y[I] + c)) & MASK
        lost |= row_mad_old(N, I, X, Y, m, pw)
        c = 1 if (X[I] | Y[I] | c) else 0
    return lost
random.seed(7)
out = []
for fid, p in FIELDS.items():
    N = {1: 4, 2: 6, 3: 8}[fid]
    found = []
    tries = 0
    while len(found) < 3:
        tries += 1
        a = random.randrange(p) | (1 << (W * (N - 1)))  # odd-ish top word helps the inverse below
        a = a % p
        aw = words(a, N)
        if aw[N - 1] % 2 == 0: continue
        b = random.randrange(p)
        inv = pow(aw[N - 1], -1, 1 << W)
        ok = False
        for it in range(8):  # steer word 1 of b until word N of X is all ones after the product phase
            X, Y = product(N, a, b)
            if X[N] == MASK: ok = True; break
            bw = words(b, N)
            bw[1] = (bw[1] + (MASK - X[N]) * inv) & MASK
            b = val(bw)
            if b >= p: break
        if not ok or b >= p: continue
        if old_mul_loses(N, a, b, p): found.append((a, b))
    R = 1 << (64 * (N // 2))
    for a, b in found:
        out.append({"fid": fid, "a_montgomery": hex(a), "b_montgomery": hex(b), "product_montgomery": hex(a * b * pow(R, -1, p) % p)})
    print("field", fid, "tries", tries)
json.dump(out, open(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), 'carry_kats.json'), 'w'), indent=1)
print(len(out))
