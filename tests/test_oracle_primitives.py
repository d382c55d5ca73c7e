"""Pin the CPU oracle's primitives against independent implementations.

The reference's tests hold no golden vectors for this path, so the
oracle is pinned piece by piece on the third-party algorithms it restates:
blake3 (Python `blake3` package = bindings to the Rust crate the reference uses),
Montgomery field arithmetic (Python integers), the NTT definition (direct DFT sum),
ChaCha20 (`cryptography`), Keccak-f[1600] (hashlib SHA3-256) and merlin (its own
published known-answer vector).
"""
import ctypes as C
import hashlib
import random
import struct

import numpy as np
import pytest

blake3_pkg = pytest.importorskip("blake3")

FIELDS = [0, 1, 2, 3, 4]


@pytest.mark.parametrize(
    "n",
    [0, 1, 31, 32, 33, 63, 64, 65, 127, 128, 1023, 1024, 1025, 2047, 2048, 2049, 3072, 3073,
     4096, 4128, 5000, 7168, 8192, 8193, 16416, 31744, 65536, 149832],
)
def test_blake3_matches_rust_crate_bindings(oracle, n):
    rnd = random.Random(n)
    data = bytes(rnd.getrandbits(8) for _ in range(n))
    assert oracle.blake3(data) == blake3_pkg.blake3(data).digest()


def test_blake3_incremental_equals_one_shot(oracle):
    """digest_update is called once per element (lcpc-2d/src/lib.rs:757-761); chunking must not matter."""
    lib = oracle.lib()
    rnd = random.Random(7)
    data = bytes(rnd.getrandbits(8) for _ in range(5000))

    class H(C.Structure):
        _fields_ = [("cv", C.c_uint32 * 8), ("chunk_counter", C.c_uint64), ("buf", C.c_uint8 * 64),
                    ("buf_len", C.c_uint8), ("blocks", C.c_uint8), ("stack", (C.c_uint32 * 8) * 54),
                    ("stack_len", C.c_uint8)]

    for piece in (1, 7, 8, 32, 64, 100, 1024, 1500):
        h = H()
        lib.orc_b3_init(C.byref(h))
        for off in range(0, len(data), piece):
            chunk = data[off:off + piece]
            lib.orc_b3_update(C.byref(h), chunk, C.c_size_t(len(chunk)))
        out = (C.c_uint8 * 32)()
        lib.orc_b3_finalize(C.byref(h), out)
        assert bytes(out) == blake3_pkg.blake3(data).digest()


@pytest.mark.parametrize("fid", FIELDS)
def test_field_constants_rederived(oracle, fid):
    """Every constant in orc_lcpc.c's table is recomputed from the modulus and generator
    in the reference's derive attributes (lcpc-test-fields/src/lib.rs:19-20,42-43,54-55,66-67)."""
    L = oracle.LIMBS[fid]
    p, g = oracle.MODULUS[fid], oracle.GENERATOR[fid]
    arr = lambda: np.zeros(4, dtype=np.uint64)
    pp, r, r2, root = arr(), arr(), arr(), arr()
    inv = C.c_uint64()
    s = C.c_int()
    nb = C.c_int()
    u64p = C.POINTER(C.c_uint64)
    oracle.lib().orc_field_constants(C.c_int(fid), pp.ctypes.data_as(u64p), C.byref(inv), r.ctypes.data_as(u64p),
                                     r2.ctypes.data_as(u64p), root.ctypes.data_as(u64p), C.byref(s), C.byref(nb))
    toint = lambda a: sum(int(x) << (64 * i) for i, x in enumerate(a[:L]))
    R = 1 << (64 * L)
    assert toint(pp) == p
    assert inv.value == (-pow(p, -1, 1 << 64)) % (1 << 64)
    assert toint(r) == R % p
    assert toint(r2) == R * R % p
    S, t = 0, p - 1
    while t % 2 == 0:
        t //= 2
        S += 1
    assert s.value == S == oracle.TWO_ADICITY[fid]
    assert nb.value == p.bit_length() == oracle.NUM_BITS[fid]
    assert toint(root) == pow(g, t, p) * R % p
    assert 2 * p < R  # one spare bit: lazy [0, 2p) values fit the limbs


@pytest.mark.parametrize("fid", FIELDS)
def test_field_arithmetic_vs_python_ints(oracle, fid):
    O = oracle
    p, R = O.MODULUS[fid], O.mont_R(fid)
    Rinv = pow(R, -1, p)
    rnd = random.Random(fid)
    a = [rnd.randrange(p) for _ in range(300)] + [0, 1, p - 1, p - 1, 0]
    b = [rnd.randrange(p) for _ in range(300)] + [p - 1, p - 1, p - 1, 1, 0]
    A, B = O.to_limbs(fid, a), O.to_limbs(fid, b)
    assert O.from_limbs(O.fe_mul(fid, A, B)) == [x * y * Rinv % p for x, y in zip(a, b)]
    assert O.from_limbs(O.fe_add(fid, A, B)) == [(x + y) % p for x, y in zip(a, b)]
    assert O.from_limbs(O.fe_sub(fid, A, B)) == [(x - y) % p for x, y in zip(a, b)]
    assert O.from_limbs(O.fe_to_canon(fid, A)) == [x * Rinv % p for x in a]
    assert O.from_limbs(O.fe_from_canon(fid, A)) == [x * R % p for x in a]
    nz = O.to_limbs(fid, [x for x in a if x])
    assert O.from_limbs(O.fe_mul(fid, O.fe_inv(fid, nz), nz)) == [R % p] * len(nz)


@pytest.mark.parametrize("fid", FIELDS)
@pytest.mark.parametrize("log_n", [1, 2, 3, 6])
def test_fft_io_is_bit_reversed_dft(oracle, fid, log_n):
    """out[bitrev(i)] = sum_j in[j] w^(ij), w = ROOT_OF_UNITY^(2^(S-k)); ifft_oi inverts it
    (the only property the reference's tests fix: lcpc_online.rs:588-601, lcpc-2d tests.rs:223-233)."""
    O = oracle
    p, n = O.MODULUS[fid], 1 << log_n
    S = O.TWO_ADICITY[fid]
    w = O.from_mont(fid, O.ntt_root(fid, log_n))[0]
    assert w == pow(pow(O.GENERATOR[fid], (p - 1) >> S, p), 1 << (S - log_n), p)
    rnd = random.Random(100 * fid + log_n)
    x = [rnd.randrange(p) for _ in range(n)]
    X = O.fft_io(fid, O.to_mont(fid, x).reshape(1, n, -1))
    got = O.from_mont(fid, X[0])
    brev = lambda i: int(format(i, f"0{log_n}b")[::-1], 2)
    for i in range(n):
        assert got[brev(i)] == sum(x[j] * pow(w, i * j, p) for j in range(n)) % p
    assert O.from_mont(fid, O.ifft_oi(fid, X)[0]) == x


@pytest.mark.parametrize("fid", FIELDS)
@pytest.mark.parametrize("log_n", [1, 4, 7])
def test_fft_io_matches_the_published_fffft_loop(oracle, fid, log_n):
    """The one convention no file in the reference pins: the `fffft` crate is a path dependency outside
    the tree.  This restates its published forward transform as recalled -- an in-place Gentleman-Sande sweep, gaps
    n/2, n/4, ..., 1, butterfly (x, y) -> (x + y, (x - y) * roots[nchunks * idx]) over the table roots[i] = w^i with
    w = ROOT_OF_UNITY^(2^(S - log n)), no reordering pass ("io" = in-order in, out-of-order out) -- in plain Python
    integers, and checks the oracle against it.  If the crate differs, this loop is the place that says how."""
    O = oracle
    p, n = O.MODULUS[fid], 1 << log_n
    S = O.TWO_ADICITY[fid]
    w = pow(pow(O.GENERATOR[fid], (p - 1) >> S, p), 1 << (S - log_n), p)
    roots = [pow(w, i, p) for i in range(max(1, n // 2))]
    rnd = random.Random(7 * fid + log_n)
    x = [rnd.randrange(p) for _ in range(n)]
    xi = list(x)
    gap = n // 2
    while gap > 0:
        nchunks = n // (2 * gap)
        for cidx in range(nchunks):
            offset = 2 * cidx * gap
            for idx in range(gap):
                neg = (xi[offset + idx] - xi[offset + idx + gap]) % p
                xi[offset + idx] = (xi[offset + idx] + xi[offset + idx + gap]) % p
                xi[offset + idx + gap] = neg * roots[nchunks * idx] % p
        gap //= 2
    got = O.from_mont(fid, O.fft_io(fid, O.to_mont(fid, x).reshape(1, n, -1))[0])
    assert got == xi


def test_keccak_f1600_via_sha3(oracle):
    lib = oracle.lib()

    def sha3_256(msg):
        rate = 136
        st = (C.c_uint64 * 25)()
        m = bytearray(msg) + b"\x06"
        while len(m) % rate:
            m += b"\x00"
        m[-1] |= 0x80
        for off in range(0, len(m), rate):
            for i in range(rate // 8):
                st[i] ^= struct.unpack_from("<Q", m, off + 8 * i)[0]
            lib.orc_keccak_f1600(st)
        return b"".join(struct.pack("<Q", st[i]) for i in range(4))

    for msg in [b"", b"abc", b"x" * 135, b"y" * 136, b"z" * 1000]:
        assert sha3_256(msg) == hashlib.sha3_256(msg).digest()


class _Rng(C.Structure):
    _fields_ = [("key", C.c_uint32 * 8), ("counter", C.c_uint64), ("stream", C.c_uint64),
                ("buf", C.c_uint32 * 16), ("idx", C.c_int), ("rounds", C.c_int)]


@pytest.mark.parametrize("stream", [0, 1, 5])
def test_chacha20_keystream(oracle, stream):
    algorithms = pytest.importorskip("cryptography.hazmat.primitives.ciphers.algorithms")
    from cryptography.hazmat.primitives.ciphers import Cipher

    lib = oracle.lib()
    lib.orc_chacha_next_u32.restype = C.c_uint32
    lib.orc_chacha_next_u64.restype = C.c_uint64
    key = bytes(range(32))
    r = _Rng()
    lib.orc_chacha_from_seed(C.byref(r), key)
    lib.orc_chacha_set_stream(C.byref(r), C.c_uint64(stream))
    words = [lib.orc_chacha_next_u32(C.byref(r)) for _ in range(200)]
    # `cryptography` takes counter(4 LE) || nonce(12); with a 64-bit counter below 2^32 and the
    # 64-bit stream id in the last 8 bytes this is the djb layout rand_chacha uses.
    nonce = struct.pack("<QQ", 0, stream)
    ks = Cipher(algorithms.ChaCha20(key, nonce), mode=None).encryptor().update(b"\x00" * 800)
    assert list(struct.unpack("<200I", ks)) == words
    # next_u64 = two consecutive words, low first
    r2 = _Rng()
    lib.orc_chacha_from_seed(C.byref(r2), key)
    lib.orc_chacha_set_stream(C.byref(r2), C.c_uint64(stream))
    for i in range(50):
        assert lib.orc_chacha_next_u64(C.byref(r2)) == words[2 * i] | (words[2 * i + 1] << 32)


def test_seed_from_u64_is_pcg32_expansion(oracle):
    lib = oracle.lib()
    r = _Rng()
    for state0 in (0, 1, 1337, 2**64 - 1):
        lib.orc_chacha_seed_from_u64(C.byref(r), C.c_uint64(state0))
        st, out = state0, []
        for _ in range(8):
            st = (st * 6364136223846793005 + 11634580027462260723) % 2**64
            xs = (((st >> 18) ^ st) >> 27) & 0xFFFFFFFF
            rot = st >> 59
            out.append(((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF)
        assert list(r.key) == out


def test_uniform_usize_is_widening_multiply_rejection(oracle):
    lib = oracle.lib()
    lib.orc_uniform_usize.restype = C.c_uint64
    lib.orc_chacha_next_u64.restype = C.c_uint64
    key = bytes(31 - i for i in range(32))
    for n in (1, 2, 3, 1000, 65536, 252931, (1 << 63) + 12345):
        a, b = _Rng(), _Rng()
        lib.orc_chacha_from_seed(C.byref(a), key)
        lib.orc_chacha_from_seed(C.byref(b), key)
        zone = (2**64 - 1) - ((2**64 - n) % n)
        for _ in range(200):
            got = lib.orc_uniform_usize(C.byref(a), C.c_uint64(n))
            while True:
                v = lib.orc_chacha_next_u64(C.byref(b))
                hi, lo = divmod(v * n, 2**64)
                if lo <= zone:
                    break
            assert got == hi < n


def test_merlin_simple_transcript_kat(oracle):
    """merlin 2.0's own conformance vector (transcript.rs `equivalence_simple`)."""
    t = oracle.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == (
        "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615")


@pytest.mark.parametrize("fid", FIELDS)
def test_field_random_is_masked_rejection_of_raw_limbs(oracle, fid):
    """ff_derive `random`: LIMBS x next_u64, top limb masked to NUM_BITS, reject >= p, raw limbs kept."""
    lib = oracle.lib()
    lib.orc_chacha_next_u64.restype = C.c_uint64
    key = bytes((7 * i + fid) % 256 for i in range(32))
    L, p = oracle.LIMBS[fid], oracle.MODULUS[fid]
    got = oracle.from_limbs(oracle.random_field_vec(fid, key, 64))
    r = _Rng()
    lib.orc_chacha_from_seed(C.byref(r), key)
    exp = []
    mask = (1 << oracle.NUM_BITS[fid]) - 1
    while len(exp) < 64:
        v = sum(lib.orc_chacha_next_u64(C.byref(r)) << (64 * l) for l in range(L)) & mask
        if v < p:
            exp.append(v)
    assert got == exp


@pytest.mark.parametrize("n", [1025, 2048, 3000, 4128, 5000, 9 * 1024, 33 * 1024 + 5, 40 * 1024])
def test_blake3_chunk_values_and_parent_tree_vs_blake3_package(oracle, n):
    """The pieces a row-sharded hash is made of (orc_b3_chunk_cv per 1024-byte chunk, orc_b3_merge_cvs over them)
    reproduce the `blake3` package's digest of the whole message."""
    import ctypes as C

    blake3_pkg = pytest.importorskip("blake3")
    lib = oracle.lib()
    data = bytes((i * 7 + 3) % 251 for i in range(n))
    n_chunks = (n + 1023) // 1024
    cvs = np.zeros((n_chunks, 32), dtype=np.uint8)
    out = (C.c_uint8 * 32)()
    for c in range(n_chunks):
        chunk = data[c * 1024:(c + 1) * 1024]
        lib.orc_b3_chunk_cv(chunk, C.c_size_t(len(chunk)), C.c_uint64(c), out)
        cvs[c] = np.frombuffer(bytes(out), dtype=np.uint8)
    lib.orc_b3_merge_cvs(cvs.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(n_chunks), out)
    assert bytes(out) == blake3_pkg.blake3(data).digest()


@pytest.mark.parametrize("fid,n_rows,n_cols", [(0, 300, 5), (1, 130, 3), (3, 70, 4), (4, 70, 2)])
def test_row_windows_of_hash_columns(oracle, fid, n_rows, n_cols):
    """hash_chunk_cvs over two row windows that meet on a chunk boundary + hash_merge == hash_columns (lib.rs:736-775)."""
    O = oracle
    comm = O.random_field_elements(fid, 3, n_rows * n_cols).reshape(n_rows, n_cols, -1)
    n_chunks = O.leaf_chunks(fid, n_rows)
    w = 8 * O.LIMBS[fid]
    c_mid = n_chunks // 2
    r_mid = (1024 - 32) // w + (c_mid - 1) * (1024 // w) if c_mid >= 1 else 0
    a = O.hash_chunk_cvs(fid, comm[:r_mid], 0, n_rows, 0, c_mid)
    b = O.hash_chunk_cvs(fid, comm[r_mid:], r_mid, n_rows, c_mid, n_chunks)
    assert np.array_equal(O.hash_merge(np.concatenate([a, b])), O.hash_columns(fid, comm))
