"""Parity of the CUDA path (through the C ABI) against the CPU oracle, bit for bit.

Mirrors the reference's own tests for this path (lcpc-2d/src/tests.rs,
lcpc-ligero-pc/src/tests.rs, lcpc-brakedown-pc/src/tests.rs): same seeded inputs on
both sides, every output array compared exactly.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = [0, 1, 2, 3, 4]


@pytest.fixture(scope="module")
def P():
    import lcpc_proof_of_storage_b200 as pkg

    return pkg


def _rand(O, fid, seed, n):
    return O.random_field_elements(fid, seed, n)


# ----------------------------------------------------------------------------- encode (K1)

@pytest.mark.parametrize("fid", FIELDS)
@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 5, 7, 9, 10, 11, 12, 13, 14, 16])
def test_ligero_encode_rows_match_fft_io(P, oracle, fid, log_n):
    """LcEncoding::encode == fffft fft_io_pc on whole rows (lcpc-ligero-pc/src/lib.rs:162-164)."""
    O = oracle
    n = 1 << log_n
    n_rows = 3 if log_n < 14 else 2
    rows = _rand(O, fid, 1000 + log_n, n_rows * n).reshape(n_rows, n, -1)
    enc = P.LigeroEncoding(fid, max(1, n // 2), n)
    got = rows.copy()
    enc.encode(got)
    assert np.array_equal(got, O.fft_io(fid, rows))


@pytest.mark.parametrize("fid,log_n", [(0, 17), (0, 18), (1, 17), (3, 15), (4, 15)])
def test_ligero_encode_large_rows(P, oracle, fid, log_n):
    """Rows that need two or more strided passes in front of the shared-memory pass."""
    O = oracle
    n = 1 << log_n
    rows = _rand(O, fid, 77, n).reshape(1, n, -1)
    enc = P.LigeroEncoding(fid, n // 2, n)
    got = rows.copy()
    enc.encode(got)
    assert np.array_equal(got, O.fft_io(fid, rows))


@pytest.mark.parametrize("fid", FIELDS)
@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 5, 7, 9, 10, 11, 12, 13, 14, 16])
def test_ligero_decode_rows_match_ifft_oi(P, oracle, fid, log_n):
    """decode_row == fffft ifft_oi on whole rows (proof-of-storage/src/lcpc_online.rs:568-573), and it undoes encode."""
    O = oracle
    n = 1 << log_n
    n_rows = 3 if log_n < 14 else 2
    rows = _rand(O, fid, 2000 + log_n, n_rows * n).reshape(n_rows, n, -1)
    enc = P.LigeroEncoding(fid, max(1, n // 2), n)
    got = rows.copy()
    enc.decode(got)
    assert np.array_equal(got, O.ifft_oi(fid, rows))
    enc.encode(got)
    assert np.array_equal(got, rows)          # encode(decode(y)) == y
    enc.encode(got)
    enc.decode(got)
    assert np.array_equal(got, rows)          # decode(encode(x)) == x


@pytest.mark.parametrize("fid,log_n", [(0, 17), (0, 18), (0, 20), (1, 17), (3, 15), (4, 15)])
def test_ligero_decode_large_rows(P, oracle, fid, log_n):
    """Rows with strided passes in front of the block pass (2^20: the shape the fused two-pass encode kernel takes)."""
    O = oracle
    n = 1 << log_n
    rows = _rand(O, fid, 78, n).reshape(1, n, -1)
    enc = P.LigeroEncoding(fid, n // 2, n)
    got = rows.copy()
    enc.decode(got)
    assert np.array_equal(got, O.ifft_oi(fid, rows))
    enc.encode(got)
    assert np.array_equal(got, rows)


def test_decode_custom_root_and_errors(P, oracle):
    O = oracle
    fid, log_n = 0, 10
    n = 1 << log_n
    rows = _rand(O, fid, 6, n).reshape(1, n, -1)
    enc = P.LigeroEncoding(fid, n // 2, n, root_of_unity=O.ntt_root(fid, log_n))
    got = rows.copy()
    enc.decode(got)
    assert np.array_equal(got, O.ifft_oi(fid, rows))
    senc = P.SdigEncoding.new(0, 3000, seed=0)
    with pytest.raises(P.ProverError) as ei:
        senc.decode(np.zeros((1, senc.n_cols, 1), dtype=np.uint64))
    assert ei.value.variant == "Encode"


def test_ligero_custom_root_of_unity(P, oracle):
    """The plan takes the n-th root from the caller (Rust passes F::ROOT_OF_UNITY.pow(2^(S-k)))."""
    O = oracle
    fid, log_n = 0, 10
    n = 1 << log_n
    rows = _rand(O, fid, 5, n).reshape(1, n, -1)
    enc = P.LigeroEncoding(fid, n // 2, n, root_of_unity=O.ntt_root(fid, log_n))
    got = rows.copy()
    enc.encode(got)
    assert np.array_equal(got, O.fft_io(fid, rows))


# ----------------------------------------------------------------------------- commit (K1+K3+K4)

COMMIT_CASES = [
    # fid, n_coeffs, n_per_row, n_cols
    (0, 1, 1, 2),            # smallest possible
    (0, 37, 5, 8),           # ragged last row, tiny tree
    (0, 3000, 100, 256),     # n_per_row not a power of two (lcpc-2d tests.rs:51)
    (0, 1 << 12, 512, 1024), # 8 rows: single-chunk leaves
    (0, 124 * 64, 64, 128),  # 124 rows: leaf = exactly one full chunk (1024 B)
    (0, 125 * 64, 64, 128),  # 125 rows: 1 chunk + 8 bytes
    (0, 700 * 32, 32, 64),   # 700 rows: 6 chunks, non power of two chunk count
    (0, 1 << 16, 2048, 4096),  # BASELINE configs[0] shape
    (1, 5000, 100, 256),
    (1, 130 * 64, 64, 128),
    (2, 3000, 60, 128),      # 24-byte elements straddle BLAKE3 blocks
    (2, 90 * 16, 16, 32),
    (3, 5000, 100, 256),
    (3, 70 * 64, 64, 128),
    # large enough for the chunk-pipelined host path (>= 8 rows and >= 8 MiB encoded), ragged tail
    (0, 300 * 2048 - 777, 2048, 4096),
    (0, 9 * 65536 - 1, 65536, 131072),
    (3, 75 * 2048 - 5, 2048, 4096),
    # strided first passes whose valid prefix is not half the row: non power-of-two n_per_row (bounds-checked loads),
    # rate 1/4, and the fused two-pass kernel with a ragged prefix
    (0, 5 * 20000 - 7, 20000, 65536),
    (0, 3 * 16384, 16384, 65536),
    (0, 2 * 50000 - 1, 50000, 131072),
    (0, 2 * 65536, 65536, 262144),
    # many BLAKE3 chunks per column: the level-parallel chunk merge (8+ chunks), even / odd / non power-of-two counts
    (0, 2000 * 64, 64, 128),        # 16 chunks
    (0, 1100 * 32 - 5, 32, 64),     # 9 chunks: the last one is carried up four levels
    (0, 5000 * 16, 16, 32),         # 40 chunks
    (1, 700 * 16, 16, 32),          # 16-byte elements, 11 chunks
    (3, 300 * 16, 16, 32),          # 32-byte elements, 10 chunks
    # a million multi-limb coefficients each: ~2e7 Montgomery products per case through every NTT pass shape
    (1, (1 << 20) - 3, 16384, 32768),
    (3, (1 << 20) - 3, 16384, 32768),
    # Ft253_192 (proof-of-storage/src/fields/ft253_192.rs): big-endian repr in the leaves, modulus words 0 / 0xffffffff
    (4, 1, 1, 2),
    (4, 5000, 100, 256),
    (4, 70 * 64, 64, 128),          # 3 chunks
    (4, 300 * 16, 16, 32),          # 10 chunks
    (4, 75 * 2048 - 5, 2048, 4096), # chunk-pipelined host path
    (4, (1 << 18) - 3, 8192, 16384),
]


@pytest.mark.parametrize("fid,n,n_per_row,n_cols", COMMIT_CASES)
def test_ligero_commit_matches_oracle(P, oracle, fid, n, n_per_row, n_cols):
    O = oracle
    coeffs = _rand(O, fid, n + fid, n)
    exp = O.commit(coeffs, O.LigeroEncoding(fid, n_per_row, n_cols))
    enc = P.LigeroEncoding(fid, n_per_row, n_cols)
    got = P.LcCommit.commit(coeffs, enc)
    assert (got.n_rows, got.n_per_row, got.n_cols) == (exp.n_rows, exp.n_per_row, exp.n_cols)
    assert np.array_equal(got.coeffs, exp.coeffs)
    assert np.array_equal(got.comm, exp.comm)
    assert np.array_equal(got.hashes, exp.hashes)
    assert got.get_root() == exp.get_root()


def test_commit_default_dims_2_16(P, oracle):
    """BASELINE configs[0]: LigeroEncoding::new(2^16) -> 32 x 2048 -> 4096 over Ft63."""
    O = oracle
    n = 1 << 16
    enc = P.LigeroEncoding.new(0, n)
    assert (enc.n_per_row, enc.n_cols) == (2048, 4096)
    assert enc.get_n_col_opens() == 309 and enc.get_n_degree_tests() == 3
    coeffs = _rand(O, 0, 1, n)
    got = P.LcCommit.commit(coeffs, enc)
    exp = O.commit(coeffs, O.LigeroEncoding.new(0, n))
    assert got.get_root() == exp.get_root()
    assert np.array_equal(got.hashes, exp.hashes)


def test_commit_device_resident_handle(P, oracle):
    """download=False keeps everything in HBM; root, lazy fields and folds still agree."""
    O = oracle
    coeffs = _rand(O, 0, 3, 5000)
    enc = P.LigeroEncoding(0, 100, 256)
    got = P.LcCommit.commit(coeffs, enc, download=False)
    exp = O.commit(coeffs, O.LigeroEncoding(0, 100, 256))
    assert got.get_root() == exp.get_root()
    assert np.array_equal(got.comm, exp.comm)


def test_commit_errors(P, oracle):
    enc = P.LigeroEncoding(0, 4, 8)
    with pytest.raises(AssertionError):
        P.LcCommit.commit(np.zeros((0, 1), dtype=np.uint64), enc)
    with pytest.raises(AssertionError):
        P.LigeroEncoding(0, 8, 8)  # n_per_row < n_cols
    with pytest.raises(AssertionError):
        P.LigeroEncoding(0, 3, 12)  # power of two
    with pytest.raises(P.LcpcError):
        P.LigeroEncoding(0, 1, 1 << 42)  # beyond the 2-adicity of Ft63


# ----------------------------------------------------------------------------- fold / open (K5, K6)

@pytest.mark.parametrize("fid", FIELDS)
def test_fold_and_open_match_oracle(P, oracle, fid):
    O = oracle
    n, n_per_row, n_cols = 6000, 100, 256
    coeffs = _rand(O, fid, 21, n)
    exp = O.commit(coeffs, O.LigeroEncoding(fid, n_per_row, n_cols))
    got = P.LcCommit.commit(coeffs, P.LigeroEncoding(fid, n_per_row, n_cols))
    tensors = _rand(O, fid, 22, 5 * exp.n_rows).reshape(5, exp.n_rows, -1)
    f = got.fold(tensors)
    fe = got.fold(tensors[:2], encoded=True)
    for t in range(5):
        assert np.array_equal(f[t], O.collapse_columns(fid, exp.coeffs, tensors[t]))
    for t in range(2):
        assert np.array_equal(fe[t], O.collapse_columns(fid, exp.comm, tensors[t]))
    assert np.array_equal(P.collapse_columns(got, tensors[0]), f[0])
    cols = [0, 1, 255, 17, 17, 128]
    opened = got.open_columns(cols)
    for c, col in zip(cols, opened):
        e = O.open_column(exp, c)
        assert np.array_equal(col.col, e.col)
        assert np.array_equal(col.path, e.path)
        assert O.verify_column_path(fid, O.LcColumn(col.col, col.path), c, got.get_root())
    assert np.array_equal(got.leaves(cols), exp.hashes[cols])
    with pytest.raises(P.ProverError) as ei:
        P.open_column(got, n_cols)
    assert ei.value.variant == "ColumnNumber"
    with pytest.raises(P.ProverError) as ei:
        P.collapse_columns(got, tensors[0][:-1])
    assert ei.value.variant == "OuterTensor"


def test_fold_many_rows_uses_row_splits(P, oracle):
    O = oracle
    fid, n_per_row, n_cols, n_rows = 0, 64, 128, 2000
    coeffs = _rand(O, fid, 31, n_rows * n_per_row)
    exp = O.commit(coeffs, O.LigeroEncoding(fid, n_per_row, n_cols))
    got = P.LcCommit.commit(coeffs, P.LigeroEncoding(fid, n_per_row, n_cols))
    tensors = _rand(O, fid, 32, 3 * n_rows).reshape(3, n_rows, -1)
    f = got.fold(tensors)
    for t in range(3):
        assert np.array_equal(f[t], O.collapse_columns(fid, exp.coeffs, tensors[t]))
    assert got.get_root() == exp.get_root()


# ----------------------------------------------------------------------------- Brakedown (K2)

def _to_pkg_csc(P, mats):
    return [P.CscMatrix(m.rows, m.cols, m.indptr, m.indices, m.data) for m in mats]


@pytest.mark.parametrize("fid,n_per_row,n_rows,seed", [(0, 150, 20, 0), (3, 150, 7, 1), (1, 400, 5, 0), (2, 120, 3, 1), (4, 150, 7, 1)])
def test_brakedown_commit_matches_oracle(P, oracle, fid, n_per_row, n_rows, seed):
    """SdigEncodingS::encode on every row + Merkle tree with a non power-of-two n_cols
    (padding leaves stay zero, lib.rs:685-695)."""
    O = oracle
    oenc = O.SdigEncoding(fid, n_per_row, seed)
    coeffs = _rand(O, fid, 41, n_rows * n_per_row - 3)
    exp = O.commit(coeffs, oenc)
    enc = P.SdigEncoding(fid, _to_pkg_csc(P, oenc.precodes), _to_pkg_csc(P, oenc.postcodes))
    assert enc.n_cols == oenc.n_cols and enc.get_n_degree_tests() == oenc.get_n_degree_tests()
    got = P.LcCommit.commit(coeffs, enc)
    assert np.array_equal(got.comm, exp.comm)
    assert np.array_equal(got.coeffs, exp.coeffs)
    assert np.array_equal(got.hashes, exp.hashes)
    # single-row encode, as verify uses it (lib.rs:912-918)
    row = np.zeros((1, enc.n_cols, O.LIMBS[fid]), dtype=np.uint64)
    row[0, :n_per_row] = exp.coeffs[1]
    enc.encode(row)
    assert np.array_equal(row[0], exp.comm[1])


# ----------------------------------------------------------------------------- proof-of-storage bytes

def _pack7(data: bytes) -> np.ndarray:
    """DataField::from_byte_vec for WriteableFt63 (data_field.rs:38-46, writable_ft63.rs:35-40)."""
    n = (len(data) + 6) // 7
    buf = np.zeros(n * 7, dtype=np.uint8)
    buf[:len(data)] = np.frombuffer(data, dtype=np.uint8)
    b = buf.reshape(n, 7).astype(np.uint64)
    out = np.zeros(n, dtype=np.uint64)
    for k in range(7):
        out |= b[:, k] << np.uint64(8 * k)
    return out.reshape(n, 1)


@pytest.mark.parametrize("n_bytes", [1, 6, 7, 8, 55, 56, 57, 598, 100003])
def test_commit_bytes_matches_oracle(P, oracle, n_bytes):
    O = oracle
    rng = np.random.default_rng(n_bytes)
    data = rng.integers(0, 256, n_bytes, dtype=np.uint8).tobytes()
    elems = _pack7(data)
    n_per_row, n_cols = (4, 8) if n_bytes < 1000 else (64, 128)
    exp = O.commit(elems, O.LigeroEncoding(0, n_per_row, n_cols))
    got = P.LcCommit.commit_bytes(data, P.LigeroEncoding(0, n_per_row, n_cols))
    assert np.array_equal(got.coeffs, exp.coeffs)
    assert np.array_equal(got.hashes, exp.hashes)


def _ft253_file(n_bytes: int, seed: int) -> bytes:
    """Random file bytes whose 31-byte groups are all below the Ft253_192 modulus (byte 24 of a group <= 0x1f)."""
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 256, n_bytes, dtype=np.uint8)
    data[24::31] &= 0x1F
    return data.tobytes()


@pytest.mark.parametrize("n_bytes", [1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 23, 24, 25, 30, 31, 32, 33, 38, 62, 63, 598, 100003])
def test_commit_bytes_ft253_192_matches_oracle(P, oracle, n_bytes):
    """Ft253_192::from_data_bytes (ft253_192.rs:18-30) fused in front of the commit: 31 bytes per element, big-endian limbs."""
    O = oracle
    data = _ft253_file(n_bytes, n_bytes)
    elems = O.pack_bytes31(data)
    n_per_row, n_cols = (4, 8) if n_bytes < 1000 else (64, 128)
    exp = O.commit(elems, O.LigeroEncoding(4, n_per_row, n_cols))
    got = P.LcCommit.commit_bytes(data, P.LigeroEncoding(4, n_per_row, n_cols))
    assert np.array_equal(got.coeffs, exp.coeffs)
    assert np.array_equal(got.comm, exp.comm)
    assert np.array_equal(got.hashes, exp.hashes)


def test_commit_bytes_ft253_192_refuses_unreduced_groups(P, oracle):
    """A group with byte 24 > 0x1f has limbs >= p; the reference keeps computing on them (not a field computation), the
    library refuses (include/lcpc_b200.h, lcpc_commit_bytes_host)."""
    data = bytearray(_ft253_file(31 * 40, 7))
    data[31 * 17 + 24] = 0x20
    with pytest.raises(ValueError):
        oracle.pack_bytes31(bytes(data))
    with pytest.raises(P.LcpcError):
        P.LcCommit.commit_bytes(bytes(data), P.LigeroEncoding(4, 4, 8))
    data[31 * 17 + 24] = 0x1F  # the largest top byte that is still below the modulus
    got = P.LcCommit.commit_bytes(bytes(data), P.LigeroEncoding(4, 4, 8))
    exp = oracle.commit(oracle.pack_bytes31(bytes(data)), oracle.LigeroEncoding(4, 4, 8))
    assert got.get_root() == exp.get_root()


# ----------------------------------------------------------------------------- full size

def test_ligero_commit_2_24_bit_exact(P, oracle):
    """BASELINE configs[1]: 2^24 coefficients over Ft63, 512 x 32768 -> 65536, bit-exact vs CPU."""
    O = oracle
    n = 1 << 24
    coeffs = _rand(O, 0, 2, n)
    enc = P.LigeroEncoding.new(0, n)
    assert (enc.n_per_row, enc.n_cols) == (32768, 65536)
    got = P.LcCommit.commit(coeffs, enc)
    exp = O.commit(coeffs, O.LigeroEncoding.new(0, n))
    assert got.get_root() == exp.get_root()
    assert np.array_equal(got.hashes, exp.hashes)
    assert np.array_equal(got.comm, exp.comm)
    # encode -> decode round trip (lcpc_online.rs:588-601) on a few rows
    back = O.ifft_oi(0, got.comm[:4])
    assert np.array_equal(back[:, :enc.n_per_row], got.coeffs[:4])
    assert not back[:, enc.n_per_row:].any()
