"""BASELINE.json's full sizes, checked through size-independent properties (the CPU oracle would take minutes and tens
of GiB here).  All bit-exact, all through the C ABI:

  * the scheme's own invariant, which `verify` relies on (lcpc-2d/src/lib.rs:912-975): for any tensor t,
    encode(fold(coeffs, t)) == fold(comm, t) -- linearity of the code and of both folds, over the whole matrix;
  * opened columns carry valid Merkle paths to the root, checked by the oracle's verify_column_path on the CPU, and equal
    the matching entries of the encoded fold;
  * the streamed commit (chunk-at-a-time column digests, lcpc_stream_*) of the same input gives the same tree as the
    resident commit (different hashing code path): proof-of-storage's "streamed root == in-memory root"
    (row_generator_iter.rs:286-364);
  * encode -> decode round trip on sampled rows (oracle ifft_oi on the CPU, lcpc_online.rs:588-601).
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import lcpc_proof_of_storage_b200 as pkg

    return pkg


P_TOP = {0: 0x46d0760000000001, 1: 0x6e754097ba20e0bf, 2: 0x453708aa3fbc8dda, 3: 0x663c799b6e4d2900}


def _fast_rand(fid, n, seed, limbs):
    """Reduced field elements straight from numpy (the oracle's seeded generator is a Python loop: minutes at 2^28):
    every limb uniform, the top limb below the modulus' top limb, so the value is below p."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 64, size=(n, limbs), dtype=np.uint64)
    a[:, limbs - 1] = rng.integers(0, P_TOP[fid], size=n, dtype=np.uint64)
    return a


def _check_invariants(P, O, enc, c, coeffs, n_tensors=2, decode=True):
    fid, L = enc.fid, enc.limbs
    tensors = O.random_field_elements(fid, 4242, n_tensors * c.n_rows).reshape(n_tensors, c.n_rows, L)
    folded = c.fold(tensors)                       # over the unencoded coefficients
    folded_enc = c.fold(tensors, encoded=True)     # over the encoded matrix
    rows = np.zeros((n_tensors, c.n_cols, L), dtype=np.uint64)
    rows[:, :c.n_per_row] = folded
    enc.encode(rows)
    assert np.array_equal(rows, folded_enc), "encode(fold(coeffs)) != fold(encode(coeffs))"
    root = c.get_root()
    cols = [0, 1, c.n_cols - 1, c.n_cols // 2, c.n_cols // 3, 12345 % c.n_cols]
    for col_idx, col in zip(cols, c.open_columns(cols)):
        assert O.verify_column_path(fid, O.LcColumn(np.ascontiguousarray(col.col), np.ascontiguousarray(col.path)), col_idx, root)
        for t in range(n_tensors):  # the opened column dotted with the tensor is the encoded fold at that column
            dot = O.collapse_columns(fid, np.ascontiguousarray(col.col).reshape(c.n_rows, 1, L), tensors[t])
            assert np.array_equal(dot[0], folded_enc[t, col_idx])
    if decode:
        # sampled rows decode back to the zero-padded coefficients
        sample = [0, c.n_rows // 2, c.n_rows - 1]
        enc_rows = np.zeros((len(sample), c.n_cols, L), dtype=np.uint64)
        flat = np.zeros((c.n_rows * c.n_per_row, L), dtype=np.uint64)
        flat[:coeffs.shape[0]] = coeffs
        for k, r in enumerate(sample):
            enc_rows[k, :c.n_per_row] = flat[r * c.n_per_row:(r + 1) * c.n_per_row]
        enc.encode(enc_rows)
        back = O.ifft_oi(fid, enc_rows)
        for k, r in enumerate(sample):
            assert np.array_equal(back[k, :c.n_per_row], flat[r * c.n_per_row:(r + 1) * c.n_per_row])
            assert not back[k, c.n_per_row:].any()
    return root


def _stream_root(P, enc, coeffs, n_rows, chunk_rows):
    from lcpc_proof_of_storage_b200 import _lib

    lib = _lib.load()
    s = C.c_void_p()
    _lib.check(lib.lcpc_stream_begin(enc.plan, n_rows, 0, None, 0, C.byref(s)))
    try:
        step = chunk_rows * enc.n_per_row
        for off in range(0, coeffs.shape[0], step):
            part = np.ascontiguousarray(coeffs[off:off + step])
            _lib.check(lib.lcpc_stream_push_elems_host(s, part.ctypes.data, part.shape[0]))
        hashes = np.empty((2 * P.next_pow2(enc.n_cols) - 1, 32), dtype=np.uint8)
        _lib.check(lib.lcpc_stream_finish(s, hashes.ctypes.data, None))
        return hashes
    finally:
        lib.lcpc_stream_free(s)


def test_ligero_ft63_2_28_invariants(P, oracle):
    """The target size of BASELINE.json: 2^28 coefficients, 2048 x 131072 -> 262144 (2 GiB in, 4 GiB encoded)."""
    O = oracle
    n = (1 << 28) - 12345  # ragged last row
    enc = P.LigeroEncoding.new(0, n)
    assert (enc.n_per_row, enc.n_cols) == (131072, 262144)
    coeffs = _fast_rand(0, n, 28, 1)
    c = P.LcCommit.commit(coeffs, enc, download=False)
    assert c.n_rows == 2048
    _check_invariants(P, O, enc, c, coeffs)
    streamed = _stream_root(P, enc, coeffs, c.n_rows, 300)
    assert streamed[-1].tobytes() == c.get_root()
    leaves = c.leaves([0, 7, 262143])
    assert np.array_equal(leaves, streamed[[0, 7, 262143]])


def _check_rows_and_leaves_against_oracle(P, O, oenc, c, coeffs, sample_rows, sample_cols):
    """Rows of the resident encoded matrix against the oracle's encode of the same coefficient rows (a wrong-but-linear
    encode passes every invariant above: this does not), and the leaves / opened columns of sampled columns against the
    oracle's hash of the same column.  A row of `comm` is read back as the encoded-matrix fold by a unit tensor."""
    fid, L = c.enc.fid, c.enc.limbs
    flat = np.zeros((c.n_rows * c.n_per_row, L), dtype=np.uint64)
    flat[:coeffs.shape[0]] = coeffs
    one = O.to_mont(fid, [1])[0]
    units = np.zeros((len(sample_rows), c.n_rows, L), dtype=np.uint64)
    for k, r in enumerate(sample_rows):
        units[k, r] = one
    got_rows = c.fold(units, encoded=True)                      # [k, n_cols, L] == comm[sample_rows]
    msg = np.zeros((len(sample_rows), c.n_cols, L), dtype=np.uint64)
    for k, r in enumerate(sample_rows):
        msg[k, :c.n_per_row] = flat[r * c.n_per_row:(r + 1) * c.n_per_row]
    exp_rows = oenc.encode_rows(msg)
    assert np.array_equal(got_rows, exp_rows), "encoded rows differ from the oracle's encode"
    opened = c.open_columns(sample_cols)
    leaves = c.leaves(sample_cols)
    for k, (j, col) in enumerate(zip(sample_cols, opened)):
        colv = np.ascontiguousarray(col.col).reshape(c.n_rows, L)
        for i, r in enumerate(sample_rows):
            assert np.array_equal(colv[r], exp_rows[i, j])
        assert O.hash_column(fid, colv) == leaves[k].tobytes(), "leaf differs from the oracle's column hash"


def test_brakedown_ft255_2_24_invariants(P, oracle):
    """BASELINE configs[2]: Brakedown code 3 over the 255-bit field, 101 x 166292 -> 252931; rows 0 / 50 / 100 of the
    encoded matrix and six leaves against the oracle (encode.rs:36-94, lib.rs:736-775)."""
    O = oracle
    n = 1 << 24
    enc = P.SdigEncoding.new(3, n, seed=0)
    assert (enc.n_per_row, enc.n_cols) == (166292, 252931)
    coeffs = _fast_rand(3, n, 24, 4)
    c = P.LcCommit.commit(coeffs, enc, download=False)
    assert c.n_rows == 101
    _check_invariants(P, O, enc, c, coeffs, decode=False)
    oenc = O.SdigEncoding(3, enc.n_per_row, 0)
    assert oenc.n_cols == enc.n_cols
    _check_rows_and_leaves_against_oracle(P, O, oenc, c, coeffs, [0, 50, 100],
                                          [0, 166291, 166292, 200000, 252930, 12345])


def test_brakedown_ft63_2_24_rows_and_leaves(P, oracle):
    """BASELINE configs[4]'s Brakedown point over the 63-bit field: 101 x 166293 -> 252932."""
    O = oracle
    n = 1 << 24
    enc = P.SdigEncoding.new(0, n, seed=0)
    assert (enc.n_per_row, enc.n_cols) == (166293, 252932)
    coeffs = _fast_rand(0, n - 77, 26, 1)  # ragged last row
    c = P.LcCommit.commit(coeffs, enc, download=False)
    assert c.n_rows == 101
    _check_invariants(P, O, enc, c, coeffs, decode=False)
    oenc = O.SdigEncoding(0, enc.n_per_row, 0)
    _check_rows_and_leaves_against_oracle(P, O, oenc, c, coeffs, [0, 50, 100],
                                          [0, 166292, 166293, 200000, 252931, 54321])


def test_ligero_ft255_2_24_invariants(P, oracle):
    O = oracle
    n = 1 << 24
    enc = P.LigeroEncoding.new(3, n)
    assert (enc.n_per_row, enc.n_cols) == (65536, 131072)
    coeffs = _fast_rand(3, n, 25, 4)
    c = P.LcCommit.commit(coeffs, enc, download=False)
    root = _check_invariants(P, O, enc, c, coeffs)
    streamed = _stream_root(P, enc, coeffs, c.n_rows, 50)
    assert streamed[-1].tobytes() == root


@pytest.mark.parametrize("fid,n_rows", [(0, 5), (0, 17), (0, 40), (0, 48), (0, 64), (0, 130), (1, 17), (2, 48), (3, 9), (3, 40), (4, 17)])
def test_brakedown_wide_levels_every_lane_group_shape(P, oracle, fid, n_rows):
    """The wide expander levels at a size where they take the many-warp kernels (k_spmv_tg for the one-limb field with
    NG = 2 ... 7 lane groups per thread and lane groups of 8 / 16 / 32 matrix rows; the pipelined k_spmv_t with split /
    Karatsuba accumulators for the others), for matrix heights that give every lane-group size and ragged last groups:
    EVERY encoded row against the oracle's encode (encode.rs:36-94), through both entry points -- the commit (message
    taken from the coefficient matrix, copied into comm by the transposing pass) and lcpc_encode_rows (in place)."""
    O = oracle
    L = P.FIELD_LIMBS[fid]
    n_per_row = 60000 + fid  # level 0: 60000 -> ~10700 outputs: thousands of warps
    enc = P.SdigEncoding.new_from_dims(fid, n_per_row, None, seed=3)
    oenc = O.SdigEncoding(fid, n_per_row, 3)
    assert oenc.n_cols == enc.n_cols
    if fid == 4:
        coeffs = _fast_rand(3, n_rows * n_per_row - 11, 31 + n_rows, L)
        coeffs[:, L - 1] &= np.uint64((1 << 56) - 1)  # below the 253-bit modulus
    else:
        coeffs = _fast_rand(fid, n_rows * n_per_row - 11, 31 + n_rows, L)
    c = P.LcCommit.commit(coeffs, enc, download=True)
    assert c.n_rows == n_rows
    flat = np.zeros((n_rows * n_per_row, L), dtype=np.uint64)
    flat[:coeffs.shape[0]] = coeffs
    msg = np.zeros((n_rows, enc.n_cols, L), dtype=np.uint64)
    msg[:, :n_per_row] = flat.reshape(n_rows, n_per_row, L)
    exp = oenc.encode_rows(msg.copy())
    assert np.array_equal(c.comm.reshape(n_rows, enc.n_cols, L), exp), "commit: encoded matrix differs from the oracle's"
    rows = msg.copy()
    enc.encode(rows)
    assert np.array_equal(rows, exp), "encode_rows: differs from the oracle's"
    cols = [0, n_per_row - 1, n_per_row, enc.n_cols - 1]
    leaves = c.leaves(cols)
    for k, j in enumerate(cols):
        assert O.hash_column(fid, np.ascontiguousarray(exp[:, j])) == leaves[k].tobytes()
