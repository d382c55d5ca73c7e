"""Host-side logic of the product (no GPU needed): transcript, challenge expansion, code
generation and dimension selection, each against the oracle's independent restatement and the
published vectors."""
import random

import numpy as np
import pytest

import lcpc_proof_of_storage_b200 as P
from lcpc_proof_of_storage_b200 import _lib


def test_transcript_kat_and_matches_oracle(oracle):
    t = P.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == (
        "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615")
    rnd = random.Random(3)
    a, b = P.Transcript(b"lcpc"), oracle.Transcript(b"lcpc")
    for i in range(300):
        label = bytes(rnd.getrandbits(8) for _ in range(rnd.randrange(0, 12)))
        if rnd.random() < 0.7:
            msg = bytes(rnd.getrandbits(8) for _ in range(rnd.choice([0, 1, 8, 32, 165, 166, 167, 400])))
            a.append_message(label, msg)
            b.append_message(label, msg)
        else:
            n = rnd.choice([1, 32, 64, 200])
            assert a.challenge_bytes(label, n) == b.challenge_bytes(label, n)
    c = a.clone()
    assert c.challenge_bytes(b"x", 32) == a.challenge_bytes(b"x", 32) == b.challenge_bytes(b"x", 32)


@pytest.mark.parametrize("fid", [0, 1, 2, 3, 4])
def test_challenge_expansion_matches_oracle(oracle, fid):
    lib = _lib.load()
    key = bytes((11 * i + fid) % 256 for i in range(32))
    L = P.FIELD_LIMBS[fid]
    out = np.zeros((500, L), dtype=np.uint64)
    _lib.check(lib.lcpc_random_field_vec(fid, key, out.ctypes.data, 500))
    assert np.array_equal(out, oracle.random_field_vec(fid, key, 500))
    for n_cols in (2, 4096, 65536, 252931):
        cols = np.zeros(309, dtype=np.uint64)
        _lib.check(lib.lcpc_random_columns(key, n_cols, cols.ctypes.data, 309))
        assert np.array_equal(cols, oracle.random_columns(key, n_cols, 309))
        assert int(cols.max()) < n_cols


@pytest.mark.parametrize("fid,n,seed,code", [(0, 150, 0, 3), (3, 150, 1, 3), (1, 1000, 7, 1), (2, 333, 2, 6), (0, 5000, 1, 4)])
def test_sdig_generation_matches_oracle(oracle, fid, n, seed, code):
    """matgen::generate: same dims, same CSC arrays bit for bit (ChaCha20 streams, rejection sampling,
    sorted indices, non-zero F::random values)."""
    pre, post = P.SdigEncoding.generate(fid, n, seed, code)
    opre, opost = oracle.sdig_generate(fid, code, n, seed)
    assert len(pre) == len(opre)
    for a, b in list(zip(pre, opre)) + list(zip(post, opost)):
        assert (a.rows, a.cols) == (b.rows, b.cols)
        assert np.array_equal(a.indptr, b.indptr)
        assert np.array_equal(a.indices, b.indices)
        assert np.array_equal(a.data, b.data)
    assert P.sdig_codeword_length(pre, post) == oracle.sdig_codeword_length(opre, opost)
    # structure: every column has d distinct sorted rows, values non-zero and < p
    for m in pre + post:
        d = int(m.indptr[1]) if m.cols else 0
        assert np.array_equal(m.indptr, np.arange(m.cols + 1, dtype=np.uint64) * d)
        idx = m.indices.reshape(m.cols, d).astype(np.int64)
        assert (np.diff(idx, axis=1) > 0).all() and idx.max() < m.rows
        assert m.data.any(axis=1).all()


def test_dimension_selection_matches_survey_tables():
    L = P.LigeroEncoding
    assert L._get_dims(0, 1 << 16) == (32, 2048, 4096)
    assert L._get_dims(0, 1 << 20) == (128, 8192, 16384)
    assert L._get_dims(0, 1 << 24) == (512, 32768, 65536)
    assert L._get_dims(0, 1 << 28) == (2048, 131072, 262144)
    assert L._get_dims(3, 1 << 24) == (256, 65536, 131072)
    assert L._get_dims(3, 1 << 24, 1, 4) == (512, 32768, 131072)
    assert L._n_col_opens(1, 2) == 309 and L._n_col_opens(1, 4) == 189
    assert L._n_degree_tests(0, 65536) == 3 and L._n_degree_tests(3, 131072) == 1
    S = P.SdigEncoding
    assert S._n_col_opens(3) == 6593
    assert S._n_per_row_for_len(3, 1 << 24) == 166292
    assert S._n_per_row_for_len(0, 1 << 24) == 166293
    assert S._n_per_row_for_len(1, 1 << 24) == 235173


def test_ligero_dims_invariants_random_lengths():
    """lcpc-ligero-pc/src/tests.rs:22-41 (get_dims over random lengths)."""
    rnd = random.Random(5)
    for _ in range(2000):
        length = rnd.randrange(1, 1 << rnd.randrange(4, 30))
        n_rows, n_per_row, n_cols = P.LigeroEncoding._get_dims(0, length)
        assert n_rows * n_per_row >= length
        assert (n_rows - 1) * n_per_row < length
        assert n_per_row * 2 == n_cols and P.LigeroEncoding._dims_ok(n_per_row, n_cols)


def test_log2_and_degree_tests():
    for i in range(31):  # lcpc-2d/src/tests.rs:127-134
        assert P.log2(1 << i) == i
    assert P.log2(3) == 2 and P.log2(5) == 3 and P.log2(252931) == 18
    assert P.n_degree_tests(128, 65536, 62) == 3
