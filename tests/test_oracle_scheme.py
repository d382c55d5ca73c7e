"""The reference's own test properties, run on the CPU oracle.

These mirror lcpc-2d/src/tests.rs, lcpc-ligero-pc/src/tests.rs and
lcpc-brakedown-pc/src/tests.rs: every reference test is differential or round-trip, so
the oracle has to satisfy the same identities before it is trusted as the checker for
the CUDA path.
"""
import random

import numpy as np
import pytest

blake3_pkg = pytest.importorskip("blake3")


def _py_leaf(O, fid, comm, j):
    """Serial statement of the leaf hash (lcpc-2d/src/lib.rs:1166-1173 merkleize_ser)."""
    w = 8 * O.LIMBS[fid]
    h = blake3_pkg.blake3()
    h.update(bytes(32))
    for v in O.from_mont(fid, comm[:, j]):
        h.update(v.to_bytes(w, "big" if O.REPR_BIG_ENDIAN[fid] else "little"))  # PrimeFieldReprEndianness
    return h.digest()


def _py_tree(leaves):
    """Serial Merkle tree (lib.rs:1175-1188)."""
    levels = [list(leaves)]
    while len(levels[-1]) > 1:
        cur = levels[-1]
        levels.append([blake3_pkg.blake3(cur[2 * i] + cur[2 * i + 1]).digest() for i in range(len(cur) // 2)])
    return [h for lvl in levels for h in lvl]


@pytest.mark.parametrize("fid", [0, 1, 2, 3, 4])
def test_merkleize_matches_serial(oracle, fid):
    """T1: lcpc-2d/src/tests.rs:136-149 (parallel merkleize == merkleize_ser)."""
    O = oracle
    n_rows, n_cols = 5 + 40 * (fid == 0), 48  # non power of two: padding leaves stay zero
    comm = O.random_field_elements(fid, 11 + fid, n_rows * n_cols).reshape(n_rows, n_cols, -1)
    leaves = O.hash_columns(fid, comm)
    exp_leaves = [_py_leaf(O, fid, comm, j) for j in range(n_cols)]
    assert [l.tobytes() for l in leaves] == exp_leaves
    np2 = O.next_pow2(n_cols)
    padded = np.zeros((np2, 32), dtype=np.uint8)
    padded[:n_cols] = leaves
    tree = O.merkle_tree(padded)
    exp = _py_tree(exp_leaves + [bytes(32)] * (np2 - n_cols))
    assert [t.tobytes() for t in tree] == exp


@pytest.mark.parametrize("fid", [0, 3, 4])
def test_collapse_columns_matches_ints(oracle, fid):
    """T2: lcpc-2d/src/tests.rs:151-165 (collapse_columns == eval_outer_ser)."""
    O = oracle
    p = O.MODULUS[fid]
    n_rows, n_per_row = 9, 70
    coeffs = O.random_field_elements(fid, 3, n_rows * n_per_row).reshape(n_rows, n_per_row, -1)
    tensor = O.random_field_elements(fid, 4, n_rows)
    got = O.from_mont(fid, O.collapse_columns(fid, coeffs, tensor))
    c = [O.from_mont(fid, coeffs[r]) for r in range(n_rows)]
    t = O.from_mont(fid, tensor)
    assert got == [sum(c[r][j] * t[r] for r in range(n_rows)) % p for j in range(n_per_row)]


def test_open_column_paths_verify(oracle):
    """T3: lcpc-2d/src/tests.rs:167-191 (64 random columns verify against the root)."""
    O = oracle
    fid = 0
    coeffs = O.random_field_elements(fid, 5, 3000)
    enc = O.LigeroEncoding(fid, 100, 256)
    comm = O.commit(coeffs, enc)
    rnd = random.Random(1)
    for _ in range(64):
        c = rnd.randrange(comm.n_cols)
        col = O.open_column(comm, c)
        assert np.array_equal(col.col, comm.comm[:, c])
        assert O.verify_column_path(fid, col, c, comm.get_root())
        assert not O.verify_column_path(fid, col, c ^ 1, comm.get_root())
    with pytest.raises(O.ProverError):
        O.open_column(comm, comm.n_cols)


@pytest.mark.parametrize("fid", [0, 1])
def test_commit_reed_solomon_structure(oracle, fid):
    """T4: lcpc-2d/src/tests.rs:193-234 -- fold of the encoded rows decodes (ifft_oi) to a
    polynomial of degree < n_per_row that evaluates like the committed one."""
    O = oracle
    p = O.MODULUS[fid]
    length = 1000
    coeffs = O.random_field_elements(fid, 6, length)
    enc = O.LigeroEncoding(fid, 37, 128)  # n_per_row need not be a power of two (tests.rs:51)
    comm = O.commit(coeffs, enc)
    assert (comm.n_rows, comm.n_per_row, comm.n_cols) == (28, 37, 128)
    x = 987654321 % p
    cf = O.from_mont(fid, comm.coeffs.reshape(-1, O.LIMBS[fid]))
    direct = sum(c * pow(x, i, p) for i, c in enumerate(cf)) % p
    roots_lo = [pow(x, j, p) for j in range(comm.n_per_row)]
    xr = pow(x, comm.n_per_row, p)
    roots_hi = O.to_mont(fid, [pow(xr, i, p) for i in range(comm.n_rows)])
    flat = O.from_mont(fid, O.collapse_columns(fid, comm.coeffs, roots_hi))
    assert sum(c * r for c, r in zip(flat, roots_lo)) % p == direct
    folded_fft = O.collapse_columns(fid, comm.comm, roots_hi)
    poly = O.from_mont(fid, O.ifft_oi(fid, folded_fft.reshape(1, comm.n_cols, -1))[0])
    assert all(v == 0 for v in poly[comm.n_per_row:])
    assert poly[:comm.n_per_row] == flat


def _tensors(O, fid, comm, x):
    p = O.MODULUS[fid]
    inner = O.to_mont(fid, [pow(x, j, p) for j in range(comm.n_per_row)])
    xr = pow(x, comm.n_per_row, p)
    outer = O.to_mont(fid, [pow(xr, i, p) for i in range(comm.n_rows)])
    return outer, inner


def _eval(O, fid, coeffs, x):
    p = O.MODULUS[fid]
    return sum(c * pow(x, i, p) for i, c in enumerate(O.from_mont(fid, coeffs))) % p


@pytest.mark.parametrize("fid,length", [(0, 1 << 12), (3, 1 << 10)])
def test_ligero_end_to_end(oracle, fid, length):
    """T5: lcpc-ligero-pc/src/tests.rs:216-283 (verify returns the true evaluation)."""
    O = oracle
    coeffs = O.random_field_elements(fid, 7, length)
    enc = O.LigeroEncoding.new(fid, length)
    comm = O.commit(coeffs, enc)
    x = 0x1234567 + fid
    outer, inner = _tensors(O, fid, comm, x)
    tr1 = O.Transcript(b"test transcript")
    tr1.append_message(b"polycommit", comm.get_root())
    tr1.append_message(b"ncols", (enc.get_n_col_opens()).to_bytes(8, "big"))
    pf = O.prove(comm, outer, enc, tr1)
    assert len(pf.columns) == enc.get_n_col_opens() == 309
    tr2 = O.Transcript(b"test transcript")
    tr2.append_message(b"polycommit", comm.get_root())
    tr2.append_message(b"ncols", (enc.get_n_col_opens()).to_bytes(8, "big"))
    enc2 = O.LigeroEncoding(fid, pf.p_eval.shape[0], pf.n_cols)
    res = O.verify(comm.get_root(), outer, inner, pf, enc2, tr2)
    assert O.from_mont(fid, res)[0] == _eval(O, fid, coeffs, x)


def test_two_proofs_share_a_transcript(oracle):
    """lcpc-2d/src/tests.rs:318-413: prover and verifier transcripts stay in sync over two proofs."""
    O = oracle
    fid = 0
    coeffs = O.random_field_elements(fid, 8, 2000)
    enc = O.LigeroEncoding(fid, 50, 128, n_col_opens=20, n_degree_tests_=2)
    comm = O.commit(coeffs, enc)
    tr1, tr2 = O.Transcript(b"t"), O.Transcript(b"t")
    for tr in (tr1, tr2):
        tr.append_message(b"polycommit", comm.get_root())
    xs = (17, 99)
    proofs = []
    for x in xs:
        outer, inner = _tensors(O, fid, comm, x)
        proofs.append(O.prove(comm, outer, enc, tr1))
    for x, pf in zip(xs, proofs):
        outer, inner = _tensors(O, fid, comm, x)
        res = O.verify(comm.get_root(), outer, inner, pf, enc, tr2)
        assert O.from_mont(fid, res)[0] == _eval(O, fid, comm.coeffs.reshape(-1, 1), x)
    assert tr1.challenge_bytes(b"x", 16) == tr2.challenge_bytes(b"x", 16)


def test_verify_rejects_tampering(oracle):
    """T8 analogue (networking/tests.rs:696-780): a flipped bit in an opened column, a path node,
    or p_eval must be rejected with the reference's error variant."""
    O = oracle
    fid = 0
    coeffs = O.random_field_elements(fid, 9, 2000)
    enc = O.LigeroEncoding(fid, 50, 128, n_col_opens=16, n_degree_tests_=1)
    comm = O.commit(coeffs, enc)
    outer, inner = _tensors(O, fid, comm, 5)

    def run(mutate):
        tr1 = O.Transcript(b"t")
        pf = O.prove(comm, outer, enc, tr1)
        mutate(pf)
        O.verify(comm.get_root(), outer, inner, pf, enc, O.Transcript(b"t"))

    run(lambda pf: None)
    with pytest.raises(O.VerifierError, match="ColumnDegree"):
        run(lambda pf: pf.columns[3].col.__setitem__((0, 0), pf.columns[3].col[0, 0] ^ np.uint64(1)))
    with pytest.raises(O.VerifierError, match="ColumnPath"):
        run(lambda pf: pf.columns[2].path.__setitem__((1, 0), pf.columns[2].path[1, 0] ^ np.uint8(1)))
    with pytest.raises(O.VerifierError, match="NumColOpens"):
        run(lambda pf: pf.columns.pop())


@pytest.mark.parametrize("fid,seed", [(0, 0), (3, 1)])
def test_brakedown_end_to_end(oracle, fid, seed):
    """lcpc-brakedown-pc/src/tests.rs:192-375 (seeds 0 and 1); also the codeword layout
    asserts of encode.rs:42,92-93 through sdig_codeword_length."""
    O = oracle
    length = 3000
    enc = O.SdigEncoding(fid, 150, seed)
    assert enc.n_cols == O.sdig_codeword_length(enc.precodes, enc.postcodes)
    coeffs = O.random_field_elements(fid, 10, length)
    comm = O.commit(coeffs, enc)
    assert comm.n_rows == 20 and comm.n_cols == enc.n_cols
    assert np.array_equal(comm.comm[:, :150], comm.coeffs)  # systematic code
    # linearity of the encoding: enc(a) + enc(b) == enc(a + b)
    a, b = comm.coeffs[0:1], comm.coeffs[1:2]
    pad = lambda v: np.concatenate([v, np.zeros((1, enc.n_cols - 150, O.LIMBS[fid]), np.uint64)], axis=1)
    ea, eb = enc.encode_rows(pad(a)), enc.encode_rows(pad(b))
    eab = enc.encode_rows(pad(O.fe_add(fid, a, b)))
    assert np.array_equal(O.fe_add(fid, ea, eb), eab)
    x = 31337
    outer, inner = _tensors(O, fid, comm, x)
    # keep the proof small: the soundness-level 6593 openings are not needed for the identity
    enc.get_n_col_opens = lambda: 40
    pf = O.prove(comm, outer, enc, O.Transcript(b"bd"))
    res = O.verify(comm.get_root(), outer, inner, pf, enc, O.Transcript(b"bd"))
    assert O.from_mont(fid, res)[0] == _eval(O, fid, coeffs, x)


def test_dimension_tables_match_survey(oracle):
    """SURVEY.md section 8 shape table (restated _get_dims / matgen::get_dims)."""
    O = oracle
    L = O.LigeroEncoding
    assert L.get_dims_for_len(0, 1 << 16) == (32, 2048, 4096)
    assert L.get_dims_for_len(0, 1 << 24) == (512, 32768, 65536)
    assert L.get_dims_for_len(0, 1 << 28) == (2048, 131072, 262144)
    assert L.get_dims_for_len(3, 1 << 24) == (256, 65536, 131072)
    assert L.get_dims_for_len(3, 1 << 24, 1, 4) == (512, 32768, 131072)
    assert L._n_col_opens(1, 2) == 309 and L._n_col_opens(1, 4) == 189
    assert L._n_degree_tests(0, 65536) == 3 and L._n_degree_tests(3, 131072) == 1
    S = O.SdigEncoding
    assert S._n_col_opens(3) == 6593
    npr = S.n_per_row_for_len(3, 1 << 24)
    pre, post = O.sdig_get_dims(3, npr, 254)
    assert npr == 166292 and [d for *_, d in pre] == [8, 8, 8, 9, 15, 6]
    assert [d for *_, d in post] == [23, 23, 24, 31, 22, 5]
    n_cols = pre[0][0] + post[-1][0] + sum(m for _, m, _ in pre[:-1]) + sum(m for _, m, _ in post)
    assert n_cols == 252931
