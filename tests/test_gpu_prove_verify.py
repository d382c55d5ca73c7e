"""prove / verify through the C ABI against the oracle: identical transcripts must give
identical proofs, verification must return the true evaluation, and tampering must be
rejected with the reference's VerifierError variant (lcpc-2d/src/tests.rs:236-413,
lcpc-ligero-pc/src/tests.rs:216-401, lcpc-brakedown-pc/src/tests.rs:192-375,
proof-of-storage/src/networking/tests.rs:696-780)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import lcpc_proof_of_storage_b200 as pkg

    return pkg


def _tensors(O, fid, n_rows, n_per_row, x):
    p = O.MODULUS[fid]
    inner = O.to_mont(fid, [pow(x, j, p) for j in range(n_per_row)])
    xr = pow(x, n_per_row, p)
    outer = O.to_mont(fid, [pow(xr, i, p) for i in range(n_rows)])
    return outer, inner


def _eval(O, fid, coeffs, x):
    p = O.MODULUS[fid]
    acc = 0
    for c in reversed(O.from_mont(fid, coeffs)):
        acc = (acc * x + c) % p
    return acc


def _start(cls, root):
    tr = cls(b"test transcript")
    tr.append_message(b"polycommit", root)
    tr.append_message(b"ncols", (309).to_bytes(8, "big"))
    return tr


def _same_proof(a, b):
    assert a.n_cols == b.n_cols
    assert np.array_equal(a.p_eval, b.p_eval)
    assert len(a.p_random_vec) == len(b.p_random_vec)
    for u, v in zip(a.p_random_vec, b.p_random_vec):
        assert np.array_equal(u, v)
    assert len(a.columns) == len(b.columns)
    for u, v in zip(a.columns, b.columns):
        assert np.array_equal(u.col, v.col) and np.array_equal(u.path, v.path)


@pytest.mark.parametrize("fid,length", [(0, 1 << 12), (0, 5000), (1, 1 << 11), (3, 1 << 10), (4, 1 << 10)])
def test_ligero_prove_verify_matches_oracle(P, oracle, fid, length):
    O = oracle
    coeffs = O.random_field_elements(fid, 70 + fid, length)
    enc = P.LigeroEncoding.new(fid, length)
    oenc = O.LigeroEncoding.new(fid, length)
    comm = P.LcCommit.commit(coeffs, enc)
    ocomm = O.commit(coeffs, oenc)
    root = comm.get_root()
    assert root == ocomm.get_root()
    x = 0xABCDEF123 + fid
    outer, inner = _tensors(O, fid, comm.n_rows, comm.n_per_row, x)
    pf = comm.prove(outer, enc, _start(P.Transcript, root))
    opf = O.prove(ocomm, outer, oenc, _start(O.Transcript, root))
    _same_proof(pf, opf)
    assert len(pf.columns) == 309
    # verifier side: fresh encoding from the proof's dimensions (tests.rs:287)
    enc2 = P.LigeroEncoding(fid, pf.get_n_per_row(), pf.get_n_cols())
    tr_v = _start(P.Transcript, root)
    res = pf.verify(root, outer, inner, enc2, tr_v)
    assert O.from_mont(fid, res)[0] == _eval(O, fid, coeffs, x)
    ores = O.verify(root, outer, inner, opf, oenc, _start(O.Transcript, root))
    assert np.array_equal(res, ores)
    # the oracle accepts the GPU proof and vice versa
    assert np.array_equal(O.verify(root, outer, inner, O.LcEvalProof(pf.n_cols, pf.p_eval, pf.p_random_vec,
                          [O.LcColumn(c.col, c.path) for c in pf.columns]), oenc, _start(O.Transcript, root)), res)
    # prover and verifier transcripts end in the same state
    tr_p = _start(P.Transcript, root)
    comm.prove(outer, enc, tr_p)
    assert tr_p.challenge_bytes(b"after", 32) == tr_v.challenge_bytes(b"after", 32)


def test_two_proofs_on_one_transcript(P, oracle):
    """lcpc-2d/src/tests.rs:318-413."""
    O = oracle
    fid, length = 0, 3000
    coeffs = O.random_field_elements(fid, 81, length)
    enc = P.LigeroEncoding.new(fid, length)
    comm = P.LcCommit.commit(coeffs, enc)
    root = comm.get_root()
    tr1, tr2 = _start(P.Transcript, root), _start(P.Transcript, root)
    xs = (17, 0x7777777)
    proofs = []
    for x in xs:
        outer, _ = _tensors(O, fid, comm.n_rows, comm.n_per_row, x)
        proofs.append(comm.prove(outer, enc, tr1))
    for x, pf in zip(xs, proofs):
        outer, inner = _tensors(O, fid, comm.n_rows, comm.n_per_row, x)
        res = pf.verify(root, outer, inner, enc, tr2)
        assert O.from_mont(fid, res)[0] == _eval(O, fid, coeffs, x)


def test_verify_rejects_tampering_with_reference_variants(P, oracle):
    O = oracle
    fid, length = 0, 4000
    coeffs = O.random_field_elements(fid, 91, length)
    enc = P.LigeroEncoding.new(fid, length)
    comm = P.LcCommit.commit(coeffs, enc)
    root = comm.get_root()
    outer, inner = _tensors(O, fid, comm.n_rows, comm.n_per_row, 5)

    def run(mutate, root_=root, outer_=outer, inner_=inner):
        pf = comm.prove(outer, enc, _start(P.Transcript, root))
        mutate(pf)
        return pf.verify(root_, outer_, inner_, enc, _start(P.Transcript, root))

    run(lambda pf: None)
    cases = [
        (lambda pf: pf.columns[7].col.__setitem__((0, 0), pf.columns[7].col[0, 0] ^ np.uint64(1)), "ColumnDegree"),
        (lambda pf: pf.columns[300].path.__setitem__((2, 5), pf.columns[300].path[2, 5] ^ np.uint8(1)), "ColumnPath"),
        (lambda pf: pf.columns.pop(), "NumColOpens"),
        # p_eval feeds the transcript, so the verifier samples other columns than the prover opened:
        # the degree check of the first column fails first (match order at lib.rs:968-973)
        (lambda pf: pf.p_eval.__setitem__((3, 0), pf.p_eval[3, 0] ^ np.uint64(1)), "ColumnDegree"),
    ]
    # non-canonical alias x + p of a column element: same leaf bytes and transcript as the honest value, but not a value
    # the reference's deserialiser (from_repr) would ever hand to its verifier -- refused before any kernel runs
    pmod = np.uint64(O.MODULUS[fid])

    def alias(pf):
        with np.errstate(over="ignore"):
            pf.columns[11].col[3, 0] = pf.columns[11].col[3, 0] + pmod
    cases.append((alias, "ColumnEval"))
    # a p_random vector of the wrong length (short or long) never reaches the flat C-ABI buffer
    cases.append((lambda pf: pf.p_random_vec.__setitem__(1, pf.p_random_vec[1][:-1]), "ColumnDegree"))
    cases.append((lambda pf: pf.p_random_vec.__setitem__(0, np.concatenate([pf.p_random_vec[0], pf.p_random_vec[0][:1]])),
                  "ColumnDegree"))
    for mutate, variant in cases:
        with pytest.raises(P.VerifierError) as ei:
            run(mutate)
        assert ei.value.variant == variant
    with pytest.raises(P.VerifierError) as ei:
        run(lambda pf: None, outer_=outer[:-1])
    assert ei.value.variant == "OuterTensor"
    bad_inner = inner.copy()
    with np.errstate(over="ignore"):
        bad_inner[0, 0] = bad_inner[0, 0] + pmod
    from lcpc_proof_of_storage_b200._lib import LcpcError
    with pytest.raises(LcpcError) as ei2:
        run(lambda pf: None, inner_=bad_inner)
    assert ei2.value.variant == "InvalidArg"
    wrong_outer = outer.copy()
    wrong_outer[1, 0] ^= np.uint64(1)  # same transcript, wrong evaluation tensor: only the eval check fails
    with pytest.raises(P.VerifierError) as ei:
        run(lambda pf: None, outer_=wrong_outer)
    assert ei.value.variant == "ColumnEval"
    with pytest.raises(P.VerifierError) as ei:
        run(lambda pf: None, inner_=inner[:-1])
    assert ei.value.variant == "InnerTensor"
    with pytest.raises(P.VerifierError) as ei:
        run(lambda pf: None, root_=bytes(32))
    assert ei.value.variant == "ColumnPath"
    with pytest.raises(P.VerifierError) as ei:
        pf = comm.prove(outer, enc, _start(P.Transcript, root))
        pf.verify(root, outer, inner, P.LigeroEncoding(fid, enc.n_per_row // 2, enc.n_cols), _start(P.Transcript, root))
    assert ei.value.variant == "EncodingDims"
    with pytest.raises(P.ProverError) as ei:
        comm.prove(outer[:-1], enc, _start(P.Transcript, root))
    assert ei.value.variant == "OuterTensor"


@pytest.mark.parametrize("fid,seed", [(0, 0), (3, 1), (4, 1)])
def test_brakedown_prove_verify_matches_oracle(P, oracle, fid, seed):
    """SdigEncoding::new(len, seed) on both sides (host-side matgen), then commit/prove/verify."""
    O = oracle
    length = 6000
    coeffs = O.random_field_elements(fid, 95, length)
    enc = P.SdigEncoding.new(fid, length, seed)
    oenc = O.SdigEncoding.new(fid, length, seed)
    assert (enc.n_per_row, enc.n_cols) == (oenc.n_per_row, oenc.n_cols)
    assert enc.get_n_col_opens() == oenc.get_n_col_opens() == 6593
    assert enc.get_n_degree_tests() == oenc.get_n_degree_tests()
    comm = P.LcCommit.commit(coeffs, enc)
    ocomm = O.commit(coeffs, oenc)
    root = comm.get_root()
    assert root == ocomm.get_root()
    x = 31337
    outer, inner = _tensors(O, fid, comm.n_rows, comm.n_per_row, x)
    pf = comm.prove(outer, enc, _start(P.Transcript, root))
    opf = O.prove(ocomm, outer, oenc, _start(O.Transcript, root))
    _same_proof(pf, opf)
    res = pf.verify(root, outer, inner, enc, _start(P.Transcript, root))
    assert O.from_mont(fid, res)[0] == _eval(O, fid, coeffs, x)


def test_ligero_2_16_commit_prove_verify(P, oracle):
    """BASELINE configs[0]: commit + prove + verify, 2^16 coefficients over the 63-bit field."""
    O = oracle
    fid, length = 0, 1 << 16
    coeffs = O.random_field_elements(fid, 1, length)
    enc = P.LigeroEncoding.new(fid, length)
    comm = P.LcCommit.commit(coeffs, enc)
    root = comm.get_root()
    x = 0x1234567890ABCDEF % O.MODULUS[fid]
    outer, inner = _tensors(O, fid, comm.n_rows, comm.n_per_row, x)
    pf = comm.prove(outer, enc, _start(P.Transcript, root))
    ocomm = O.commit(coeffs, O.LigeroEncoding.new(fid, length))
    opf = O.prove(ocomm, outer, O.LigeroEncoding.new(fid, length), _start(O.Transcript, root))
    _same_proof(pf, opf)
    res = pf.verify(root, outer, inner, enc, _start(P.Transcript, root))
    assert O.from_mont(fid, res)[0] == _eval(O, fid, coeffs, x)
