"""proof-of-storage commit path on the GPU (mirrors proof-of-storage/src/tests.rs and lcpc_online.rs
tests): file bytes -> commit, Leaves / Columns request kinds, the client's retrievability check,
rejection after tampering, and the tall-vs-wide polynomial evaluation identity."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import lcpc_proof_of_storage_b200 as pkg

    return pkg


@pytest.fixture(scope="module")
def pos():
    from lcpc_proof_of_storage_b200 import pos as m

    return m


def _file(n, seed=4):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8).tobytes()


@pytest.mark.parametrize("n_bytes,dims", [(598, "4->8"), (598, "square"), (70001, "64->128"), (3_000_000, "default")])
def test_upload_then_retrievability_proof(P, pos, oracle, n_bytes, dims):
    """client.rs:51-281 flow: commit, choose columns from seed 1337, local leaves vs server columns + paths."""
    O = oracle
    data = _file(n_bytes)
    if dims == "square":
        d = pos.Square()
    elif dims == "default":
        pre, enc, _ = pos.get_aspect_ratio_default_from_file_len(n_bytes)
        d = pos.Specified(pre, enc)
    else:
        a, b = dims.split("->")
        d = pos.Specified(int(a), int(b))
    comm = pos.convert_file_data_to_commit(data, pos.Commit(), d)
    # same commitment as the oracle on the packed elements
    elems = O.pack_bytes7(data)
    exp = O.commit(elems, O.LigeroEncoding(0, comm.n_per_row, comm.n_cols))
    assert comm.get_root() == exp.get_root()
    assert np.array_equal(comm.coeffs, exp.coeffs)
    soundness = pos.get_soundness_from_matrix_dims(comm.n_per_row, comm.n_cols)
    cols = pos.get_column_indicies_from_random_seed(1337, soundness, comm.n_cols)
    # client: leaves only; server: columns with paths (recommit from the raw bytes, server.rs:670-682)
    leaves = pos.convert_file_data_to_commit(data, pos.Leaves(cols), d)
    assert np.array_equal(leaves, exp.hashes[cols])
    served = pos.convert_file_data_to_commit(data, pos.ColumnsWithPath(cols), d)
    pos.client_verify_commitment(comm.get_root(), leaves, cols, served, soundness)
    bare = pos.convert_file_data_to_commit(data, pos.ColumnsWithoutPath(cols), d)
    for a, b, c in zip(bare, served, cols):
        assert np.array_equal(a, b.col) and np.array_equal(a, exp.comm[:, c])
        assert pos.hash_column_to_digest(b) == exp.hashes[c].tobytes()
    # T8 (networking/tests.rs:696-780): two bytes of the server's copy change -> the proof must fail
    bad = bytearray(data)
    bad[n_bytes // 2] ^= 0x01
    bad[n_bytes // 3] ^= 0x80
    served_bad = pos.convert_file_data_to_commit(bytes(bad), pos.ColumnsWithPath(cols), d)
    if len(cols) == comm.n_cols or n_bytes > 1000:  # small shapes open every column: always detected
        with pytest.raises(P.VerifierError):
            pos.client_verify_commitment(comm.get_root(), leaves, cols, served_bad, soundness)
    with pytest.raises(P.VerifierError) as ei:
        pos.client_verify_commitment(comm.get_root(), leaves, cols, served, soundness - 1)
    assert ei.value.variant == "NumColOpens"
    with pytest.raises(P.VerifierError) as ei:
        pos.client_online_verify_column_paths(bytes(32), cols, served)
    assert ei.value.variant == "ColumnEval"


def test_encode_then_decode_row(P, pos, oracle):
    """The reference's own test (lcpc_online.rs:588-601): commit a short vector as a single-row matrix, decode the encoded
    row, get the coefficients back."""
    O = oracle
    row = O.random_field_elements(0, 17, 4)
    enc = P.LigeroEncoding(0, 1 << 4, 1 << 8)
    commit = P.LcCommit.commit(row, enc)
    back = pos.decode_row(commit.comm.reshape(commit.n_rows, 1 << 8, 1), enc)
    assert np.array_equal(back[0, :16], commit.coeffs.reshape(-1, 1)[:16])
    assert not back[0, 16:].any()


def test_polynomial_evaluation_tall_vs_wide(P, pos, oracle):
    """lcpc_online.rs:629-674 / networking/tests.rs:374-466: the fold of the encoded matrix decodes to
    the fold of the file's elements, and evaluating through a tall or a wide layout agrees."""
    O = oracle
    p = O.MODULUS[0]
    data = _file(7 * 4096)
    elems = O.pack_bytes7(data)
    x = 123456789
    results = []
    for pre, enc in [(64, 128), (256, 512)]:
        comm = pos.convert_file_data_to_commit(data, pos.Commit(), pos.Specified(pre, enc))
        xr = pow(x, pre, p)
        left = O.to_mont(0, [pow(xr, i, p) for i in range(comm.n_rows)])
        folded = pos.verifiable_polynomial_evaluation(comm, left)
        assert np.array_equal(folded, O.collapse_columns(0, comm.comm, left))
        poly = O.from_mont(0, O.ifft_oi(0, folded.reshape(1, enc, 1))[0])
        # decode_row on the GPU (lcpc_online.rs:568-573) gives the same coefficients
        assert O.from_mont(0, pos.decode_row(folded.reshape(1, enc, 1), comm.enc)[0]) == poly
        assert all(v == 0 for v in poly[pre:])
        results.append(sum(c * pow(x, j, p) for j, c in enumerate(poly[:pre])) % p)
    direct = sum(c * pow(x, i, p) for i, c in enumerate(O.from_mont(0, elems))) % p
    assert results[0] == results[1] == direct
