"""Streaming commit (proof-of-storage EncodedFileWriter / ColumnDigestAccumulator) and row edits
(FileHandler::edit_bytes) through the C ABI, against the in-memory oracle commit: the reference's own test for this
path is "streamed root == in-memory root" (row_generator_iter.rs:286-364, lcpc_online/tests.rs) and
"edit then re-derive == commit of the edited file" (file_handler tests)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import lcpc_proof_of_storage_b200 as pkg

    return pkg


def _stream_elems(P, enc, elems, max_rows, block_rows, pushes):
    from lcpc_proof_of_storage_b200 import _lib

    lib = _lib.load()
    s = C.c_void_p()
    _lib.check(lib.lcpc_stream_begin(enc.plan, max_rows, block_rows, None, 0, C.byref(s)))
    try:
        off = 0
        for n in pushes:
            part = np.ascontiguousarray(elems[off:off + n])
            _lib.check(lib.lcpc_stream_push_elems_host(s, part.ctypes.data, part.shape[0]))
            off += n
        assert off == elems.shape[0]
        hashes = np.empty((2 * P.next_pow2(enc.n_cols) - 1, 32), dtype=np.uint8)
        rows = C.c_size_t()
        _lib.check(lib.lcpc_stream_finish(s, hashes.ctypes.data, C.byref(rows)))
        return hashes, rows.value
    finally:
        lib.lcpc_stream_free(s)


@pytest.mark.parametrize("fid,n_per_row,n_cols,n_rows,ragged,block_rows", [
    (0, 64, 128, 1, 5, 1),          # a single short row: one chunk, ROOT on the chunk itself
    (0, 64, 128, 124, 0, 50),       # leaf = exactly one full chunk
    (0, 64, 128, 125, 0, 50),       # one chunk + 8 bytes
    (0, 64, 128, 700, 13, 97),      # 6 chunks, blocks not aligned with chunks
    (0, 64, 128, 700, 0, 1000),     # everything in one block
    (0, 2048, 4096, 300, 777, 128),
    (1, 64, 128, 130, 3, 17),       # 16-byte elements: 64 rows per chunk
    (2, 16, 32, 200, 0, 23),        # 24-byte elements straddle chunk boundaries
    (3, 64, 128, 70, 9, 11),        # 32-byte elements
    (4, 64, 128, 70, 9, 11),        # Ft253_192: big-endian repr
])
def test_stream_commit_equals_in_memory_commit(P, oracle, fid, n_per_row, n_cols, n_rows, ragged, block_rows):
    O = oracle
    n = n_rows * n_per_row - ragged
    elems = O.random_field_elements(fid, 900 + fid, n)
    exp = O.commit(elems, O.LigeroEncoding(fid, n_per_row, n_cols))
    enc = P.LigeroEncoding(fid, n_per_row, n_cols)
    # pushes of whole rows in uneven batches, the ragged tail last
    pushes, left, k = [], n, 1
    while left > 0:
        take = min(left, k * n_per_row)
        pushes.append(take)
        left -= take
        k = k * 3 + 1
    hashes, rows = _stream_elems(P, enc, elems, n_rows, block_rows, pushes)
    assert rows == exp.n_rows
    assert np.array_equal(hashes, exp.hashes)


def test_stream_rejects_push_after_ragged_row_and_overflow(P, oracle):
    from lcpc_proof_of_storage_b200 import _lib

    lib = _lib.load()
    enc = P.LigeroEncoding(0, 8, 16)
    elems = oracle.random_field_elements(0, 1, 100)
    s = C.c_void_p()
    _lib.check(lib.lcpc_stream_begin(enc.plan, 4, 0, None, 0, C.byref(s)))
    try:
        assert lib.lcpc_stream_push_elems_host(s, elems.ctypes.data, 12) == 0      # 1.5 rows
        assert lib.lcpc_stream_push_elems_host(s, elems.ctypes.data, 8) == -6      # LCPC_ERR_DIMS
    finally:
        lib.lcpc_stream_free(s)
    _lib.check(lib.lcpc_stream_begin(enc.plan, 2, 0, None, 0, C.byref(s)))
    try:
        assert lib.lcpc_stream_push_elems_host(s, elems.ctypes.data, 24) == -1     # 3 rows > max_rows: TooBig
        assert lib.lcpc_stream_finish(s, None, None) != 0                          # nothing committed
    finally:
        lib.lcpc_stream_free(s)


@pytest.mark.parametrize("n_bytes,pre,enc_cols,read_size", [(598, 4, 8, 100), (100003, 64, 128, 4096), (3_000_000, 512, 1024, 1 << 20)])
def test_convert_unencoded_file_layout_and_tree(P, oracle, tmp_path, n_bytes, pre, enc_cols, read_size):
    """EncodedFileWriter::convert_unencoded_file: .porenc column-major canonical bytes, .portree digests, .meta JSON."""
    from lcpc_proof_of_storage_b200 import encoded_file as EF

    O = oracle
    rng = np.random.default_rng(n_bytes)
    data = rng.integers(0, 256, n_bytes, dtype=np.uint8).tobytes()
    src, dst, tree_f, meta_f = [str(tmp_path / x) for x in ("raw.bin", "file.porenc", "file.portree", "file.meta")]
    with open(src, "wb") as f:
        f.write(data)
    meta, tree = EF.EncodedFileWriter.convert_unencoded_file(src, dst, tree_f, meta_f, pre, enc_cols, read_size=read_size)
    exp = O.commit(O.pack_bytes7(data), O.LigeroEncoding(0, pre, enc_cols))
    assert meta.rows_written == exp.n_rows and meta.row_capacity == 2 * exp.n_rows and meta.bytes_of_data == n_bytes
    assert (meta.pre_encoded_size, meta.encoded_size) == (pre, enc_cols) and len(meta.ulid) == 26
    assert tree.root() == exp.get_root()
    assert np.array_equal(tree.digests, exp.hashes)
    assert os.path.getsize(dst) == meta.row_capacity * enc_cols * 8
    canon = O.fe_to_canon(0, exp.comm.reshape(-1, 1)).reshape(exp.n_rows, enc_cols)
    for c in (0, 1, enc_cols // 2 + 1, enc_cols - 1):
        assert np.array_equal(EF.read_encoded_column(dst, meta, c), canon[:, c])
    whole = np.fromfile(dst, dtype="<u8").reshape(enc_cols, meta.row_capacity)
    assert np.array_equal(whole[:, :exp.n_rows], canon.T) and not whole[:, exp.n_rows:].any()
    # the side files round-trip
    with open(meta_f, "rb") as f:
        assert EF.EncodedFileMetadata.read_from_file(f) == meta
    with open(tree_f, "rb") as f:
        t2 = EF.MerkleTree.from_bytes(f.read())
    assert t2.root() == tree.root() and len(t2) == 2 * enc_cols - 1
    # MerkleTree::get_path agrees with open_column's path
    col = O.open_column(exp, 3)
    assert [bytes(p) for p in col.path] == tree.get_path(3)


@pytest.mark.parametrize("fid,n_per_row,n_cols,n_rows,row0,k", [
    (0, 64, 128, 700, 0, 1), (0, 64, 128, 700, 123, 3), (0, 64, 128, 700, 699, 1), (0, 64, 128, 700, 100, 400),
    (0, 64, 128, 50, 7, 2),      # single-chunk leaves
    (2, 16, 32, 200, 41, 5), (3, 64, 128, 70, 30, 9), (4, 64, 128, 70, 30, 9),
])
def test_update_rows_equals_recommit(P, oracle, fid, n_per_row, n_cols, n_rows, row0, k):
    O = oracle
    n = n_rows * n_per_row
    elems = O.random_field_elements(fid, 31 + fid, n)
    enc = P.LigeroEncoding(fid, n_per_row, n_cols)
    c = P.LcCommit.commit(elems, enc)
    new_rows = O.random_field_elements(fid, 77, k * n_per_row)
    hashes = c.update_rows(row0, new_rows)
    edited = elems.copy()
    edited[row0 * n_per_row:(row0 + k) * n_per_row] = new_rows
    exp = O.commit(edited, O.LigeroEncoding(fid, n_per_row, n_cols))
    assert np.array_equal(hashes, exp.hashes)
    assert c.get_root() == exp.get_root()
    assert np.array_equal(c.comm, exp.comm) and np.array_equal(c.coeffs, exp.coeffs)
    # the handle itself is up to date: openings and folds come from the edited matrix
    col = c.open_columns([5])[0]
    e = O.open_column(exp, 5)
    assert np.array_equal(col.col, e.col) and np.array_equal(col.path, e.path)
    with pytest.raises(Exception):
        c.update_rows(n_rows, new_rows[:n_per_row])


def test_edit_bytes_matches_commit_of_edited_file(P, oracle):
    from lcpc_proof_of_storage_b200 import encoded_file as EF

    O = oracle
    rng = np.random.default_rng(5)
    data = bytearray(rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes())
    enc = P.LigeroEncoding(0, 64, 128)
    c = P.LcCommit.commit_bytes(bytes(data), enc)
    patch = bytes(rng.integers(0, 256, 1500, dtype=np.uint8))
    original, tree = EF.edit_bytes(c, len(data), 100_001, patch)
    assert original == bytes(data[100_001:100_001 + 1500])
    data[100_001:100_001 + 1500] = patch
    exp = O.commit(O.pack_bytes7(bytes(data)), O.LigeroEncoding(0, 64, 128))
    assert tree.root() == exp.get_root() == c.get_root()
    with pytest.raises(ValueError):
        EF.edit_bytes(c, len(data), len(data) - 3, b"12345")


@pytest.mark.parametrize("fid,n_per_row,n_cols,n_rows,row0,k", [
    (0, 64, 128, 50, 50, 3),       # single chunk stays a single chunk
    (0, 64, 128, 120, 119, 10),    # grows from one chunk to two: the ROOT-flagged chunk is re-hashed
    (0, 64, 128, 700, 700, 1), (0, 64, 128, 700, 699, 300),
    (2, 16, 32, 200, 199, 40), (3, 64, 128, 70, 70, 33),
])
def test_append_rows_equals_recommit(P, oracle, fid, n_per_row, n_cols, n_rows, row0, k):
    O = oracle
    elems = O.random_field_elements(fid, 11 + fid, n_rows * n_per_row)
    c = P.LcCommit.commit(elems, P.LigeroEncoding(fid, n_per_row, n_cols))
    new_rows = O.random_field_elements(fid, 78, k * n_per_row)
    hashes = c.append_rows(row0, new_rows)
    grown = np.concatenate([elems[:row0 * n_per_row], new_rows])
    exp = O.commit(grown, O.LigeroEncoding(fid, n_per_row, n_cols))
    assert c.n_rows == exp.n_rows
    assert np.array_equal(hashes, exp.hashes) and c.get_root() == exp.get_root()
    assert np.array_equal(c.comm, exp.comm) and np.array_equal(c.coeffs, exp.coeffs)
    with pytest.raises(Exception):
        c.append_rows(0, new_rows[:n_per_row])  # an append must reach the end


def test_append_bytes_matches_commit_of_longer_file(P, oracle):
    from lcpc_proof_of_storage_b200 import encoded_file as EF

    O = oracle
    rng = np.random.default_rng(6)
    data = rng.integers(0, 256, 100_003, dtype=np.uint8).tobytes()
    more = rng.integers(0, 256, 54_321, dtype=np.uint8).tobytes()
    c = P.LcCommit.commit_bytes(data, P.LigeroEncoding(0, 64, 128))
    tree = EF.append_bytes(c, len(data), more)
    exp = O.commit(O.pack_bytes7(data + more), O.LigeroEncoding(0, 64, 128))
    assert tree.root() == exp.get_root() == c.get_root()
    assert np.array_equal(c.coeffs, exp.coeffs)
