"""proof-of-storage host logic (no GPU): dimension defaults, soundness, column choice, byte packing."""
import numpy as np
import pytest

from lcpc_proof_of_storage_b200 import pos


def test_default_aspect_ratio_matches_survey_table():
    # 4 GiB file -> 32768 -> 65536, 309 openings (server.rs:1139-1182)
    assert pos.get_aspect_ratio_default_from_file_len(1 << 32) == (32768, 65536, 309)
    assert pos.get_aspect_ratio_default_from_field_len(598) == (32, 64, 64)
    assert pos.get_soundness_from_matrix_dims(4, 8) == 8  # capped at the column count
    assert pos.get_soundness_from_matrix_dims(32768, 65536) == 309
    assert pos.dims_ok(4, 8) and not pos.dims_ok(5, 8) and not pos.dims_ok(4, 12) and not pos.dims_ok(0, 2)


@pytest.mark.parametrize("seed,amount,max_index", [(1337, 309, 65536), (1337, 8, 8), (7, 5, 3), (0, 64, 64), (42, 20, 1000)])
def test_column_choice_matches_oracle(oracle, seed, amount, max_index):
    got = pos.get_column_indicies_from_random_seed(seed, amount, max_index)
    assert got == oracle.pos_choose_columns(seed, amount, max_index)
    assert len(got) == min(amount, max_index) and len(set(got)) == len(got)  # without replacement
    assert all(0 <= c < max_index for c in got)


@pytest.mark.parametrize("n", [0, 1, 6, 7, 8, 13, 14, 598, 10001])
def test_byte_packing_round_trip(oracle, n):
    """fields.rs:286-383 round trips: bytes -> elements -> bytes (zero padded to a multiple of 7)."""
    rng = np.random.default_rng(n)
    data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    e = pos.convert_byte_vec_to_field_elements_vec(data)
    assert e.shape == ((n + 6) // 7, 1)
    assert np.array_equal(e, oracle.pack_bytes7(data))
    assert int(e.max(initial=0)) < (1 << 56)
    back = pos.field_vec_to_byte_vec(e)
    assert back[:n] == data and not any(back[n:])


def test_dimension_errors_follow_the_reference_messages():
    with pytest.raises(ValueError, match="empty file"):
        pos.convert_file_data_to_commit(b"", pos.Commit(), pos.Square())
    for dims, msg in [(pos.Specified(0, 8), "pre-encoded columns must be greater than 0"),
                      (pos.Specified(4, 1), "pencoded columns must be greater than 0"),
                      (pos.Specified(4, 12), "power of 2"),
                      (pos.Specified(8, 8), "greater than the number of columns")]:
        with pytest.raises(ValueError, match=msg):
            pos.convert_file_data_to_commit(b"x" * 100, pos.Commit(), dims)


# ----------------------------------------------------------------------------- encoded_file.py host pieces (no GPU)

def test_merkle_tree_mirror_matches_oracle_paths(oracle):
    """MerkleTree::{new layout, root, get_path, to_bytes, from_bytes} (lcpc_online/merkle_tree.rs:7-100) on a tree built by
    the oracle: get_path must equal open_column's path (lcpc-2d lib.rs:841-851)."""
    import numpy as np

    from lcpc_proof_of_storage_b200 import encoded_file as EF

    O = oracle
    c = O.commit(O.random_field_elements(0, 3, 700), O.LigeroEncoding(0, 16, 32))
    t = EF.MerkleTree(c.hashes)
    assert t.width == 32 and len(t) == 63 and t.root() == c.get_root()
    for col in (0, 1, 17, 31):
        assert t.get_path(col) == [bytes(p) for p in O.open_column(c, col).path]
    assert t.get_path(32) is None
    t2 = EF.MerkleTree.from_bytes(t.to_bytes())
    assert t2.root() == t.root() and t2[5] == t[5]
    import pytest

    with pytest.raises(ValueError):
        EF.MerkleTree(np.zeros((6, 32), dtype=np.uint8))  # not 2^k - 1 digests


def test_encoded_file_metadata_json_round_trip(tmp_path):
    """encoded_file_metadata.rs:5-27: the serde_json field names."""
    import json

    from lcpc_proof_of_storage_b200 import encoded_file as EF

    m = EF.EncodedFileMetadata(EF._new_ulid(), 4, 8, 22, 44, 598)
    path = tmp_path / "f.meta"
    with open(path, "wb") as f:
        m.write_to_file(f)
    raw = json.loads(path.read_text())
    assert list(raw) == ["ulid", "pre_encoded_size", "encoded_size", "rows_written", "row_capacity", "bytes_of_data"]
    assert len(raw["ulid"]) == 26 and set(raw["ulid"]) <= set("0123456789ABCDEFGHJKMNPQRSTVWXYZ")
    with open(path, "rb") as f:
        assert EF.EncodedFileMetadata.read_from_file(f) == m


def test_data_byte_capacity_per_data_field():
    """DataField::DATA_BYTE_CAPACITY: 7 for WriteableFt63 (writable_ft63.rs:30), 31 for Ft253_192 (ft253_192.rs:15);
    the other fields carry no file bytes."""
    import pytest

    from lcpc_proof_of_storage_b200 import pos
    from lcpc_proof_of_storage_b200.lcpc2d import FT63, FT127, FT253_192

    assert pos.data_byte_capacity(FT63) == 7 and pos.data_byte_capacity(FT253_192) == 31
    with pytest.raises(ValueError):
        pos.data_byte_capacity(FT127)
