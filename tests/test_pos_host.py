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
