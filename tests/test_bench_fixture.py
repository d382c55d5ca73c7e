"""tests/golden/bench_roots.json (the parity gate of bench.py's multi-GPU lines) against the CPU oracle.

Only the 1- and 2-GPU workloads are recomputed here (2^24 and 2^25 coefficients, a few seconds); the 4- and 8-GPU
roots come from the same generator, tests/golden/make_bench_roots.py.
"""
import json
import os

import pytest

import bench
from oracle import lcpc_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("n_gpus", [1, 2])
def test_bench_root_fixture_matches_oracle(n_gpus):
    with open(os.path.join(HERE, "golden", "bench_roots.json")) as f:
        roots = json.load(f)["roots_by_n_gpus"]
    assert set(roots) == {"1", "2", "4", "8"}
    O.build()
    enc = O.LigeroEncoding(bench.FID, bench.N_PER_ROW, bench.N_COLS)
    n_total = bench.ROWS_PER_GPU * n_gpus * bench.N_PER_ROW
    got = O.commit(bench.make_coeffs(2, n_total), enc).get_root().hex()
    assert got == roots[str(n_gpus)]
    assert bench.expected_root(n_gpus) == got


def test_integer_pipe_model_numbers():
    """bench.py's second roofline: floors from the essential instruction counts at 148 SMs / 1965 MHz,
    against the kernel times of profiles/r02c_bench_n1.json (the build with the 6 W + 2 I + 5 A product)."""
    with open(os.path.join(os.path.dirname(HERE), "profiles", "r02c_bench_n1.json")) as f:
        line = json.load(f)
    rep = bench.integer_pipe_report(line["roofline"]["kernels_ms_per_step"], 148, 1965)
    assert set(rep) == {"k_ntt_strided", "k_ntt_block", "k_hash_chunks"}
    for k, v in rep.items():
        assert 0.5 < v["frac"] < 1.0, (k, v)   # a floor: never above the measurement, and these kernels are close to it
    # ALU pipe binds: (2 * 30.4 + 2 * 73.3) cycles per warp-element * 2^25 / 32 elements / (4 * 148) / 1965 MHz
    assert abs(rep["k_ntt_block"]["floor_ms"] - 0.1869) < 1e-3
    assert bench.integer_pipe_report({"k_merkle_levels": 0.03}, 148, 1965) == {}
    assert bench.integer_pipe_report(line["roofline"]["kernels_ms_per_step"], 148, None) == {}


def test_synthetic_streams_numpy_and_torch_agree():
    """bench.py fills its inputs on the device with the torch form; the committed known answers were computed by the CPU
    oracle on the numpy form (tests/golden/make_bench_roots.py)."""
    import numpy as np
    import torch

    from lcpc_proof_of_storage_b200 import synth as S

    a = S.ft63_np(2, 5000)
    assert np.array_equal(a, bench.make_coeffs(2, 5000))
    assert np.array_equal(a, S.ft63_torch(2, 5000, "cpu").numpy().view(np.uint64).reshape(-1, 1))
    assert np.array_equal(S.ft63_np(2, 100, start=4900), a[4900:])
    assert (a < np.uint64(S.P63)).all()
    b = S.ft255_np(3, 333)
    assert np.array_equal(b, S.ft255_torch(3, 333, "cpu").numpy().view(np.uint64).reshape(-1, 4))
    assert (b[:, 3] < np.uint64(1 << 61)).all()
    assert np.array_equal(S.bytes_np(4, 1001), S.bytes_torch(4, 1001, "cpu").numpy())
    assert np.array_equal(S.bytes_np(4, 1001)[800:], S.bytes_torch(4, 201, "cpu", 800).numpy())


def test_bench_known_answers_cover_every_case():
    with open(os.path.join(HERE, "golden", "bench_roots.json")) as f:
        g = json.load(f)
    assert set(g["configs"]) == set(bench.CONFIG_CASES)
    assert {"ligero_root", "fold_sha256", "fold_encoded_sha256", "open_cols_sha256", "open_paths_sha256", "bytes_root",
            "brakedown_root"} <= set(g["parity"])


def test_parity_known_answers_match_oracle():
    """The small sharded cases of bench.py's parity_checks, recomputed by the oracle (a second or two)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_bench_roots", os.path.join(HERE, "golden", "make_bench_roots.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    O.build()
    with open(os.path.join(HERE, "golden", "bench_roots.json")) as f:
        assert m.parity_answers() == json.load(f)["parity"]
