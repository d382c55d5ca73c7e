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
    against the kernel times of profiles/r01g_bench.json."""
    with open(os.path.join(os.path.dirname(HERE), "profiles", "r01g_bench.json")) as f:
        line = json.load(f)
    rep = bench.integer_pipe_report(line["roofline"]["kernels_ms_per_step"], 148, 1965)
    assert set(rep) == {"k_ntt_strided", "k_ntt_block", "k_hash_chunks"}
    for k, v in rep.items():
        assert 0.5 < v["frac"] < 1.0, (k, v)   # a floor: never above the measurement, and these kernels are close to it
    assert abs(rep["k_ntt_block"]["floor_ms"] - 0.2144) < 1e-3
    assert bench.integer_pipe_report({"k_merkle_levels": 0.03}, 148, 1965) == {}
    assert bench.integer_pipe_report(line["roofline"]["kernels_ms_per_step"], 148, None) == {}
