#!/usr/bin/env python
"""Generates tests/golden/bench_roots.json: the Merkle root of bench.py's workload at 1, 2, 4 and 8 GPUs.

    python tests/golden/make_bench_roots.py

bench.py commits N x 512 rows x 32768 -> 65536 (Ft63, seed-2 coefficient stream) at N GPUs.  The CPU oracle leg of
bench.py runs at N = 1 only, so the multi-GPU lines are checked against these roots instead: computed here, once,
by the CPU oracle (oracle/) on the same seeded input -- "derived, not reference-attested" like the other oracle
fixtures (tests/golden/make_golden.py).  8 GPUs = 2^27 coefficients, a 2 GiB encoded matrix: about 11 s on 8 cores.
"""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (workload constants and the seeded generator only; no GPU code is touched)
from oracle import lcpc_oracle as O  # noqa: E402


def _sha(a) -> str:
    import hashlib

    import numpy as np

    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def config_answers() -> dict:
    """Known answers of bench.py's `configs` block (the other BASELINE.json configurations), by the CPU oracle on the
    numpy form of the same seeded inputs (lcpc_proof_of_storage_b200/synth.py)."""
    import numpy as np

    from lcpc_proof_of_storage_b200 import synth as S

    out = {}
    c = bench.CONFIG_CASES
    # (n_dt + 1) folds of the headline workload's coefficients
    k = c["fold_ft63_2^24"]
    coeffs = S.ft63_np(2, 1 << 24).reshape(512, 32768, 1)
    tens = S.ft63_np(k["tensor_seed"], k["n_tensors"] * 512).reshape(k["n_tensors"], 512, 1)
    out["fold_ft63_2^24"] = {"sha256": _sha(np.stack([O.collapse_columns(0, coeffs, t) for t in tens]))}
    print("fold 2^24", out["fold_ft63_2^24"], flush=True)
    # Ligero Ft63 2^28 commit + its folds
    k = c["ligero_ft63_2^28"]
    coeffs = S.ft63_np(k["seed"], 1 << 28)
    enc = O.LigeroEncoding(0, k["n_per_row"], k["n_cols"])
    comm = O.commit(coeffs, enc)
    out["ligero_ft63_2^28"] = {"root": comm.get_root().hex()}
    print("ligero 2^28", out["ligero_ft63_2^28"], flush=True)
    k = c["fold_ft63_2^28"]
    n_rows = (1 << 28) // 131072
    tens = S.ft63_np(k["tensor_seed"], k["n_tensors"] * n_rows).reshape(k["n_tensors"], n_rows, 1)
    out["fold_ft63_2^28"] = {"sha256": _sha(np.stack([O.collapse_columns(0, comm.coeffs, t) for t in tens]))}
    print("fold 2^28", out["fold_ft63_2^28"], flush=True)
    del comm, coeffs
    # Brakedown code 3 over Ft255, 2^24 coefficients
    k = c["brakedown_ft255_2^24"]
    enc = O.SdigEncoding.new(3, 1 << 24, k["code_seed"])
    assert (enc.n_per_row, enc.n_cols) == (k["n_per_row"], k["n_cols"])
    comm = O.commit(S.ft255_np(k["seed"], 1 << 24), enc)
    out["brakedown_ft255_2^24"] = {"root": comm.get_root().hex()}
    print("brakedown ft255 2^24", out["brakedown_ft255_2^24"], flush=True)
    out.update(config_answers_small())
    return out


def config_answers_small() -> dict:
    """The two configs[4] points added late in round 2 (`--only-small` merges just these into the committed file)."""
    from lcpc_proof_of_storage_b200 import synth as S

    out = {}
    c = bench.CONFIG_CASES
    k = c["brakedown_ft63_2^24"]
    enc = O.SdigEncoding.new(0, 1 << 24, k["code_seed"])
    assert (enc.n_per_row, enc.n_cols) == (k["n_per_row"], k["n_cols"])
    out["brakedown_ft63_2^24"] = {"root": O.commit(S.ft63_np(k["seed"], 1 << 24), enc).get_root().hex()}
    print("brakedown ft63 2^24", out["brakedown_ft63_2^24"], flush=True)
    k = c["ligero_ft63_2^20"]
    enc = O.LigeroEncoding(0, k["n_per_row"], k["n_cols"])
    out["ligero_ft63_2^20"] = {"root": O.commit(S.ft63_np(k["seed"], 1 << 20), enc).get_root().hex()}
    print("ligero ft63 2^20", out["ligero_ft63_2^20"], flush=True)
    return out


def parity_answers() -> dict:
    """Known answers of the small sharded cases bench.py checks at N > 1 before its timed region."""
    import numpy as np

    from lcpc_proof_of_storage_b200 import synth as S

    c = bench.PARITY_CASES
    out = {}
    k = c["ligero"]
    enc = O.LigeroEncoding(0, k["n_per_row"], k["n_cols"])
    comm = O.commit(S.ft63_np(k["seed"], k["n_rows"] * k["n_per_row"]), enc)
    out["ligero_root"] = comm.get_root().hex()
    tens = S.ft63_np(k["tensor_seed"], 2 * k["n_rows"]).reshape(2, k["n_rows"], 1)
    out["fold_sha256"] = _sha(np.stack([O.collapse_columns(0, comm.coeffs, t) for t in tens]))
    out["fold_encoded_sha256"] = _sha(np.stack([O.collapse_columns(0, comm.comm, t) for t in tens]))
    opened = [O.open_column(comm, j) for j in k["open"]]
    out["open_cols_sha256"] = _sha(np.stack([o.col for o in opened]))
    out["open_paths_sha256"] = _sha(np.stack([o.path for o in opened]))
    k = c["bytes"]
    enc = O.LigeroEncoding(0, k["n_per_row"], k["n_cols"])
    out["bytes_root"] = O.commit(O.pack_bytes7(S.bytes_np(k["seed"], k["n_bytes"]).tobytes()), enc).get_root().hex()
    k = c["brakedown"]
    enc = O.SdigEncoding(0, k["n_per_row"], k["code_seed"])
    assert enc.n_cols == k["n_cols"], enc.n_cols
    out["brakedown_root"] = O.commit(S.ft63_np(k["seed"], k["n_rows"] * k["n_per_row"] - 5), enc).get_root().hex()
    print("parity", out, flush=True)
    return out


def main() -> None:
    O.build()
    O.set_threads(os.cpu_count() or 1)
    path = os.path.join(HERE, "bench_roots.json")
    if "--only-small" in sys.argv:
        with open(path) as f:
            out = json.load(f)
        out["configs"].update(config_answers_small())
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
            f.write("\n")
        return
    enc = O.LigeroEncoding(bench.FID, bench.N_PER_ROW, bench.N_COLS)
    roots = {}
    for world in (1, 2, 4, 8):
        n_total = bench.ROWS_PER_GPU * world * bench.N_PER_ROW
        roots[str(world)] = O.commit(bench.make_coeffs(2, n_total), enc).get_root().hex()
        print(world, roots[str(world)], flush=True)
    out = {
        "source": "derived, not reference-attested: CPU oracle (oracle/) on bench.make_coeffs(2, N * 2^24)",
        "workload": "Ligero Ft63 rho=1/2 BLAKE3, N x 512 rows x 32768 -> 65536",
        "roots_by_n_gpus": roots,
        "configs": config_answers(),
        "parity": parity_answers(),
    }
    with open(os.path.join(HERE, "bench_roots.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    main()
