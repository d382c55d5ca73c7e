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


def main() -> None:
    O.build()
    O.set_threads(os.cpu_count() or 1)
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
    }
    with open(os.path.join(HERE, "bench_roots.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    main()
