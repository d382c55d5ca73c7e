#!/usr/bin/env python
"""Generates tests/golden/pos_4gib.json: BASELINE.json configs[3], the proof-of-storage commit of a 4 GiB synthetic file
(bytes = lcpc_proof_of_storage_b200.synth.bytes_np(4, 2^32); WriteableFt63 7-byte packing; default aspect of
networking/server.rs:1139-1182: 18725 rows x 32768 -> 65536) and its retrievability proof: the 309 columns of
get_column_indicies_from_random_seed(1337, ..) (networking/client.rs:443-456).  Computed once by the CPU oracle
(about 15 GiB of RAM, a minute or two): "derived, not reference-attested" like the other oracle fixtures.

    python tests/golden/make_pos_golden.py
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from lcpc_proof_of_storage_b200 import synth as S  # noqa: E402
from oracle import lcpc_oracle as O  # noqa: E402

N_BYTES = 1 << 32
PRE, ENC, OPENS = 32768, 65536, 309


def main() -> None:
    O.build()
    O.set_threads(os.cpu_count() or 1)
    n_elems = (N_BYTES + 6) // 7
    coeffs = np.zeros((n_elems, 1), dtype=np.uint64)
    step = 7 * (1 << 24)                      # bytes per slab: a whole number of elements and of stream words
    for b0 in range(0, N_BYTES, step):
        nb = min(step, N_BYTES - b0)
        data = S.splitmix_np(4, (nb + 7) // 8, b0 // 8).view(np.uint8)[:nb]
        e0 = b0 // 7
        coeffs[e0:e0 + (nb + 6) // 7] = O.pack_bytes7(data.tobytes())
    comm = O.commit(coeffs, O.LigeroEncoding(0, PRE, ENC))
    cols = O.pos_choose_columns(1337, OPENS, ENC)
    leaves = np.stack([comm.hashes[c] for c in cols])
    vals = np.stack([O.open_column(comm, c).col for c in cols])
    paths = np.stack([O.open_column(comm, c).path for c in cols])
    out = {
        "source": "derived, not reference-attested: CPU oracle (oracle/) on synth.bytes_np(4, 2^32)",
        "workload": f"proof-of-storage commit, 4 GiB file, WriteableFt63, {comm.n_rows} rows x {PRE} -> {ENC}; {OPENS} columns (seed 1337)",
        "n_rows": int(comm.n_rows), "root": comm.get_root().hex(), "columns": [int(c) for c in cols],
        "leaves_sha256": hashlib.sha256(np.ascontiguousarray(leaves).tobytes()).hexdigest(),
        "column_values_sha256": hashlib.sha256(np.ascontiguousarray(vals).tobytes()).hexdigest(),
        "paths_sha256": hashlib.sha256(np.ascontiguousarray(paths).tobytes()).hexdigest(),
    }
    with open(os.path.join(HERE, "pos_4gib.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print(out["root"], out["n_rows"])


if __name__ == "__main__":
    main()
