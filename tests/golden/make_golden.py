#!/usr/bin/env python
"""Generates tests/golden/vectors.json -- the committed known-answer fixtures of the commit path.

    python tests/golden/make_golden.py

The reference's own tests hold no golden vectors for this path and the Rust
reference cannot be built in this image, so the fixtures come from two sources, labelled per entry:

* "independent": computed WITHOUT the oracle -- BLAKE3 digests from the Python `blake3` package (bindings
  to the Rust crate the reference links), and NTT / Montgomery values from Python big-integer arithmetic
  following the definitions (out[bitrev(i)] = sum_j in[j] w^(ij), w = ROOT_OF_UNITY^(2^(S-k)); repr = a R^-1).
  A small Ligero commitment (encode + leaves + tree) is also rebuilt from those two alone.
* "derived": produced by the CPU oracle (oracle/) on seeded inputs; "derived, not reference-attested".
  They pin today's behaviour of oracle AND CUDA path against accidental change.

Both the oracle (`-m "not gpu"`) and the CUDA library (`-m gpu`) are tested against this file
(tests/test_golden.py).  Nothing here reads /root/reference.
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

# field id -> (modulus, limbs, two-adicity S, multiplicative generator); lcpc-test-fields/src/lib.rs:18-70
FIELDS = {
    0: (0x46d0760000000001, 1, 41, 10),
    1: (0x6e754097ba20e0bf7f2bd90000000001, 2, 40, 3),
    2: (0x453708aa3fbc8dda936888270ceecbcdd246820000000001, 3, 41, 5),
    3: (0x663c799b6e4d2900fda9df04b9575969ef73c79086595f3002a4f20000000001, 4, 41, 5),
    # Ft253_192, proof-of-storage/src/fields/ft253_192.rs:6-10: (2^61 - 1) * 2^192 + 1, big-endian repr
    4: (0x1fffffffffffffff000000000000000000000000000000000000000000000001, 4, 192, 3),
}
REPR_BYTE_ORDER = {0: "little", 1: "little", 2: "little", 3: "little", 4: "big"}  # PrimeFieldReprEndianness


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes() if isinstance(a, np.ndarray) else a).hexdigest()


def bitrev(i: int, bits: int) -> int:
    return int(format(i, f"0{bits}b")[::-1], 2) if bits else 0


def pattern(n: int) -> bytes:
    return bytes(i % 251 for i in range(n))


def lcg_values(seed: int, n: int, p: int):
    """Tiny deterministic generator of canonical field values (Python ints only)."""
    x, out = seed * 0x9E3779B97F4A7C15 + 12345, []
    for _ in range(n):
        v = 0
        for _ in range(5):
            x = (x * 6364136223846793005 + 1442695040888963407) % (1 << 64)
            v = (v << 64) | x
        out.append(v % p)
    return out


def independent_section():
    import blake3

    out = {"blake3": [], "ntt": [], "ligero_commit": []}
    # operand pairs on which a dual-accumulator Montgomery product that adds a chain-end carry into a data word loses
    # that carry (found by find_carry_kats.py with 32-bit words; expected products are Python integers)
    with open(os.path.join(HERE, "carry_kats.json")) as f:
        out["carry_chain_mul"] = json.load(f)
    for n in [0, 1, 63, 64, 65, 1023, 1024, 1025, 2048, 3072, 4128, 5000, 8192, 31744 + 32]:
        out["blake3"].append({"len": n, "input": "bytes(i % 251 for i in range(len))", "digest": blake3.blake3(pattern(n)).hexdigest()})
    for fid, (p, limbs, S, g) in FIELDS.items():
        R = 1 << (64 * limbs)
        t = (p - 1) >> S
        root = pow(g, t, p)
        for k in (3, 5):
            n = 1 << k
            w = pow(root, 1 << (S - k), p)
            vals = lcg_values(100 * fid + k, n, p)
            res = [0] * n
            for i in range(n):
                res[bitrev(i, k)] = sum(vals[j] * pow(w, i * j, p) for j in range(n)) % p
            out["ntt"].append({"fid": fid, "log_n": k, "input_canonical": [hex(v) for v in vals],
                               "input_montgomery": [hex(v * R % p) for v in vals],
                               "output_montgomery": [hex(v * R % p) for v in res]})
        # a whole Ligero commitment from definitions: 3 rows x 4 -> 8 columns, 10 coefficients (ragged last row)
        n_rows, npr, nc, k = 3, 4, 8, 3
        w = pow(root, 1 << (S - k), p)
        coeffs = lcg_values(7 + fid, 10, p)
        padded = coeffs + [0] * (n_rows * npr - len(coeffs))
        comm = []
        for r in range(n_rows):
            row = padded[r * npr:(r + 1) * npr] + [0] * (nc - npr)
            enc = [0] * nc
            for i in range(nc):
                enc[bitrev(i, k)] = sum(row[j] * pow(w, i * j, p) for j in range(nc)) % p
            comm.append(enc)
        # leaf = BLAKE3(0^32 || to_repr() of the column's elements): the canonical value, little-endian (big-endian for
        # Ft253_192); the stored (Montgomery) limbs are value*R mod p, so the repr is the integer "value" itself
        leaves = []
        for c in range(nc):
            h = blake3.blake3()
            h.update(bytes(32))
            for r in range(n_rows):
                h.update(comm[r][c].to_bytes(8 * limbs, REPR_BYTE_ORDER[fid]))
            leaves.append(h.digest())
        level, tree = leaves, list(leaves)
        while len(level) > 1:
            level = [blake3.blake3(level[2 * i] + level[2 * i + 1]).digest() for i in range(len(level) // 2)]
            tree += level
        out["ligero_commit"].append({"fid": fid, "n_rows": n_rows, "n_per_row": npr, "n_cols": nc,
                                     "coeffs_montgomery": [hex(v * R % p) for v in coeffs],
                                     "comm_montgomery": [[hex(v * R % p) for v in row] for row in comm],
                                     "hashes": [d.hex() for d in tree], "root": tree[-1].hex()})
    return out


def derived_section():
    from oracle import lcpc_oracle as O

    O.build()
    out = {"ligero": [], "brakedown": [], "pos_bytes": [], "prove": []}
    for fid, n, npr, nc in [(0, 1 << 16, 2048, 4096), (0, 125 * 64, 64, 128), (0, 3000, 100, 256), (1, 5000, 100, 256),
                            (2, 3000, 60, 128), (3, 5000, 100, 256), (0, 700 * 32, 32, 64), (4, 5000, 100, 256)]:
        coeffs = O.random_field_elements(fid, 1000 + fid, n)
        c = O.commit(coeffs, O.LigeroEncoding(fid, npr, nc))
        tensor = O.random_field_elements(fid, 2000 + fid, c.n_rows)
        col = O.open_column(c, nc // 3)
        out["ligero"].append({"fid": fid, "n": n, "n_per_row": npr, "n_cols": nc, "coeff_seed": 1000 + fid, "tensor_seed": 2000 + fid,
                              "root": c.get_root().hex(), "comm_sha256": sha(c.comm), "hashes_sha256": sha(c.hashes),
                              "fold_sha256": sha(O.collapse_columns(fid, c.coeffs, tensor)), "open_column": nc // 3,
                              "open_col_sha256": sha(col.col), "open_path_sha256": sha(col.path)})
    for fid, npr, n_rows, seed in [(0, 150, 20, 0), (3, 150, 7, 1), (1, 400, 5, 0), (4, 150, 7, 1)]:
        enc = O.SdigEncoding(fid, npr, seed)
        coeffs = O.random_field_elements(fid, 41, n_rows * npr - 3)
        c = O.commit(coeffs, enc)
        out["brakedown"].append({"fid": fid, "n_per_row": npr, "n_rows": n_rows, "code_seed": seed, "coeff_seed": 41,
                                 "n_cols": enc.n_cols, "precode_data_sha256": [sha(m.data) for m in enc.precodes],
                                 "postcode_indices_sha256": [sha(m.indices) for m in enc.postcodes],
                                 "root": c.get_root().hex(), "comm_sha256": sha(c.comm), "hashes_sha256": sha(c.hashes)})
    for n_bytes, npr, nc in [(598, 4, 8), (100003, 64, 128)]:
        data = pattern(n_bytes)
        c = O.commit(O.pack_bytes7(data), O.LigeroEncoding(0, npr, nc))
        out["pos_bytes"].append({"n_bytes": n_bytes, "n_per_row": npr, "n_cols": nc, "input": "bytes(i % 251 for i in range(n_bytes))",
                                 "root": c.get_root().hex(), "hashes_sha256": sha(c.hashes), "coeffs_sha256": sha(c.coeffs)})
    # full prove transcripts: commitment, challenges and openings all hashed (Ft63; Ft253_192 for the big-endian
    # transcript messages)
    for fid, n in [(0, 1 << 12), (4, 1 << 10)]:
        _prove_case(O, out, fid, n)
    return out


def _prove_case(O, out, fid, n):
    enc = O.LigeroEncoding.new(fid, n)
    coeffs = O.random_field_elements(fid, 5, n)
    c = O.commit(coeffs, enc)
    x = O.random_field_elements(fid, 6, 1)
    outer = O.random_field_elements(fid, 8, c.n_rows)
    tr = O.Transcript(b"golden")
    tr.append_message(b"polycommit", c.get_root())
    tr.append_message(b"rate", b"0.5")
    proof = O.prove(c, outer, enc, tr)
    out["prove"].append({"fid": fid, "n": n, "coeff_seed": 5, "outer_seed": 8, "transcript_label": "golden",
                         "messages": [["polycommit", "<root>"], ["rate", "0.5"]], "root": c.get_root().hex(),
                         "p_eval_sha256": sha(proof.p_eval), "p_random_sha256": sha(np.stack(proof.p_random_vec)),
                         "columns_sha256": sha(np.stack([col.col for col in proof.columns])),
                         "paths_sha256": sha(np.stack([col.path for col in proof.columns])),
                         "challenge_after_prove": tr.challenge_bytes(b"check", 32).hex()})
    del x


def main():
    doc = {"about": "known-answer fixtures for the lcpc commit path; see make_golden.py (independent = no oracle involved; "
                    "derived = CPU oracle output, not reference-attested)",
           "independent": independent_section(), "derived": derived_section()}
    path = os.path.join(HERE, "vectors.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
