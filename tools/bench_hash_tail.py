#!/usr/bin/env python
"""Timing of the column-hash + Merkle tail variants on one GPU (2^24 Ft63 bench shape by default):
  old   = k_hash_chunks + k_hash_merge + k_merkle_levels x2 (lcpc_dev_hash_columns + lcpc_dev_merkle_tree)
  tree  = k_hash_tree with LCPC_HT_GROUP in the given list (lcpc_dev_merkleize)
Prints ms per call (CUDA events, 50 calls after 5 warm-ups) and checks that every variant gives the same tree."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lcpc_proof_of_storage_b200 as P
from lcpc_proof_of_storage_b200 import _lib


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    n_rows, n_cols = int(sys.argv[1]) if len(sys.argv) > 1 else 512, int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    groups = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "1,8,32,64,128,0".split(","))]
    lib = _lib.load()
    ctx = P.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(1)
    mat = torch.randint(0, 5102708120182849537, (n_rows * n_cols,), dtype=torch.int64, device="cuda", generator=g)
    np2 = P.next_pow2(n_cols)
    ref = torch.zeros((2 * np2 - 1) * 32, dtype=torch.uint8, device="cuda")

    def old():
        _lib.check(lib.lcpc_dev_hash_columns(ctx.handle, 0, mat.data_ptr(), n_rows, n_cols, n_cols, ref.data_ptr()))
        _lib.check(lib.lcpc_dev_merkle_tree(ctx.handle, ref.data_ptr(), np2))

    print(f"old (4 launches): {timeit(old):.4f} ms")
    os.environ["LCPC_HASH_MODE"] = "1"
    out1 = torch.zeros_like(ref)

    def two():
        _lib.check(lib.lcpc_dev_merkleize(ctx.handle, 0, mat.data_ptr(), n_rows, n_cols, n_cols, out1.data_ptr()))

    print(f"k_hash_chunks + k_merge_tree: {timeit(two):.4f} ms  same tree: {bool(torch.equal(out1, ref))}")
    os.environ["LCPC_HASH_MODE"] = "2"
    for grp in groups:
        os.environ["LCPC_HT_GROUP"] = str(grp)
        out = torch.zeros_like(ref)

        def new():
            _lib.check(lib.lcpc_dev_merkleize(ctx.handle, 0, mat.data_ptr(), n_rows, n_cols, n_cols, out.data_ptr()))

        t = timeit(new)
        print(f"k_hash_tree group={grp}: {t:.4f} ms  same tree: {bool(torch.equal(out, ref))}")


if __name__ == "__main__":
    main()
