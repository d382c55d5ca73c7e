#!/usr/bin/env python
"""k_merge_tree alone (leaf merge + Merkle tree, one launch) on the shapes a rank sees: n_chunks x n_cols chaining values.
LCPC_MERGE_PAR: 0 = per-thread walk, 4 / 8 / 16 / 32 = level-wise merge with that many columns per CTA."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lcpc_proof_of_storage_b200 as P
from lcpc_proof_of_storage_b200 import _lib

lib = _lib.load()
ctx = P.Context(0, stream=torch.cuda.current_stream().cuda_stream)
for n_chunks, n_cols in [(5, 65536), (33, 8192), (17, 16384), (9, 32768), (147, 8192), (147, 65536), (17, 262144)]:
    cvs = torch.randint(0, 256, (n_chunks * n_cols * 32,), dtype=torch.uint8, device="cuda")
    ref = None
    line = f"{n_chunks:4d} chunks x {n_cols:6d} cols:"
    for mode in ("0", "4", "8", "16", "32"):
        os.environ["LCPC_MERGE_PAR"] = mode
        tree = torch.zeros((2 * n_cols - 1) * 32, dtype=torch.uint8, device="cuda")

        def fn():
            _lib.check(lib.lcpc_dev_hash_merge_tree(ctx.handle, cvs.data_ptr(), n_cols, n_chunks, tree.data_ptr(), n_cols, 0))

        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(20):
            fn()
        b.record()
        torch.cuda.synchronize()
        if ref is None:
            ref = tree.clone()
        line += f"  par={mode}: {a.elapsed_time(b) / 20 * 1e3:7.1f} us{'' if torch.equal(tree, ref) else ' MISMATCH'}"
    print(line)
