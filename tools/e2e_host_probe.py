#!/usr/bin/env python
"""lcpc_commit_host on bench.py's 2^24 workload (pinned host coefficients in, encoded matrix + tree out) under different
row-chunk schedules of the PCIe-in / encode / PCIe-out pipeline (LCPC_COMMIT_CHUNKS, LCPC_COMMIT_RAMP)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import lcpc_proof_of_storage_b200 as P
from lcpc_proof_of_storage_b200 import _lib
from lcpc_proof_of_storage_b200 import synth as S

lib = _lib.load()
ctx = P.Context(0)
enc = P.LigeroEncoding(0, bench.N_PER_ROW, bench.N_COLS, ctx=ctx)
n = 1 << 24
h = torch.from_numpy(S.ft63_np(2, n).view(np.int64).reshape(-1)).pin_memory()
h_comm = torch.empty(512 * 65536, dtype=torch.int64).pin_memory()
h_hashes = torch.empty((2 * 65536 - 1) * 32, dtype=torch.uint8).pin_memory()
for chunks, ramp in [(8, 0), (8, 1), (12, 1), (16, 0), (16, 1), (6, 1), (4, 1)]:
    os.environ["LCPC_COMMIT_CHUNKS"], os.environ["LCPC_COMMIT_RAMP"] = str(chunks), str(ramp)

    def step():
        _lib.check(lib.lcpc_commit_host(enc.plan, h.data_ptr(), n, None, h_comm.data_ptr(), h_hashes.data_ptr(), None))

    for _ in range(3):
        step()
    t0 = time.perf_counter()
    for _ in range(10):
        step()
    dt = (time.perf_counter() - t0) / 10
    print(f"chunks {chunks} ramp {ramp}: {dt * 1e3:.3f} ms  root {bytes(h_hashes[-32:].numpy()).hex()[:16]}")
