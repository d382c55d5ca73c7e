#!/usr/bin/env python
"""The one-process, N-device path of the C ABI (lcpc_ctx_create_multi) on bench.py's weak-scaling workload: N x 512 rows x
32768 -> 65536 over Ft63, committed through lcpc_commit_host from pinned host memory, root read back; the root is checked
against tests/golden/bench_roots.json.  One JSON line.

    python tools/bench_multi_abi.py [--gpus N] [--steps K]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import bench
import lcpc_proof_of_storage_b200 as P
from lcpc_proof_of_storage_b200 import _lib
from lcpc_proof_of_storage_b200 import synth as S


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    lib = _lib.load()
    n_rows = bench.ROWS_PER_GPU * args.gpus
    n = n_rows * bench.N_PER_ROW
    ctx = P.Context.multi(list(range(args.gpus)))
    enc = P.LigeroEncoding(bench.FID, bench.N_PER_ROW, bench.N_COLS, ctx=ctx)
    h = torch.from_numpy(S.ft63_np(2, n).view(np.int64).reshape(-1)).pin_memory()
    root = torch.empty(32, dtype=torch.uint8).pin_memory()
    keep = C.c_void_p()

    def step():
        _lib.check(lib.lcpc_commit_host(enc.plan, h.data_ptr(), n, None, None, None, C.byref(keep)))
        _lib.check(lib.lcpc_commit_root(keep, root.data_ptr()))
        lib.lcpc_commit_free(keep)

    for _ in range(3):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    got = bytes(root.numpy()).hex()
    print(json.dumps({"case": "multi_device_abi_commit", "n_gpus": args.gpus, "api": "lcpc_ctx_create_multi + lcpc_commit_host(keep) + "
                      "lcpc_commit_root (one process; pinned host coefficients in, root out)", "workload": bench.workload_name(args.gpus),
                      "ms_per_commit": dt * 1e3, "elements_per_s": n / dt, "h2d_bytes": n * 8, "root": got,
                      "root_matches_golden": got == bench.expected_root(args.gpus)}))


if __name__ == "__main__":
    main()
