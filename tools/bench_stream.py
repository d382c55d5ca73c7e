#!/usr/bin/env python
"""Streaming proof-of-storage commit of a synthetic file on ONE GPU (lcpc_stream_*: O(block) device memory),
with and without the column-major encoded-file image (the .porenc sink).  BASELINE.json configs[3] names a 4 GiB file
sharded over 8 GPUs (tools/bench_pos.py); this is the same file through one device, as the reference's
EncodedFileWriter::convert_unencoded_file processes it.

    python tools/bench_stream.py [--gib 4] [--push-mib 256] [--sink-gib 0.5]
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

import lcpc_proof_of_storage_b200 as P
from lcpc_proof_of_storage_b200 import _lib, pos


def run(n_bytes, push_bytes, with_sink, host):
    lib = _lib.load()
    pre, enc_cols, _ = pos.get_aspect_ratio_default_from_file_len(n_bytes)
    n_rows = -(-(-(-n_bytes // 7)) // pre)
    enc = P.LigeroEncoding(P.FT63, pre, enc_cols)
    row_bytes = 7 * pre
    push_bytes = max(row_bytes, push_bytes // row_bytes * row_bytes)
    sink = None
    if with_sink:
        sink = torch.zeros(2 * n_rows * enc_cols * 8, dtype=torch.uint8)
    s = C.c_void_p()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(lib.lcpc_stream_begin(enc.plan, n_rows, 0, sink.data_ptr() if with_sink else None, 2 * n_rows if with_sink else 0,
                                     C.byref(s)))
    t1 = time.perf_counter()
    off = 0
    while off < n_bytes:
        take = min(push_bytes, n_bytes - off)
        _lib.check(lib.lcpc_stream_push_bytes_host(s, host.data_ptr() + off, take))
        off += take
    hashes = np.empty((2 * enc_cols - 1, 32), dtype=np.uint8)
    rows = C.c_size_t()
    _lib.check(lib.lcpc_stream_finish(s, hashes.ctypes.data, C.byref(rows)))
    t2 = time.perf_counter()
    lib.lcpc_stream_free(s)
    return {"file_GiB": n_bytes / (1 << 30), "shape": [rows.value, pre, enc_cols], "sink": with_sink,
            "setup_s": round(t1 - t0, 3), "stream_s": round(t2 - t1, 4), "file_GBps": n_bytes / (t2 - t1) / 1e9,
            "elements_per_s": (n_bytes / 7) / (t2 - t1), "root": bytes(hashes[-1]).hex()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=4.0)
    ap.add_argument("--push-mib", type=int, default=256)
    ap.add_argument("--sink-gib", type=float, default=0.5)
    args = ap.parse_args()
    n_bytes = int(args.gib * (1 << 30))
    rng = np.random.default_rng(4)
    tile = torch.from_numpy(rng.integers(0, 256, 1 << 26, dtype=np.uint8))
    host = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    for o in range(0, n_bytes, 1 << 26):
        n = min(1 << 26, n_bytes - o)
        host[o:o + n] = tile[:n]
        host[o] = (o >> 26) & 0xFF  # tiles differ
    run(min(n_bytes, 1 << 28), args.push_mib << 20, False, host)  # warm-up (context, plan, allocator)
    out = [run(n_bytes, args.push_mib << 20, False, host)]
    # parity of the streamed root against the resident commit (same library, different path) at a size that fits
    small = min(n_bytes, 1 << 28)
    pre, enc_cols, _ = pos.get_aspect_ratio_default_from_file_len(small)
    c = P.LcCommit.commit_bytes(bytes(host[:small].numpy()), P.LigeroEncoding(P.FT63, pre, enc_cols), download=False)
    out.append({"check": "streamed root == resident commit root", "file_GiB": small / (1 << 30),
                "ok": run(small, args.push_mib << 20, False, host)["root"] == c.get_root().hex()})
    if args.sink_gib > 0:
        out.append(run(int(args.sink_gib * (1 << 30)), args.push_mib << 20, True, host))
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
