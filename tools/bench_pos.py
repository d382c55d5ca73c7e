#!/usr/bin/env python
"""BASELINE.json configs[3]: proof-of-storage commit of a synthetic file sharded over N GPUs, plus a
random-column retrievability proof (open 309 columns with paths) and its verification on rank 0.

    torchrun --nproc-per-node N tools/bench_pos.py [--gib 4] [--steps 5]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import lcpc_proof_of_storage_b200 as P
from lcpc_proof_of_storage_b200 import pos
from lcpc_proof_of_storage_b200.sharded import ShardedLigeroCommitter


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=4.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--hashing", default="auto", choices=["auto", "columns", "rows"])
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local))
    n_bytes = int(args.gib * (1 << 30))
    pre, enc_cols, soundness = pos.get_aspect_ratio_default_from_file_len(n_bytes)
    n_elems = (n_bytes + 6) // 7
    n_rows = (n_elems + pre - 1) // pre
    stream = torch.cuda.current_stream()
    ctx = P.Context(local, stream=stream.cuda_stream)
    enc = P.LigeroEncoding(P.FT63, pre, enc_cols, ctx=ctx)
    sc = ShardedLigeroCommitter(enc, n_rows, None, hashing=args.hashing)
    lo, hi = sc.byte_range(n_bytes)
    # the file is the seeded stream of synth.py (seed 4), the same bytes at any number of GPUs: the root and the opened
    # columns of the 4 GiB file are checked against tests/golden/pos_4gib.json (CPU oracle, tests/golden/make_pos_golden.py)
    from lcpc_proof_of_storage_b200 import synth as S

    lo8 = lo - lo % 8
    data = S.bytes_torch(4, hi - lo8, torch.device("cuda", local), lo8)[lo - lo8:].clone()
    for _ in range(2):
        sc.commit_bytes(data)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        sc.commit_bytes(data)
    e1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    cols = pos.get_column_indicies_from_random_seed(1337, soundness, enc_cols)
    sc.open_columns(cols)  # first call: NCCL channel set-up, first-use allocations
    t_open_dev = None
    if sc.hashing == "rows":
        # the retrievability proof on the device: gather kernels + one all-gather, result left in rank 0's HBM
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(5):
            sc.open_columns_dev(cols)
        e3.record(stream)
        torch.cuda.synchronize()
        t_open_dev = e2.elapsed_time(e3) / 5
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    opened = sc.open_columns(cols)   # + one device -> host copy of the column values and paths
    torch.cuda.synchronize()
    t_open = time.perf_counter() - t0
    if rank == 0:
        import hashlib

        root = sc.root()
        t0 = time.perf_counter()
        pos.client_online_verify_column_paths(root, cols, opened, ctx)
        t_verify = time.perf_counter() - t0
        line = {"case": "pos_commit", "n_gpus": world, "file_GiB": args.gib, "shape": [n_rows, pre, enc_cols],
                "hashing": sc.hashing, "cv_fused": sc.cv_fused, "fused_nvlink": sc.fused, "ms_per_commit": round(ms, 3),
                "file_GBps": n_bytes / ms / 1e6, "elements_per_s": n_elems / ms * 1e3,
                "open_columns": len(cols), "open_on_device_ms": None if t_open_dev is None else round(t_open_dev, 3),
                "open_to_host_ms": round(t_open * 1e3, 3), "open_bytes": len(cols) * (n_rows * 8 + opened[0].path.size),
                "verify_paths_ms": round(t_verify * 1e3, 3), "root": root.hex()}
        try:
            with open(os.path.join(ROOT, "tests", "golden", "pos_4gib.json")) as f:
                want = json.load(f)
        except OSError:
            want = None
        if want and args.gib == 4.0:
            vals = np.stack([o.col for o in opened])
            paths = np.stack([o.path for o in opened])
            line["matches_known_answer"] = {
                "root": root.hex() == want["root"], "columns": cols == want["columns"],
                "column_values": hashlib.sha256(np.ascontiguousarray(vals).tobytes()).hexdigest() == want["column_values_sha256"],
                "paths": hashlib.sha256(np.ascontiguousarray(paths).tobytes()).hexdigest() == want["paths_sha256"]}
            lv, _ = pos._verify_columns(ctx, opened, None, None)
            line["matches_known_answer"]["leaves"] = hashlib.sha256(np.ascontiguousarray(lv).tobytes()).hexdigest() == want["leaves_sha256"]
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
