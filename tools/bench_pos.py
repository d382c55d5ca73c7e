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
    ap.add_argument("--hashing", default="columns", choices=["columns", "rows"])
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
    g = torch.Generator(device="cuda")
    g.manual_seed(4 + rank)
    data = torch.randint(0, 256, (hi - lo,), dtype=torch.uint8, device="cuda", generator=g)
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
    sc.open_columns(cols)  # first call: NCCL sets up the point-to-point channels to rank 0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    opened = sc.open_columns(cols)
    torch.cuda.synchronize()
    t_open = time.perf_counter() - t0
    if rank == 0:
        root = sc.root()
        t0 = time.perf_counter()
        pos.client_online_verify_column_paths(root, cols, opened, ctx)
        t_verify = time.perf_counter() - t0
        print(json.dumps({"case": "pos_commit", "n_gpus": world, "file_GiB": args.gib, "shape": [n_rows, pre, enc_cols],
                          "fused_nvlink": sc.fused, "ms_per_commit": round(ms, 3), "file_GBps": n_bytes / ms / 1e6,
                          "elements_per_s": n_elems / ms * 1e3, "open_309_columns_s": round(t_open, 3),
                          "verify_309_paths_s": round(t_verify, 4), "root": root.hex()}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
