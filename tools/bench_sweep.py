#!/usr/bin/env python
"""BASELINE.json configs[4]: scaling sweep 2^20 - 2^28 coefficients, Ligero and Brakedown commit + the eval-proof row
folds ((n_degree_tests + 1) sequential single-tensor folds, as `prove` issues them), on N GPUs (one process per GPU),
strong scaling: ONE polynomial of 2^n coefficients whose rows are sharded over the ranks.

    python tools/bench_sweep.py                         # 1 GPU
    torchrun --nproc-per-node N tools/bench_sweep.py [--log-n 20 22 24 26 28] [--schemes ligero63 brakedown63 brakedown255]

One JSON line per (scheme, n).  Times are CUDA events on the launching stream, max over ranks.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import lcpc_proof_of_storage_b200 as P
from lcpc_proof_of_storage_b200.sharded import ShardedCommitter

MOD_TOP = {0: 0x46d07600, 3: 0x663c799b}


def rand_elems(fid, n, seed):
    L = P.FIELD_LIMBS[fid]
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    a = torch.randint(-(1 << 63), (1 << 63) - 1, (n, L), dtype=torch.int64, device="cuda", generator=g)
    # top limb below the modulus' top word: reduced, uniform enough for timing
    a[:, L - 1] = torch.randint(0, MOD_TOP[fid] << 32, (n,), dtype=torch.int64, device="cuda", generator=g)
    return a.reshape(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", dest="logs", type=int, nargs="*", default=[20, 22, 24, 26, 28])
    ap.add_argument("--schemes", nargs="*", default=["ligero63", "brakedown63", "brakedown255"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--hashing", default="auto", choices=["auto", "columns", "rows"],
                    help="rows: hash BLAKE3 chunks where the rows are and re-shard chaining values (sharded.py)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29544")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()
    ctx = P.Context(local, stream=stream.cuda_stream)
    for scheme in args.schemes:
        fid = 3 if scheme.endswith("255") else 0
        L = P.FIELD_LIMBS[fid]
        for log_n in args.logs:
            n = 1 << log_n
            if scheme.startswith("ligero"):
                enc = P.LigeroEncoding.new(fid, n, ctx=ctx)
            else:
                enc = P.SdigEncoding.new(fid, n, seed=0, ctx=ctx)
            n_rows, npr, n_cols = enc.get_dims(n)
            if n_rows * n_cols * L * 8 * 2.5 / world > 150e9:  # encoded shard + exchange buffers must fit in HBM
                if rank == 0:
                    print(json.dumps({"case": f"{scheme}_2^{log_n}", "n_gpus": world, "skipped": "does not fit"}), flush=True)
                continue
            try:
                sc = ShardedCommitter(enc, n_rows, None, hashing=args.hashing)
            except ValueError:  # rows mode needs multi-chunk leaves and elements that divide a chunk
                sc = ShardedCommitter(enc, n_rows, None)
            r0, cnt = sc.rows[rank]
            coeffs = rand_elems(fid, max(cnt, 1) * npr, 100 + log_n + rank)[:cnt * npr * L]
            n_dt = enc.get_n_degree_tests()
            tensors = [rand_elems(fid, n_rows, 7 + i) for i in range(n_dt + 1)]
            if world > 1:
                for t in tensors:
                    dist.broadcast(t, 0)

            def timed(fn):
                for _ in range(2):
                    fn()
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(args.steps):
                    fn()
                e1.record(stream)
                dist.barrier()
                torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                return float(ms.item())

            ms_commit = timed(lambda: sc.commit(coeffs))

            def folds():
                for t in tensors:
                    sc.fold(t)

            ms_fold = timed(folds)
            if rank == 0:
                w = 8 * L
                alg = n * w + n_rows * n_cols * w + (2 * P.next_pow2(n_cols) - 1) * 32
                print(json.dumps({"case": f"{scheme}_2^{log_n}", "n_gpus": world, "shape": [n_rows, npr, n_cols],
                                  "hashing": sc.hashing, "cv_fused": sc.cv_fused, "fused_nvlink": sc.fused, "ms_commit": round(ms_commit, 4), "coeffs_per_s": n / ms_commit * 1e3,
                                  "commit_algorithmic_GBps": alg / ms_commit / 1e6, "n_folds": n_dt + 1,
                                  "ms_folds": round(ms_fold, 4), "fold_GBps": (n_dt + 1) * n * w / ms_fold / 1e6,
                                  "root": sc.root().hex()}), flush=True)
            del sc, coeffs, tensors, enc
            torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
