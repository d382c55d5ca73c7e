import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch, torch.distributed as dist
import lcpc_proof_of_storage_b200 as P
from lcpc_proof_of_storage_b200 import synth as S
from lcpc_proof_of_storage_b200.sharded import ShardedLigeroCommitter
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29544")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
use_side = len(sys.argv) > 1 and sys.argv[1] == "side"
if use_side:
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
stream = torch.cuda.current_stream()
ctx = P.Context(0, stream=stream.cuda_stream)
enc = P.LigeroEncoding(0, 32768, 65536, ctx=ctx)
sc = ShardedLigeroCommitter(enc, 512, dist.group.WORLD, hashing="rows")
h = torch.from_numpy(S.ft63_np(2, 1 << 24).view(np.int64).reshape(-1)).pin_memory()
h_comm = torch.empty(512 * 65536, dtype=torch.int64).pin_memory()
for full in (True, False):
    for chunks in (8, 4, 16):
        for _ in range(3):
            sc.commit_host(h, h_comm if full else None, chunks); sc.wait_host_copies(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            sc.commit_host(h, h_comm if full else None, chunks); sc.wait_host_copies(); torch.cuda.synchronize()
        print("side" if use_side else "default", "full" if full else "root", chunks, (time.perf_counter() - t0) / 10 * 1e3, "ms", sc.root().hex()[:16])
dist.destroy_process_group()
