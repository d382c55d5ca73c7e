"""The one-process-per-GPU commit on real GPUs (needs >= 2 devices; the single-GPU round-end run
skips it): NCCL all-to-all path and the fused encode + peer-store path, both against the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, fused, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import lcpc_proof_of_storage_b200 as P
        from lcpc_proof_of_storage_b200.sharded import ShardedLigeroCommitter, row_partition
        from oracle import lcpc_oracle as O

        fid, n_rows, n_per_row, n_cols = 0, 37, 4096, 8192
        n = n_rows * n_per_row - 11
        coeffs = np.zeros((n_rows * n_per_row, 1), dtype=np.uint64)
        coeffs[:n] = O.random_field_elements(fid, 5, n)
        ctx = P.Context(rank, stream=torch.cuda.current_stream().cuda_stream)
        enc = P.LigeroEncoding(fid, n_per_row, n_cols, ctx=ctx)
        sc = ShardedLigeroCommitter(enc, n_rows, None, fused=fused)
        assert sc.fused == fused
        r0, cnt = row_partition(n_rows, world)[rank]
        local = torch.from_numpy(coeffs.reshape(n_rows, n_per_row)[r0:r0 + cnt].copy().view(np.int64).reshape(-1)).cuda()
        for _ in range(3):  # repeated commits reuse the symmetric buffers
            sc.commit(local)
        root_elems = sc.root() if rank == 0 else None
        # the same commitment from file bytes (each element < 2^56, so packing 7-byte groups reproduces it)
        file_bytes = (coeffs[:n].reshape(-1) & np.uint64((1 << 56) - 1)).view(np.uint8).reshape(-1, 8)[:, :7].reshape(-1)
        lo, hi = sc.byte_range(file_bytes.shape[0])
        sc2 = ShardedLigeroCommitter(enc, n_rows, None, fused=fused)
        sc2.commit_bytes(torch.from_numpy(file_bytes[lo:hi].copy()).cuda())
        if rank == 0:
            exp2 = O.commit(O.pack_bytes7(file_bytes.tobytes()), O.LigeroEncoding(fid, n_per_row, n_cols))
            assert sc2.root() == exp2.get_root()
        if fused:
            # deferred commits: the tree of commit k appears once commit k+1 has issued its encode, or at flush();
            # three symmetric buffers rotate, so run more than three and alternate two inputs
            coeffs_b = np.zeros_like(coeffs)
            coeffs_b[:n] = O.random_field_elements(fid, 6, n)
            local_b = torch.from_numpy(coeffs_b.reshape(n_rows, n_per_row)[r0:r0 + cnt].copy().view(np.int64).reshape(-1)).cuda()
            root_b = O.commit(coeffs_b[:n], O.LigeroEncoding(fid, n_per_row, n_cols)).get_root() if rank == 0 else None
            seen = []
            for i in range(7):
                sc.commit(local_b if i % 2 else local, defer=True)
                if i and rank == 0:  # commit i-1 is finished now
                    seen.append(bytes(sc.top[-32:].cpu().numpy()))
            sc.flush()
            if rank == 0:
                seen.append(sc.root())
                assert seen == [root_b if i % 2 else root_elems for i in range(7)]
            with_pending_ok = True
            sc.commit(local, defer=True)
            try:
                if rank == 0:
                    sc.root()
                    with_pending_ok = False
            except AssertionError:
                pass
            sc.flush()
            assert with_pending_ok
        sc.commit(local)
        hashes = sc.gather_hashes()
        tensors = O.random_field_elements(fid, 7, 2 * n_rows).reshape(2, n_rows, 1)
        t_dev = torch.from_numpy(tensors.view(np.int64).reshape(-1).copy()).cuda()
        folded = sc.fold(t_dev)
        folded_enc = sc.fold_encoded(t_dev)  # the proof-of-storage fold over the encoded matrix
        cols = [0, n_cols - 1, 4096, 4095]
        opened = sc.open_columns(cols)
        if rank == 0:
            exp = O.commit(coeffs[:n], O.LigeroEncoding(fid, n_per_row, n_cols))
            ok = sc.root() == exp.get_root()
            ok &= np.array_equal(hashes.cpu().numpy().reshape(-1, 32), exp.hashes)
            f = folded.cpu().numpy().view(np.uint64).reshape(2, n_per_row, 1)
            fe = folded_enc.cpu().numpy().view(np.uint64).reshape(2, n_cols, 1)
            for t in range(2):
                ok &= np.array_equal(f[t], O.collapse_columns(fid, exp.coeffs, tensors[t]))
                ok &= np.array_equal(fe[t], O.collapse_columns(fid, exp.comm, tensors[t]))
            for c, col in zip(cols, opened):
                e = O.open_column(exp, c)
                ok &= np.array_equal(col.col, e.col) and np.array_equal(col.path, e.path)
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def _worker_brakedown(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import lcpc_proof_of_storage_b200 as P
        from lcpc_proof_of_storage_b200.sharded import ShardedCommitter, row_partition
        from oracle import lcpc_oracle as O

        ok = True
        # both exchanges of the column-block mode: the transposing passes storing into the owners' blocks of the padded
        # column range (fused: symmetric memory, NVLink), and the NCCL all-to-all over the padded leaf range
        for fid, n_per_row, n_rows, seed, fused in [(0, 300, 23, 0, None), (3, 150, 9, 1, None), (0, 300, 23, 0, False),
                                                    (1, 260, 17, 2, None)]:
            L = O.LIMBS[fid]
            oenc = O.SdigEncoding(fid, n_per_row, seed)
            n = n_rows * n_per_row - 7
            coeffs = np.zeros((n_rows * n_per_row, L), dtype=np.uint64)
            coeffs[:n] = O.random_field_elements(fid, 5, n)
            ctx = P.Context(rank, stream=torch.cuda.current_stream().cuda_stream)
            enc = P.SdigEncoding.new_from_dims(fid, n_per_row, oenc.n_cols, seed=seed, ctx=ctx)
            sc = ShardedCommitter(enc, n_rows, None, fused=fused)
            assert sc.fused == (fused is None)  # n_cols is not a power of two: the blocks are cut from the padded range
            r0, cnt = row_partition(n_rows, world)[rank]
            local = torch.from_numpy(coeffs.reshape(n_rows, n_per_row, L)[r0:r0 + cnt].copy().view(np.int64).reshape(-1)).cuda()
            sc.commit(local)
            hashes = sc.gather_hashes()
            tensors = O.random_field_elements(fid, 7, 2 * n_rows).reshape(2, n_rows, L)
            folded = sc.fold(torch.from_numpy(tensors.view(np.int64).reshape(-1).copy()).cuda())
            cols = [0, oenc.n_cols - 1, oenc.n_cols // 2]
            opened = sc.open_columns(cols)
            if rank == 0:
                exp = O.commit(coeffs[:n], oenc)
                ok &= sc.root() == exp.get_root()
                ok &= np.array_equal(hashes.cpu().numpy().reshape(-1, 32), exp.hashes)
                f = folded.cpu().numpy().view(np.uint64).reshape(2, n_per_row, L)
                for t in range(2):
                    ok &= np.array_equal(f[t], O.collapse_columns(fid, exp.coeffs, tensors[t]))
                for c, col in zip(cols, opened):
                    e = O.open_column(exp, c)
                    ok &= np.array_equal(col.col, e.col) and np.array_equal(col.path, e.path)
        if rank == 0:
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_sharded_brakedown_commit_two_gpus():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_brakedown, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


@pytest.mark.parametrize("fused", [False, True])
def test_sharded_commit_two_gpus(fused):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fused, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


# ----------------------------------------------------------------------------- row-sharded hashing (lower traffic)

@pytest.mark.parametrize("fid,n_rows,n_per_row,n_cols,split_chunk", [
    (0, 300, 64, 128, 1),     # 3 chunks: 124 rows | 128 + 48 rows
    (0, 700, 64, 128, 3),     # 6 chunks, last one partial
    (0, 512, 2048, 4096, 2),  # the bench's rows per GPU: 4 full chunks + a 32-byte tail chunk
    (1, 130, 64, 128, 2),     # 16-byte elements
    (3, 70, 64, 128, 1),      # 32-byte elements: 31 rows in chunk 0
    (4, 70, 64, 128, 2),      # Ft253_192: big-endian repr
])
def test_hash_chunk_range_and_merge_one_gpu(oracle, fid, n_rows, n_per_row, n_cols, split_chunk):
    """lcpc_dev_hash_chunk_range over two row windows that meet on a chunk boundary + lcpc_dev_hash_merge
    == hash_columns over the whole matrix (what two ranks of a row-hashed commit compute between them)."""
    import torch

    import lcpc_proof_of_storage_b200 as P
    from lcpc_proof_of_storage_b200.sharded import GpuOps

    O = oracle
    L = O.LIMBS[fid]
    w = 8 * L
    # GpuOps works on torch tensors: the library context must run on torch's current stream (as sharded.py's callers do)
    ctx = P.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    enc = P.LigeroEncoding(fid, n_per_row, n_cols, ctx=ctx)
    ops = GpuOps(enc)
    coeffs = O.random_field_elements(fid, 17, n_rows * n_per_row)
    exp = O.commit(coeffs, O.LigeroEncoding(fid, n_per_row, n_cols))
    comm = ops.encode(torch.from_numpy(coeffs.view(np.int64).reshape(-1).copy()).cuda(), n_rows)
    assert np.array_equal(comm.cpu().numpy().view(np.uint64).reshape(n_rows, n_cols, L), exp.comm)
    n_chunks = O.leaf_chunks(fid, n_rows)
    row_split = (1024 - 32) // w + (split_chunk - 1) * (1024 // w)
    lo = comm[:row_split * n_cols * L].clone()   # separate allocations, as on two ranks
    hi = comm[row_split * n_cols * L:].clone()
    cv_lo = ops.hash_chunk_range(lo, 0, n_rows, n_cols, 0, split_chunk)
    cv_hi = ops.hash_chunk_range(hi, row_split, n_rows, n_cols, split_chunk, n_chunks)
    exp_cvs = O.hash_chunk_cvs(fid, exp.comm[:, :8], 0, n_rows, 0, n_chunks)  # the oracle's values for 8 columns
    got_cvs = torch.cat([cv_lo, cv_hi]).cpu().numpy().reshape(n_chunks, n_cols, 32)
    assert np.array_equal(got_cvs[:, :8], exp_cvs)
    leaves = torch.zeros(n_cols * 32, dtype=torch.uint8, device="cuda")
    ops.hash_merge(torch.cat([cv_lo, cv_hi]), n_cols, n_chunks, leaves)
    assert np.array_equal(leaves.cpu().numpy().reshape(n_cols, 32), exp.hashes[:n_cols])
    with pytest.raises(P.LcpcError):  # the chunk range must start inside the row window
        ops.hash_chunk_range(hi, row_split, n_rows, n_cols, split_chunk - 1, n_chunks)


@pytest.mark.parametrize("n_chunks,n_cols,n_leaves", [
    (1, 300, 512), (2, 37, 64), (5, 1000, 1024), (8, 256, 256), (9, 700, 1024), (33, 520, 1024), (147, 100, 128),
    (200, 40, 64), (700, 9, 16),   # 8..159 chunks take the level-wise merge, the others the per-thread walk
])
def test_hash_merge_tree_one_launch(oracle, n_chunks, n_cols, n_leaves):
    """lcpc_dev_hash_merge_tree (leaf merge + the whole Merkle tree, one launch, last-CTA ticket for the top levels)
    against the BLAKE3 parent tree and merkle_tree of the oracle on random chaining values, padding leaves included."""
    import torch

    import lcpc_proof_of_storage_b200 as P
    from lcpc_proof_of_storage_b200 import _lib

    O = oracle
    rng = np.random.default_rng(n_chunks * 1000 + n_cols)
    cvs = rng.integers(0, 256, size=(n_chunks, n_cols, 32), dtype=np.uint8)
    leaves = np.zeros((n_leaves, 32), dtype=np.uint8)
    leaves[:n_cols] = O.hash_merge(cvs) if n_chunks > 1 else cvs[0]
    exp = O.merkle_tree(leaves)
    ctx = P.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    lib = _lib.load()
    d_cvs = torch.from_numpy(cvs.reshape(-1)).cuda()
    for rep in range(2):   # twice: the ticket counter must be left at zero
        tree = torch.full(((2 * n_leaves - 1) * 32,), 0xAB, dtype=torch.uint8, device="cuda")
        if n_chunks == 1:
            tree[:n_cols * 32] = d_cvs
        _lib.check(lib.lcpc_dev_hash_merge_tree(ctx.handle, d_cvs.data_ptr(), n_cols, n_chunks, tree.data_ptr(), n_leaves, 0))
        assert np.array_equal(tree.cpu().numpy().reshape(-1, 32), exp)


def _worker_rows(rank, world, port, q, cv_fused=False):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import lcpc_proof_of_storage_b200 as P
        from lcpc_proof_of_storage_b200.sharded import ShardedLigeroCommitter
        from oracle import lcpc_oracle as O

        ok = True
        for fid, n_rows, n_per_row, n_cols in [(0, 700, 2048, 4096), (3, 70, 512, 1024)]:
            L = O.LIMBS[fid]
            n = n_rows * n_per_row - 11
            coeffs = np.zeros((n_rows * n_per_row, L), dtype=np.uint64)
            coeffs[:n] = O.random_field_elements(fid, 5, n)
            ctx = P.Context(rank, stream=torch.cuda.current_stream().cuda_stream)
            enc = P.LigeroEncoding(fid, n_per_row, n_cols, ctx=ctx)
            sc = ShardedLigeroCommitter(enc, n_rows, None, hashing="rows", fused=cv_fused)
            assert sc.hashing == "rows" and not sc.fused and sc.cv_fused == cv_fused
            r0, cnt = sc.rows[rank]
            local = torch.from_numpy(coeffs.reshape(n_rows, n_per_row, L)[r0:r0 + cnt].copy().view(np.int64).reshape(-1)).cuda()
            for _ in range(3):  # repeated commits rotate the chaining-value stores of the fused exchange
                sc.commit(local)
            hashes = sc.gather_hashes()
            tensors = O.random_field_elements(fid, 7, 2 * n_rows).reshape(2, n_rows, L)
            t_dev = torch.from_numpy(tensors.view(np.int64).reshape(-1).copy()).cuda()
            folded = sc.fold(t_dev)
            folded_enc = sc.fold_encoded(t_dev)
            cols = [0, n_cols - 1, n_cols // 2, n_cols // 2 - 1]
            opened = sc.open_columns(cols)
            dev_open = sc.open_columns_dev(cols)   # the same openings left on the device (rank 0)
            # the end-to-end form: pinned host rows in, encoded rows out, PCIe / encode overlapped over row chunks
            h_in = local.cpu().pin_memory()
            h_out = torch.empty(cnt * n_cols * L, dtype=torch.int64).pin_memory()
            sc.commit_host(h_in, h_out, n_chunks=3)
            sc.wait_host_copies()
            torch.cuda.synchronize()
            exp = O.commit(coeffs[:n], O.LigeroEncoding(fid, n_per_row, n_cols))
            ok &= np.array_equal(h_out.numpy().view(np.uint64).reshape(cnt, n_cols, L), exp.comm[r0:r0 + cnt])
            if rank == 0:
                ok &= sc.root() == exp.get_root()
                ok &= np.array_equal(dev_open[0].cpu().numpy().view(np.uint64), np.stack([O.open_column(exp, c).col for c in cols]))
                ok &= np.array_equal(dev_open[1].cpu().numpy(), np.stack([O.open_column(exp, c).path for c in cols]))
                ok &= np.array_equal(hashes.cpu().numpy().reshape(-1, 32), exp.hashes)
                f = folded.cpu().numpy().view(np.uint64).reshape(2, n_per_row, L)
                fe = folded_enc.cpu().numpy().view(np.uint64).reshape(2, n_cols, L)
                for t in range(2):
                    ok &= np.array_equal(f[t], O.collapse_columns(fid, exp.coeffs, tensors[t]))
                    ok &= np.array_equal(fe[t], O.collapse_columns(fid, exp.comm, tensors[t]))
                for c, col in zip(cols, opened):
                    e = O.open_column(exp, c)
                    ok &= np.array_equal(col.col, e.col) and np.array_equal(col.path, e.path)
        if rank == 0:
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_row_hashed_commit_one_gpu():
    """The row-hashed committer on a one-rank NCCL group (what the driver's one-GPU test box can run): commit, commit_host,
    folds, openings on the host and on the device against the oracle."""
    import torch.multiprocessing as mp

    sk = socket.socket()
    sk.bind(("127.0.0.1", 0))
    port = sk.getsockname()[1]
    sk.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_worker_rows, args=(0, 1, port, q, False))
    p.start()
    p.join(300)
    assert p.exitcode == 0
    assert q.get(timeout=5) is True


@pytest.mark.parametrize("cv_fused", [False, True])
def test_row_hashed_commit_two_gpus(cv_fused):
    """cv_fused: the hash kernel stores the chaining values into the owners' stores over NVLink (k_hash_chunks_scatter,
    symmetric memory + one barrier) instead of the NCCL all-to-all."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_rows, args=(r, 2, port, q, cv_fused)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


@pytest.mark.parametrize("fid,n_rows,n_per_row,n_cols,n_peers", [(0, 512, 2048, 4096, 2), (0, 700, 64, 128, 4), (3, 70, 64, 128, 2),
                                                                 (1, 130, 64, 128, 8)])
def test_hash_chunk_range_scatter_one_gpu(oracle, fid, n_rows, n_per_row, n_cols, n_peers):
    """k_hash_chunks_scatter with every "peer" store on the same GPU: store g must hold the chaining values of column
    block g, [chunk][column in block], exactly as lcpc_dev_hash_chunk_range computes them."""
    import ctypes as C

    import torch

    import lcpc_proof_of_storage_b200 as P
    from lcpc_proof_of_storage_b200.sharded import GpuOps

    O = oracle
    ctx = P.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    enc = P.LigeroEncoding(fid, n_per_row, n_cols, ctx=ctx)
    ops = GpuOps(enc)
    coeffs = O.random_field_elements(fid, 23, n_rows * n_per_row)
    comm = ops.encode(torch.from_numpy(coeffs.view(np.int64).reshape(-1).copy()).cuda(), n_rows)
    n_chunks = O.leaf_chunks(fid, n_rows)
    cb = n_cols // n_peers
    ref = ops.hash_chunk_range(comm, 0, n_rows, n_cols, 0, n_chunks).view(n_chunks, n_cols, 32)
    stores = [torch.zeros(n_chunks * cb * 32, dtype=torch.uint8, device="cuda") for _ in range(n_peers)]
    ptrs = (C.c_void_p * n_peers)(*[s.data_ptr() for s in stores])
    ops.hash_chunk_range_scatter(comm, 0, n_rows, n_cols, 0, n_chunks, ptrs)
    for g in range(n_peers):
        assert torch.equal(stores[g].view(n_chunks, cb, 32), ref[:, g * cb:(g + 1) * cb]), g
    # and the leaves that come out of a store are the commitment's
    exp = O.commit(coeffs, O.LigeroEncoding(fid, n_per_row, n_cols))
    leaves = torch.zeros(cb * 32, dtype=torch.uint8, device="cuda")
    ops.hash_merge(stores[n_peers - 1], cb, n_chunks, leaves)
    assert np.array_equal(leaves.cpu().numpy().reshape(cb, 32), exp.hashes[(n_peers - 1) * cb:n_peers * cb])
