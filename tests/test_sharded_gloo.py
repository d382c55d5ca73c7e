"""The N > 1 path on CPU: world_size-2 (and 4) gloo runs of ShardedLigeroCommitter with the
numerical back end replaced by the oracle (injected by this test -- the product never imports
oracle/).  Checks the host logic: row partition, all-to-all packing/splits, subtree + top-of-tree
assembly, sharded fold and column opening, all bit-identical to the single-process commit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleOps:
    """Same interface as sharded.GpuOps, computed by the CPU oracle (test double)."""

    def __init__(self, O, fid, n_per_row, n_cols, oenc=None):
        self.O, self.fid, self.n_per_row, self.n_cols = O, fid, n_per_row, n_cols
        self.L = O.LIMBS[fid]
        self.oenc = oenc  # an oracle encoding with encode_rows (Brakedown); None = Ligero fft_io

    def _np(self, t):
        return t.numpy().view(np.uint64)

    def encode(self, coeffs, n_rows):
        rows = np.zeros((n_rows, self.n_cols, self.L), dtype=np.uint64)
        rows[:, :self.n_per_row] = self._np(coeffs).reshape(n_rows, self.n_per_row, self.L)
        if n_rows:
            out = self.oenc.encode_rows(rows) if self.oenc is not None else self.O.fft_io(self.fid, rows)
        else:
            out = rows
        return torch.from_numpy(out.view(np.int64).reshape(-1).copy())

    def hash_columns(self, mat, n_rows, row_stride, n_cols, out):
        m = self._np(mat).reshape(n_rows, row_stride, self.L)[:, :n_cols]
        out[:n_cols * 32] = torch.from_numpy(self.O.hash_columns(self.fid, np.ascontiguousarray(m)).reshape(-1))

    def merkle_tree(self, hashes, n_leaves):
        arr = hashes.numpy().reshape(-1, 32)
        hashes[:] = torch.from_numpy(self.O.merkle_tree(arr[:n_leaves].copy()).reshape(-1))

    def hash_chunk_range(self, mat, row_base, n_rows_total, n_cols, chunk0, chunk_end):
        n_local = mat.numel() // (n_cols * self.L)
        m = self._np(mat).reshape(n_local, n_cols, self.L)
        cvs = self.O.hash_chunk_cvs(self.fid, m, row_base, n_rows_total, chunk0, chunk_end)
        return torch.from_numpy(cvs.reshape(-1).copy())

    def hash_merge(self, cvs, n_cols, n_chunks, out):
        leaves = self.O.hash_merge(cvs.numpy().reshape(n_chunks, n_cols, 32))
        out[:n_cols * 32] = torch.from_numpy(leaves.reshape(-1))

    def fold(self, mat, n_rows, width, row_stride, tensors, n_tensors):
        m = self._np(mat).reshape(n_rows, row_stride, self.L)[:, :width]
        t = self._np(tensors).reshape(n_tensors, n_rows, self.L)
        out = np.stack([self.O.collapse_columns(self.fid, np.ascontiguousarray(m), t[i]) for i in range(n_tensors)])
        return torch.from_numpy(out.view(np.int64).reshape(-1).copy())

    def add_partials(self, parts, n_parts, n):
        p = self._np(parts).reshape(n_parts, n, self.L)
        acc = p[0].copy()
        for k in range(1, n_parts):
            acc = self.O.fe_add(self.fid, acc, p[k])
        return torch.from_numpy(acc.view(np.int64).reshape(-1).copy())


class _Enc:
    def __init__(self, fid, n_per_row, n_cols):
        self.fid, self.n_per_row, self.n_cols = fid, n_per_row, n_cols


def _worker(rank, world, port, fid, n_rows, n_per_row, n_cols, q, brakedown_seed=None, hashing="columns"):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import lcpc_oracle as O
        from lcpc_proof_of_storage_b200.sharded import ShardedLigeroCommitter, row_partition

        O.set_threads(1)
        L = O.LIMBS[fid]
        n = n_rows * n_per_row - 5  # ragged last row
        coeffs = np.zeros((n_rows * n_per_row, L), dtype=np.uint64)
        coeffs[:n] = O.random_field_elements(fid, 99, n)
        oenc = None
        if brakedown_seed is not None:  # Brakedown: n_cols is whatever the code generator gives (not a power of two)
            oenc = O.SdigEncoding(fid, n_per_row, brakedown_seed)
            n_cols = oenc.n_cols
        enc = _Enc(fid, n_per_row, n_cols)
        sc = ShardedLigeroCommitter(enc, n_rows, None, ops=OracleOps(O, fid, n_per_row, n_cols, oenc), hashing=hashing)
        r0, cnt = sc.rows[rank]
        if hashing == "columns":
            assert (r0, cnt) == row_partition(n_rows, world)[rank]
        local = coeffs.reshape(n_rows, n_per_row, L)[r0:r0 + cnt]
        sc.commit(torch.from_numpy(np.ascontiguousarray(local).view(np.int64).reshape(-1)))
        hashes = sc.gather_hashes()
        tensors = O.random_field_elements(fid, 7, 2 * n_rows).reshape(2, n_rows, L)
        folded = sc.fold(torch.from_numpy(tensors.view(np.int64).reshape(-1).copy()))
        folded_enc = sc.fold_encoded(torch.from_numpy(tensors.view(np.int64).reshape(-1).copy()))
        cols = [0, n_cols - 1, n_cols // 2, 3]
        opened = sc.open_columns(cols)
        if rank == 0:
            exp = O.commit(coeffs[:n], oenc if oenc is not None else O.LigeroEncoding(fid, n_per_row, n_cols))
            ok = sc.root() == exp.get_root()
            ok &= np.array_equal(hashes.numpy().reshape(-1, 32), exp.hashes)
            f = folded.numpy().view(np.uint64).reshape(2, n_per_row, L)
            fe = folded_enc.numpy().view(np.uint64).reshape(2, n_cols, L)
            for t in range(2):
                ok &= np.array_equal(f[t], O.collapse_columns(fid, exp.coeffs, tensors[t]))
                ok &= np.array_equal(fe[t], O.collapse_columns(fid, exp.comm, tensors[t]))  # lcpc_online.rs:454-484
            for c, col in zip(cols, opened):
                e = O.open_column(exp, c)
                ok &= np.array_equal(col.col, e.col) and np.array_equal(col.path, e.path)
                ok &= O.verify_column_path(fid, O.LcColumn(np.ascontiguousarray(col.col), np.ascontiguousarray(col.path)), c,
                                           exp.get_root())
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,fid,n_rows,n_per_row,n_cols", [(1, 0, 5, 16, 32), (2, 0, 7, 16, 32), (2, 3, 4, 8, 16), (4, 0, 10, 16, 64)])
def test_sharded_commit_matches_single_process(oracle, world, fid, n_rows, n_per_row, n_cols):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fid, n_rows, n_per_row, n_cols, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


@pytest.mark.parametrize("world,fid,n_rows,n_per_row,seed", [(1, 0, 5, 150, 0), (2, 0, 9, 150, 0), (4, 1, 6, 120, 1)])
def test_sharded_brakedown_commit_matches_single_process(oracle, world, fid, n_rows, n_per_row, seed):
    """Brakedown rows shard the same way; the PADDED leaf range (np2 > n_cols) is what gets split into column blocks,
    so the all-zero padding leaves fall into the last ranks' subtrees."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fid, n_rows, n_per_row, 0, q, seed)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


@pytest.mark.parametrize("world,fid,n_rows,n_per_row,n_cols", [
    (1, 0, 300, 8, 16),     # one rank: 3 chunks, no exchange
    (2, 0, 300, 8, 16),     # 124 + 128 + 48 rows: ranks own 1 and 2 chunks
    (4, 0, 700, 8, 16),     # 6 chunks over 4 ranks
    (4, 0, 130, 8, 16),     # 2 chunks over 4 ranks: two ranks own nothing
    (2, 3, 70, 8, 16),      # 32-byte elements: 31 rows in chunk 0, 32 per chunk after
    (2, 1, 130, 8, 16),     # 16-byte elements
])
def test_row_hashed_commit_matches_single_process(oracle, world, fid, n_rows, n_per_row, n_cols):
    """hashing="rows": chunk-aligned row blocks, chunk chaining values computed where the rows are, only those
    re-sharded (the lower-traffic exchange).  Root, tree, folds and openings equal the single-process commit."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fid, n_rows, n_per_row, n_cols, q, None, "rows")) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


@pytest.mark.parametrize("world,fid,n_rows,n_per_row,seed", [(2, 0, 300, 150, 0), (4, 1, 140, 120, 1)])
def test_row_hashed_brakedown_commit_matches_single_process(oracle, world, fid, n_rows, n_per_row, seed):
    """hashing="rows" with a non power-of-two n_cols: the chaining values of the real columns are re-sharded over the
    PADDED leaf range (short / empty last blocks), padding leaves stay zero."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fid, n_rows, n_per_row, 0, q, seed, "rows")) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_chunk_row_partition():
    from lcpc_proof_of_storage_b200.sharded import chunk_row_partition

    rows, chunks = chunk_row_partition(1, 4096, 8)  # the 8-GPU bench shape: 124 + 31 * 128 + 4 rows = 33 chunks
    assert chunks == [(0, 4), (4, 8), (8, 12), (12, 16), (16, 20), (20, 24), (24, 28), (28, 33)]
    assert rows == [(0, 508)] + [(508 + 512 * i, 512) for i in range(6)] + [(3580, 516)]
    rows, chunks = chunk_row_partition(4, 300, 4)   # 32-byte elements: 31 rows, then 32 per chunk
    assert sum(c for _, c in rows) == 300 and chunks[-1][1] == (32 + 300 * 32 + 1023) // 1024
    for (r0, cnt), (c0, c1) in zip(rows, chunks):
        assert r0 == (0 if c0 == 0 else 31 + 32 * (c0 - 1))
    assert chunk_row_partition(3, 100, 2) is None   # 24-byte elements straddle chunk boundaries
    assert chunk_row_partition(1, 100, 2) is None   # single-chunk leaf


def test_row_partition():
    from lcpc_proof_of_storage_b200.sharded import row_partition

    assert row_partition(7, 2) == [(0, 4), (4, 3)]
    assert row_partition(2, 4) == [(0, 1), (1, 1), (2, 0), (2, 0)]
    assert row_partition(512, 8) == [(64 * i, 64) for i in range(8)]
