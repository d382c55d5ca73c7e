"""One-process-per-GPU commitment: rows sharded for encoding, an all-to-all re-shards the
encoded matrix to column blocks for hashing, every rank builds its Merkle subtree and rank 0
joins the subtree roots (SURVEY.md section 8 e).

The reference has no multi-process path; what is sharded here is the loop structure of
lcpc-2d/src/lib.rs:651-700: `par_chunks_mut` over rows (:677-682) becomes the row shard,
`hash_columns`' column recursion (:736-775) becomes the column shard, and the top
log2(world) levels of `merkle_tree` (:777-815) are computed once on rank 0.  The result
(root, tree, fold vectors, opened columns) is bit-identical to the single-GPU commit.

A second hashing mode (hashing="rows", SURVEY.md section 8e's lower-traffic alternative) keeps the encoded matrix
row-sharded: BLAKE3 hashes the 1024-byte chunks of a leaf independently, so each rank computes the chunk chaining values
of ALL columns over its own (chunk-aligned) rows and only those 32-byte values are re-sharded to column blocks.

`torch.distributed` carries the plumbing (NCCL over NVLink on GPUs; gloo in the CPU tests,
where the numerical back end is injected by the test).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .lcpc2d import FIELD_LIMBS, LcColumn, LigeroEncoding, SdigEncoding, log2, next_pow2


def chunk_row_partition(limbs: int, n_rows: int, world: int):
    """Row blocks whose boundaries are BLAKE3 chunk boundaries of the column leaves (the
    lower-traffic hashing): the leaf stream of a column is 32 zero bytes followed by 8*limbs bytes per row, BLAKE3
    cuts it into 1024-byte chunks, so chunk 0 holds the first (1024 - 32) / w rows and every later chunk 1024 / w.
    Rank q owns chunks [c0, c1) and exactly their rows; the boundaries are the chunk boundaries nearest to the even
    split q * n_rows / world.  Returns (rows, chunks): [(row0, count)] and [(c0, c1)] per rank, or None when the
    elements straddle chunk boundaries (24-byte elements) or the leaf is a single chunk."""
    w = 8 * limbs
    if 1024 % w:
        return None
    n_chunks = (32 + n_rows * w + 1023) // 1024
    if n_chunks < 2:
        return None
    first, per = (1024 - 32) // w, 1024 // w

    def start(c: int) -> int:  # first row of chunk c; start(n_chunks) = n_rows
        return 0 if c == 0 else min(n_rows, first + (c - 1) * per)

    bounds = [0]
    for q in range(1, world):
        target = q * n_rows / world
        c = min(range(bounds[-1], n_chunks + 1), key=lambda k: (abs(start(k) - target), k))
        bounds.append(c)
    bounds.append(n_chunks)
    chunks = [(bounds[q], bounds[q + 1]) for q in range(world)]
    rows = [(start(c0), start(c1) - start(c0)) for c0, c1 in chunks]
    return rows, chunks


def row_partition(n_rows: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced row blocks: rank q owns rows [start, start + count)."""
    base, rem = divmod(n_rows, world)
    out, start = [], 0
    for q in range(world):
        cnt = base + (1 if q < rem else 0)
        out.append((start, cnt))
        start += cnt
    return out


class GpuOps:
    """Numerical back end on the local GPU: the device-pointer entry points of the C ABI."""

    def __init__(self, enc):
        self.enc = enc
        self.lib = _lib.load()
        self.fid = enc.fid
        self.L = FIELD_LIMBS[enc.fid]
        self.device = torch.device("cuda", enc.ctx.device)
        # The lcpc_dev_* calls enqueue on the context's stream, the torch work around them (allocations, packs, NCCL
        # collectives, symmetric-memory barriers) on torch's current stream, and nothing orders two different streams:
        # the pack could read rows the encode has not written yet.  The context must have been created on the stream
        # torch is using -- Context(device, stream=torch.cuda.current_stream().cuda_stream).
        cur = torch.cuda.current_stream(self.device).cuda_stream
        if enc.ctx.stream != cur:
            raise ValueError("sharded commit: the encoding's Context enqueues on stream %#x but torch's current stream on %s "
                             "is %#x; create it with Context(device, stream=torch.cuda.current_stream().cuda_stream)"
                             % (enc.ctx.stream, self.device, cur))

    def encode(self, coeffs: torch.Tensor, n_rows: int) -> torch.Tensor:
        comm = torch.empty(n_rows * self.enc.n_cols * self.L, dtype=torch.int64, device=self.device)
        if n_rows:
            _lib.check(self.lib.lcpc_dev_encode(self.enc.plan, coeffs.data_ptr(), n_rows, comm.data_ptr()))
        return comm

    def encode_into(self, coeffs: torch.Tensor, n_rows: int, comm: torch.Tensor) -> None:
        """Encode n_rows rows of `coeffs` into the preallocated `comm` (both may be slices of larger buffers)."""
        if n_rows:
            _lib.check(self.lib.lcpc_dev_encode(self.enc.plan, coeffs.data_ptr(), n_rows, comm.data_ptr()))

    def encode_scatter(self, coeffs: torch.Tensor, n_rows: int, row0: int, scratch: torch.Tensor, peer_ptrs) -> None:
        """Fused encode + re-shard: the transform's last pass stores each row block straight into the
        owning rank's column-block matrix (peer HBM over NVLink), see lcpc_dev_encode_scatter."""
        if n_rows:
            _lib.check(self.lib.lcpc_dev_encode_scatter(self.enc.plan, coeffs.data_ptr(), n_rows, row0,
                                                        scratch.data_ptr(), peer_ptrs, len(peer_ptrs)))

    def pack_bytes7(self, data: torch.Tensor, n_elems_padded: int) -> torch.Tensor:
        """WriteableFt63::from_data_bytes over this rank's slice of the file; the tail is zero elements."""
        out = torch.zeros(n_elems_padded, dtype=torch.int64, device=self.device)
        if data.numel():
            _lib.check(self.lib.lcpc_dev_pack_bytes7(self.enc.ctx.handle, data.data_ptr(), data.numel(), out.data_ptr()))
        return out

    def hash_columns(self, mat: torch.Tensor, n_rows: int, row_stride: int, n_cols: int, out: torch.Tensor) -> None:
        _lib.check(self.lib.lcpc_dev_hash_columns(self.enc.ctx.handle, self.fid, mat.data_ptr(), n_rows, row_stride,
                                                  n_cols, out.data_ptr()))

    def merkle_tree(self, hashes: torch.Tensor, n_leaves: int) -> None:
        _lib.check(self.lib.lcpc_dev_merkle_tree(self.enc.ctx.handle, hashes.data_ptr(), n_leaves))

    def hash_chunk_range(self, mat: torch.Tensor, row_base: int, n_rows_total: int, n_cols: int, chunk0: int,
                         chunk_end: int) -> torch.Tensor:
        """BLAKE3 chaining values of chunks [chunk0, chunk_end) of every column's leaf, from this rank's encoded rows:
        [chunk_end - chunk0, n_cols, 32] bytes."""
        out = torch.empty((chunk_end - chunk0) * n_cols * 32, dtype=torch.uint8, device=self.device)
        if chunk_end > chunk0:
            _lib.check(self.lib.lcpc_dev_hash_chunk_range(self.enc.ctx.handle, self.fid, mat.data_ptr(), row_base, n_rows_total,
                                                          n_cols, n_cols, chunk0, chunk_end, out.data_ptr()))
        return out

    def hash_chunk_range_scatter(self, mat: torch.Tensor, row_base: int, n_rows_total: int, n_cols: int, chunk0: int,
                                 chunk_end: int, peer_ptrs) -> None:
        """Same values, stored by the kernel into the owning ranks' chaining-value stores (peer HBM over NVLink)."""
        if chunk_end > chunk0:
            _lib.check(self.lib.lcpc_dev_hash_chunk_range_scatter(self.enc.ctx.handle, self.fid, mat.data_ptr(), row_base,
                                                                  n_rows_total, n_cols, n_cols, chunk0, chunk_end, peer_ptrs,
                                                                  len(peer_ptrs)))

    def hash_merge(self, cvs: torch.Tensor, n_cols: int, n_chunks: int, out: torch.Tensor) -> None:
        _lib.check(self.lib.lcpc_dev_hash_merge(self.enc.ctx.handle, cvs.data_ptr(), n_cols, n_chunks, out.data_ptr()))

    def hash_merge_tree(self, cvs: torch.Tensor, n_cols: int, n_chunks: int, tree: torch.Tensor, n_leaves: int,
                        cv_stride: int = 0) -> None:
        """hash_merge + merkle_tree in one launch (padding leaves written as zero by the kernel); cv_stride = columns per
        chunk row of `cvs` when that is wider than n_cols (a scatter store of a block with padding columns)."""
        _lib.check(self.lib.lcpc_dev_hash_merge_tree(self.enc.ctx.handle, cvs.data_ptr(), n_cols, n_chunks, tree.data_ptr(),
                                                     n_leaves, cv_stride))

    def merkleize(self, mat: torch.Tensor, n_rows: int, row_stride: int, n_cols: int, tree: torch.Tensor) -> None:
        """hash_columns + merkle_tree over next_pow2(n_cols) leaves in one launch."""
        _lib.check(self.lib.lcpc_dev_merkleize(self.enc.ctx.handle, self.fid, mat.data_ptr(), n_rows, row_stride, n_cols,
                                               tree.data_ptr()))

    def fold(self, mat: torch.Tensor, n_rows: int, width: int, row_stride: int, tensors: torch.Tensor,
             n_tensors: int) -> torch.Tensor:
        out = torch.zeros(n_tensors * width * self.L, dtype=torch.int64, device=self.device)
        if n_rows:
            _lib.check(self.lib.lcpc_dev_fold(self.enc.ctx.handle, self.fid, mat.data_ptr(), n_rows, width, row_stride,
                                              tensors.data_ptr(), n_tensors, out.data_ptr()))
        return out

    def gather_columns(self, mat: torch.Tensor, n_rows: int, row_stride: int, d_cols: torch.Tensor) -> torch.Tensor:
        """out[i][r] = mat[r][cols[i]]: [n, n_rows * L] int64 on the device (k_gather_columns)."""
        n = d_cols.numel()
        out = torch.empty(n, n_rows * self.L, dtype=torch.int64, device=self.device)
        if n and n_rows:
            _lib.check(self.lib.lcpc_dev_gather_columns(self.enc.ctx.handle, self.fid, mat.data_ptr(), n_rows, row_stride,
                                                        d_cols.data_ptr(), n, out.data_ptr()))
        return out

    def gather_paths(self, tree: torch.Tensor, n_leaves: int, d_cols: torch.Tensor) -> torch.Tensor:
        """Sibling digests of every index inside a flat tree over n_leaves: [n, log2(n_leaves), 32] uint8 (k_gather_paths)."""
        n, depth = d_cols.numel(), log2(n_leaves)
        out = torch.empty(n, depth, 32, dtype=torch.uint8, device=self.device)
        if n and depth:
            _lib.check(self.lib.lcpc_dev_gather_paths(self.enc.ctx.handle, tree.data_ptr(), n_leaves, d_cols.data_ptr(), n,
                                                      out.data_ptr()))
        return out

    def add_partials(self, parts: torch.Tensor, n_parts: int, n: int) -> torch.Tensor:
        out = torch.empty(n * self.L, dtype=torch.int64, device=self.device)
        _lib.check(self.lib.lcpc_dev_add_partials(self.enc.ctx.handle, self.fid, parts.data_ptr(), n_parts, n,
                                                  out.data_ptr()))
        return out


class ShardedLigeroCommitter:
    """Commit one (n_rows_total x n_per_row) coefficient matrix across the ranks of `group`.

    Rank q passes its own row block (row_partition) to `commit`.  After it, every rank holds
    its column block of the encoded matrix and the matching Merkle subtree; rank 0 also holds
    the top of the tree and the root.
    """

    def __init__(self, enc, n_rows_total: int, group=None, ops=None, fused: Optional[bool] = None, hashing: str = "columns"):
        """hashing="auto": row hashing with the chaining values stored into the owners' stores over NVLink when the shape
        allows it (GPU back end, 2..16 ranks, power-of-two n_cols, leaves of at least two chunks, elements that divide a
        chunk) -- 3 % of the exchange volume of column blocks and the measured winner from 2 GPUs up
        (profiles/r02_scaling.md) -- else column blocks.
        hashing="columns": the encoded matrix is re-sharded to column blocks and every rank hashes whole
        columns.  hashing="rows": every rank hashes the BLAKE3 chunks of ALL columns that its own rows make up
        (chunk-aligned row blocks, chunk_row_partition) and only the 32-byte chunk chaining values are re-sharded to
        column blocks -- 3 % of the volume for 8-byte elements; the encoded matrix stays row-sharded."""
        self.enc = enc
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.ops = ops or GpuOps(enc)
        self.L = FIELD_LIMBS[enc.fid]
        self.n_rows = n_rows_total
        self.n_per_row, self.n_cols = enc.n_per_row, enc.n_cols
        self.np2 = next_pow2(self.n_cols)
        if self.world & (self.world - 1) or self.world > self.np2:
            raise ValueError("world size must be a power of two not exceeding the padded column count")
        # the PADDED leaf range is sharded, so padding (all-zero) leaves fall in the last shards
        self.cb = self.np2 // self.world
        self.rows = row_partition(n_rows_total, self.world)
        self.hashing, self.chunks = "columns", None
        if hashing == "auto":
            part = chunk_row_partition(self.L, n_rows_total, self.world)
            # row hashing cuts the rows on chunk boundaries: a leaf of few chunks (a short matrix: Brakedown's 13 - 404
            # rows) leaves some ranks with far more rows than others, or none -- 202 rows of 8 bytes are two chunks, i.e.
            # two busy ranks at any world size (profiles/r02_sweep.md).  Take it when the largest block is within 30 % of
            # the even split, the even row split with column blocks otherwise.
            if part is None or max(cnt for _, cnt in part[0]) > 1.3 * n_rows_total / self.world:
                hashing, fused = "columns", None
            else:  # stores over NVLink where peers can be mapped, the NCCL all-to-all of the chaining values otherwise
                hashing, fused = "rows", isinstance(self.ops, GpuOps) and 1 < self.world <= 16
        if hashing == "rows":
            part = chunk_row_partition(self.L, n_rows_total, self.world)
            if part is None:
                raise ValueError("hashing='rows' needs elements that divide a 1024-byte chunk and leaves of at least two chunks")
            self.rows, self.chunks = part
            self.hashing = "rows"
            self.n_chunks = self.chunks[-1][1]
            # fused=True: k_hash_chunks_scatter stores the chaining values straight into the owners' stores (symmetric
            # memory, below) instead of the NCCL all-to-all.  Opt-in until it has been timed at 4 and 8 GPUs.
            self._cv_fused = bool(fused)
            fused = False
        elif hashing != "columns":
            raise ValueError("hashing must be 'columns' or 'rows'")
        self.comm_rows: Optional[torch.Tensor] = None   # hashing="rows": [rows_local, n_cols, L], my rows, all columns
        self._cv_hdl = self._cv_peer_ptrs = self._cv_bufs = None
        self.row0, self.rows_local = self.rows[self.rank]
        self.col0 = self.rank * self.cb
        self.cols_local = max(0, min(self.n_cols, self.col0 + self.cb) - self.col0)  # real columns in my block
        self.coeffs_local: Optional[torch.Tensor] = None
        self.comm_cols: Optional[torch.Tensor] = None   # [n_rows_total, col_stride, L]: my column block, all rows
        self.col_stride = self.cb                        # row stride of comm_cols (n_cols on the single-rank path)
        self.subtree: Optional[torch.Tensor] = None     # [(2*cb-1)*32] uint8
        self.top: Optional[torch.Tensor] = None         # rank 0: [(2*world-1)*32] uint8
        # Fused path (GPU back end, world > 1, power-of-two n_cols): the column-block matrix lives in
        # symmetric memory, every rank's encode kernel writes its rows into all peers' matrices over
        # NVLink, and no separate pack / all-to-all pass exists.  fused=None tries it and falls back
        # to NCCL all_to_all_single when peer mapping is not available.
        self._symm = self._hdl = self._peer_ptrs = self._scratch = None
        self._pending: Optional[int] = None  # buffer index of a commit whose exchange is issued but not yet hashed
        want = fused if fused is not None else (isinstance(self.ops, GpuOps) and self.world > 1)
        # the transform's last pass stores whole shared-memory blocks (2^12 / 2^11 / 2^10 elements for 1 / 2 / 3-4 limbs): a
        # column block must hold at least one
        # (Brakedown plans have no such constraint and no power-of-two n_cols: their transposing passes store element by
        # element into the owners' blocks of the PADDED column range, lcpc_dev_encode_scatter)
        sdig = isinstance(enc, SdigEncoding)
        ntt_block = 1 << min(log2(self.n_cols), {1: 12, 2: 11}.get(self.L, 10))
        if fused is None and not sdig and self.cb < ntt_block:
            want = False
        if want and isinstance(self.ops, GpuOps) and self.world > 1 and (sdig or self.np2 == self.n_cols) and self.world <= 16:
            try:
                import ctypes as C

                import torch.distributed._symmetric_memory as symm_mem

                dev = torch.device("cuda", enc.ctx.device)
                # Three matrices used in turn.  Commit k is finished (barrier, hash, tree) either right after its own
                # encode or, with defer=True, after the encode of commit k+1 has been issued -- the exchange of commit k
                # then drains over NVLink, and slower ranks catch up, behind useful work.  Encode j writes buffer j % 3,
                # last read by the hashing of commit j-3; encode j is issued after this rank's barrier of commit j-2,
                # which every rank reaches only after hashing commit j-3: no extra synchronisation is needed.
                per = self.n_rows * self.cb * self.L
                self._symm = symm_mem.empty(3 * per, dtype=torch.int64, device=dev)
                self._hdl = symm_mem.rendezvous(self._symm, group if group is not None else dist.group.WORLD)
                self._peer_ptrs = [(C.c_void_p * self.world)(*[int(p) + b * per * 8 for p in self._hdl.buffer_ptrs])
                                   for b in range(3)]
                self._bufs = [self._symm[b * per:(b + 1) * per] for b in range(3)]
                self._k = 0
                self._scratch = torch.empty(1 if sdig else max(1, self.rows_local) * self.n_cols * self.L, dtype=torch.int64,
                                            device=dev)
            except Exception:
                if fused:
                    raise
                self._symm = self._hdl = self._peer_ptrs = self._scratch = None

        if self.hashing == "rows" and self._cv_fused:
            if not (isinstance(self.ops, GpuOps) and 1 < self.world <= 16):
                raise ValueError("fused chaining-value exchange needs the GPU back end and 2..16 ranks")
            import ctypes as C

            import torch.distributed._symmetric_memory as symm_mem

            dev = torch.device("cuda", enc.ctx.device)
            # Two chaining-value stores used alternately ([n_chunks][cb][32 B] each): commit k+2 writes the store commit k
            # was merged from, and a rank can only issue that write after its barrier of commit k+1, which every rank
            # reaches after its own merge of commit k (stream order): one barrier per commit is enough.
            per = self.n_chunks * self.cb * 32
            self._cv_symm = symm_mem.empty(2 * per, dtype=torch.uint8, device=dev)
            self._cv_hdl = symm_mem.rendezvous(self._cv_symm, group if group is not None else dist.group.WORLD)
            self._cv_peer_ptrs = [(C.c_void_p * self.world)(*[int(p) + b * per for p in self._cv_hdl.buffer_ptrs])
                                  for b in range(2)]
            self._cv_bufs = [self._cv_symm[b * per:(b + 1) * per] for b in range(2)]
            self._cv_k = 0

    @property
    def fused(self) -> bool:
        return self._hdl is not None

    @property
    def cv_fused(self) -> bool:
        """hashing="rows" with the chaining-value exchange done by the hash kernel's own peer stores."""
        return self._cv_hdl is not None

    # ------------------------------------------------------------------ commit
    def commit(self, coeffs_local: torch.Tensor, defer: bool = False) -> None:
        """defer=True (fused path only): the encode + exchange of this commit is issued, its hashing and tree follow when
        the next commit has issued its encode, or at flush().  Every rank must make the same sequence of calls."""
        L, cb, W = self.L, self.cb, self.world
        assert coeffs_local.numel() == self.rows_local * self.n_per_row * L
        dev = coeffs_local.device
        self.coeffs_local = coeffs_local
        if self.fused:
            b = self._k % 3
            self._k += 1
            self.ops.encode_scatter(coeffs_local, self.rows_local, self.row0, self._scratch, self._peer_ptrs[b])
            self.flush()          # the previous commit, if it was deferred
            self._pending = b
            if not defer:
                self.flush()
            return
        comm = self.ops.encode(coeffs_local, self.rows_local)  # [rows_local, n_cols, L]
        if self.hashing == "rows":
            self._commit_row_hashed(comm, dev)
            return
        if W == 1:
            # a single rank owns every column: hash the encoded matrix where it is (row stride n_cols, no padding copy)
            self.comm_cols = comm
            self.col_stride = self.n_cols
            self._finish_tree(dev)
            return
        # pack: one contiguous slab per destination rank = that rank's column block of my rows
        if self.np2 != self.n_cols:
            padded = torch.zeros(self.rows_local, self.np2, L, dtype=torch.int64, device=dev)
            padded[:, :self.n_cols] = comm.view(self.rows_local, self.n_cols, L)
            comm3 = padded
        else:
            comm3 = comm.view(self.rows_local, self.n_cols, L)
        send = comm3.view(self.rows_local, W, cb * L).transpose(0, 1).contiguous()  # [W, rows_local, cb*L]
        recv = torch.empty(self.n_rows * cb * L, dtype=torch.int64, device=dev)
        in_splits = [self.rows_local * cb * L] * W
        out_splits = [cnt * cb * L for (_, cnt) in self.rows]
        if W > 1:
            dist.all_to_all_single(recv, send.view(-1), out_splits, in_splits, group=self.group)
        else:
            recv = send.view(-1)
        self.comm_cols = recv  # row-major [n_rows_total, cb, L] because row blocks arrive in rank order
        self._finish_tree(dev)

    def commit_host(self, h_coeffs: torch.Tensor, h_comm: Optional[torch.Tensor] = None, n_chunks: int = 8) -> None:
        """End-to-end form of commit() for row hashing: this rank's coefficient rows come from PINNED host memory and, when
        `h_comm` is given, its encoded rows go back to pinned host memory, with the three stages overlapped over row
        chunks the way lcpc_commit_host does it on one GPU: chunk k+1 crosses PCIe while chunk k is encoded and chunk k-1
        leaves (copy streams on both sides of torch's current stream).  Column hashing needs every row and follows the
        last chunk.  Bit-identical to commit(): rows are independent."""
        assert self.hashing == "rows" and isinstance(self.ops, GpuOps), "commit_host: row hashing on the GPU back end"
        L, npr, nc = self.L, self.n_per_row, self.n_cols
        dev = self.ops.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_s_in", None) is None:
            self._s_in, self._s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self._h_coeffs_dev = torch.empty(max(1, self.rows_local) * npr * L, dtype=torch.int64, device=dev)
            self._h_comm_dev = torch.empty(max(1, self.rows_local) * nc * L, dtype=torch.int64, device=dev)
        d_coeffs, comm = self._h_coeffs_dev, self._h_comm_dev
        # the previous call's copies out of `comm` and its kernels reading `d_coeffs` are done before either is rewritten
        main.wait_stream(self._s_out)
        self._s_in.wait_stream(main)
        rows = self.rows_local
        k = max(1, min(n_chunks, rows))
        per = (rows + k - 1) // k
        for r0 in range(0, rows, per):
            nr = min(per, rows - r0)
            a, b = r0 * npr * L, (r0 + nr) * npr * L
            with torch.cuda.stream(self._s_in):
                d_coeffs[a:b].copy_(h_coeffs[a:b], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self._s_in)
            main.wait_event(ev_in)
            ca, cb_ = r0 * nc * L, (r0 + nr) * nc * L
            self.ops.encode_into(d_coeffs[a:b], nr, comm[ca:cb_])
            if h_comm is not None:
                ev_enc = torch.cuda.Event()
                ev_enc.record(main)
                self._s_out.wait_event(ev_enc)
                with torch.cuda.stream(self._s_out):
                    h_comm[ca:cb_].copy_(comm[ca:cb_], non_blocking=True)
        self.coeffs_local = d_coeffs
        self._commit_row_hashed(comm, dev)

    def wait_host_copies(self) -> None:
        """Joins commit_host's device -> host copies onto torch's current stream."""
        if getattr(self, "_s_out", None) is not None:
            torch.cuda.current_stream(self.ops.device).wait_stream(self._s_out)

    def _commit_row_hashed(self, comm: torch.Tensor, dev) -> None:
        """Chunk chaining values of all columns from my rows, an all-to-all of those (32 bytes per chunk and column),
        then the BLAKE3 parent tree per column of my column block, my Merkle subtree and the join of the roots."""
        cb, W = self.cb, self.world
        self.comm_rows = comm
        c0, c1 = self.chunks[self.rank]
        if self.cv_fused:
            b = self._cv_k % 2
            self._cv_k += 1
            self.ops.hash_chunk_range_scatter(comm, self.row0, self.n_rows, self.n_cols, c0, c1, self._cv_peer_ptrs[b])
            self._cv_hdl.barrier(channel=b)  # every rank's chaining values for my column block have landed
            recv = self._cv_bufs[b]
        else:
            recv = self._exchange_chunk_values(comm, c0, c1, dev)
        # chunk order = rank order, so recv is [n_chunks, cols_local, 32]: the chaining-value store of my column block
        if self.subtree is None or self.subtree.device != dev:
            self.subtree = torch.zeros((2 * cb - 1) * 32, dtype=torch.uint8, device=dev)
            self._roots = torch.empty(W * 32, dtype=torch.uint8, device=dev)
            self.top = torch.zeros((2 * W - 1) * 32, dtype=torch.uint8, device=dev)
        if self.cols_local and hasattr(self.ops, "hash_merge_tree"):
            # leaves of my column block from the chaining values + my whole subtree: one launch (a scatter store is cb wide
            # also when the block holds fewer real columns, the NCCL form is compact)
            self.ops.hash_merge_tree(recv, self.cols_local, self.n_chunks, self.subtree, cb, cb if self.cv_fused else 0)
            self._join_subtrees(tree_done=True)
            return
        if self.cols_local:  # padding leaves are never written: they stay zero
            self.ops.hash_merge(recv, self.cols_local, self.n_chunks, self.subtree)
        self._join_subtrees()

    def _exchange_chunk_values(self, comm: torch.Tensor, c0: int, c1: int, dev) -> torch.Tensor:
        """NCCL form of the exchange: my chunks' chaining values of every column, all-to-all to the column owners."""
        cb, W = self.cb, self.world
        cvs = self.ops.hash_chunk_range(comm, self.row0, self.n_rows, self.n_cols, c0, c1)  # [c1 - c0, n_cols, 32]
        if W > 1:
            # one slab per destination = my chunks of that rank's REAL columns (the padded leaf range is what is split
            # into column blocks, so with a non power-of-two n_cols the last blocks are short or empty)
            width = [max(0, min(self.n_cols, (r + 1) * cb) - r * cb) for r in range(W)]
            if self.np2 == self.n_cols:  # equal blocks: one strided copy
                send = cvs.view(c1 - c0, W, cb * 32).transpose(0, 1).contiguous().view(-1)
            else:
                cvs3 = cvs.view(c1 - c0, self.n_cols, 32)
                send = torch.cat([cvs3[:, r * cb:r * cb + width[r]].reshape(-1) for r in range(W)])
            recv = torch.empty(self.n_chunks * self.cols_local * 32, dtype=torch.uint8, device=dev)
            in_splits = [(c1 - c0) * width[r] * 32 for r in range(W)]
            out_splits = [(b - a) * self.cols_local * 32 for a, b in self.chunks]
            dist.all_to_all_single(recv, send, out_splits, in_splits, group=self.group)
        else:
            recv = cvs
        return recv

    def flush(self) -> None:
        """Finish the deferred commit, if any (collective: every rank calls it at the same point)."""
        if self._pending is None:
            return
        b, self._pending = self._pending, None
        self._hdl.barrier(channel=b)  # every rank's rows of that commit have landed in my column block
        self.comm_cols = self._bufs[b]
        self._finish_tree(self.comm_cols.device)

    def _finish_tree(self, dev) -> None:
        cb, W = self.cb, self.world
        recv = self.comm_cols
        # leaves of my column block + my subtree (padding leaves are never written: they stay zero)
        if self.subtree is None or self.subtree.device != dev:
            self.subtree = torch.zeros((2 * cb - 1) * 32, dtype=torch.uint8, device=dev)
            self._roots = torch.empty(W * 32, dtype=torch.uint8, device=dev)
            self.top = torch.zeros((2 * W - 1) * 32, dtype=torch.uint8, device=dev)
        if self.cols_local == cb and hasattr(self.ops, "merkleize"):
            self.ops.merkleize(recv, self.n_rows, self.col_stride, cb, self.subtree)  # leaves + subtree: one launch
            self._join_subtrees(tree_done=True)
            return
        if self.cols_local:
            self.ops.hash_columns(recv, self.n_rows, self.col_stride, self.cols_local, self.subtree)
        self._join_subtrees()

    def _join_subtrees(self, tree_done: bool = False) -> None:
        cb, W = self.cb, self.world
        if not tree_done:
            self.ops.merkle_tree(self.subtree, cb)
        # only the subtree roots travel (32 bytes per rank); rank 0 joins them
        my_root = self.subtree[-32:]
        if W > 1:
            dist.all_gather_into_tensor(self._roots, my_root, group=self.group)
            if self.rank == 0:
                self.top[:W * 32] = self._roots
                self.ops.merkle_tree(self.top, W)
        else:
            self.top[:32] = my_root

    # ------------------------------------------------------------------ proof-of-storage bytes
    def byte_range(self, n_bytes_total: int, rank: Optional[int] = None) -> Tuple[int, int]:
        """[start, end) of the file bytes that make up this rank's rows (7 bytes per element)."""
        r0, cnt = self.rows[self.rank if rank is None else rank]
        lo = min(n_bytes_total, r0 * self.n_per_row * 7)
        hi = min(n_bytes_total, (r0 + cnt) * self.n_per_row * 7)
        return lo, hi

    def commit_bytes(self, local_bytes: torch.Tensor) -> None:
        """proof-of-storage commit of a file sharded by row blocks: 7-byte packing on the device, then commit."""
        assert self.L == 1 and local_bytes.dtype == torch.uint8
        self.commit(self.ops.pack_bytes7(local_bytes, self.rows_local * self.n_per_row))

    def root(self) -> bytes:
        """LcCommit::get_root on rank 0."""
        assert self.rank == 0 and self.top is not None
        assert self._pending is None, "a deferred commit is pending: call flush() on every rank first"
        return bytes(self.top[-32:].cpu().numpy())

    def gather_hashes(self) -> Optional[torch.Tensor]:
        """The full flat tree `LcCommit.hashes` ([np2 | np2/2 | ... | 1] digests) on rank 0."""
        W, cb = self.world, self.cb
        assert self._pending is None, "a deferred commit is pending: call flush() on every rank first"
        dev = self.subtree.device
        if W == 1:
            return self.subtree.clone()
        parts = [torch.empty_like(self.subtree) for _ in range(W)] if self.rank == 0 else None
        dist.gather(self.subtree, parts, dst=dist.get_global_rank(self.group, 0) if self.group else 0, group=self.group)
        if self.rank != 0:
            return None
        out = torch.empty((2 * self.np2 - 1) * 32, dtype=torch.uint8, device=dev)
        pos, sub_off, n = 0, 0, cb
        while n >= 1:  # levels that live inside the subtrees
            for q in range(W):
                out[pos:pos + n * 32] = parts[q][sub_off:sub_off + n * 32]
                pos += n * 32
            sub_off += n * 32
            n //= 2
        out[pos:] = self.top[W * 32:]  # levels above the subtree roots
        return out

    # ------------------------------------------------------------------ fold (prove)
    def fold(self, tensors: torch.Tensor) -> torch.Tensor:
        """collapse_columns over the unencoded coefficients for a batch of tensors
        ([n_tensors, n_rows_total, L] on every rank): partial sums over the local rows, an
        all-gather of the partials, and a local modular sum (NCCL has no mod-p reduction)."""
        L = self.L
        n_t = tensors.numel() // (self.n_rows * L)
        t3 = tensors.view(n_t, self.n_rows, L)
        local_t = t3[:, self.row0:self.row0 + self.rows_local].contiguous()
        part = self.ops.fold(self.coeffs_local, self.rows_local, self.n_per_row, self.n_per_row, local_t, n_t)
        if self.world == 1:
            return part
        allp = torch.empty(self.world * part.numel(), dtype=torch.int64, device=part.device)
        dist.all_gather_into_tensor(allp, part, group=self.group)
        return self.ops.add_partials(allp, self.world, n_t * self.n_per_row)

    def fold_encoded(self, tensors: torch.Tensor) -> torch.Tensor:
        """The proof-of-storage fold over the ENCODED matrix (verifiable_polynomial_evaluation,
        proof-of-storage/src/lcpc_online.rs:454-484): out[t][j] = sum_r tensors[t][r] * comm[r][j] for all n_cols columns,
        [n_tensors, n_cols, L] on every rank.  With column blocks each rank folds its own columns over all rows and the
        blocks are concatenated; with row blocks (hashing="rows") it is the coefficient fold's scheme: partial sums over
        the local rows, all-gather, modular sum."""
        L, W, cb = self.L, self.world, self.cb
        assert self._pending is None, "a deferred commit is pending: call flush() on every rank first"
        n_t = tensors.numel() // (self.n_rows * L)
        t3 = tensors.view(n_t, self.n_rows, L)
        if self.hashing == "rows":
            local_t = t3[:, self.row0:self.row0 + self.rows_local].contiguous()
            part = self.ops.fold(self.comm_rows, self.rows_local, self.n_cols, self.n_cols, local_t, n_t)
            if W == 1:
                return part
            allp = torch.empty(W * part.numel(), dtype=torch.int64, device=part.device)
            dist.all_gather_into_tensor(allp, part, group=self.group)
            return self.ops.add_partials(allp, W, n_t * self.n_cols)
        dev = self.comm_cols.device
        if W == 1:
            return self.ops.fold(self.comm_cols, self.n_rows, self.n_cols, self.col_stride, t3.contiguous(), n_t)
        block = torch.zeros(n_t, cb, L, dtype=torch.int64, device=dev)  # my column block, padded to the block width
        if self.cols_local:
            mine = self.ops.fold(self.comm_cols, self.n_rows, self.cols_local, self.col_stride, t3.contiguous(), n_t)
            block[:, :self.cols_local] = mine.view(n_t, self.cols_local, L)
        allb = torch.empty(W * block.numel(), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allb, block.view(-1), group=self.group)
        out = allb.view(W, n_t, cb, L).permute(1, 0, 2, 3).reshape(n_t, W * cb, L)[:, :self.n_cols]
        return out.contiguous().view(-1)

    # ------------------------------------------------------------------ open
    def open_columns_dev(self, cols: Sequence[int]):
        """open_column for each index with everything on the device (row-sharded matrix, GPU back end): every rank
        gathers its rows of the requested columns (k_gather_columns) and the path levels inside its own subtree
        (k_gather_paths), ONE all-gather moves both, and rank 0 assembles [n, n_rows, L] column values and
        [n, depth, 32] paths with the siblings above the subtree roots appended.  Returns the two device tensors on rank
        0, None elsewhere.  The retrievability proof of proof-of-storage (networking/client.rs:443-456,
        lcpc_online.rs:226-237) is this call on 309 columns."""
        L, cb, W = self.L, self.cb, self.world
        assert self.hashing == "rows" and self.comm_rows is not None
        n = len(cols)
        dev = self.comm_rows.device
        idx = torch.tensor(list(cols), dtype=torch.int64, device=dev)
        max_rows = max(cnt for _, cnt in self.rows)
        depth = log2(self.n_cols)
        depth_sub = min(depth, log2(cb))
        vw = max_rows * L
        payload = torch.zeros(n, vw + depth_sub * 4, dtype=torch.int64, device=dev)
        if self.rows_local:
            payload[:, :self.rows_local * L] = self.ops.gather_columns(self.comm_rows, self.rows_local, self.n_cols, idx)
        if depth_sub:
            # every rank gathers from its own subtree at the column's offset inside a block; rank 0 keeps the owner's
            local = idx & (cb - 1)
            payload[:, vw:] = self.ops.gather_paths(self.subtree, cb, local).view(n, depth_sub * 32).view(torch.int64)
        if W > 1:
            allp = torch.empty(W, n, vw + depth_sub * 4, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allp.view(-1), payload.view(-1), group=self.group)
        else:
            allp = payload.view(1, n, -1)
        if self.rank != 0:
            return None
        vals = torch.cat([allp[q, :, :cnt * L] for q, (_, cnt) in enumerate(self.rows)], dim=1).view(n, self.n_rows, L)
        owner = idx // cb
        ar = torch.arange(n, device=dev)
        parts = []
        if depth_sub:
            parts.append(allp[owner, ar, vw:].contiguous().view(torch.uint8).view(n, depth_sub, 32))
        if depth > depth_sub:
            parts.append(self.ops.gather_paths(self.top, W, owner))
        paths = torch.cat(parts, dim=1) if parts else torch.empty(n, 0, 32, dtype=torch.uint8, device=dev)
        return vals, paths

    def open_columns(self, cols: Sequence[int]) -> Optional[List[LcColumn]]:
        """open_column (lcpc-2d/src/lib.rs:818-855) for each index.  The owner of a column block sends
        the column values together with the part of the Merkle path that lies inside its subtree;
        rank 0 appends the siblings above the subtree roots.  Result on rank 0 (None elsewhere)."""
        import numpy as np

        L, cb, W = self.L, self.cb, self.world
        assert self._pending is None, "a deferred commit is pending: call flush() on every rank first"
        for c in cols:
            if not 0 <= c < self.n_cols:
                from .lcpc2d import ProverError

                raise ProverError("ColumnNumber", "bad column number")
        if self.hashing == "rows" and hasattr(self.ops, "gather_paths"):
            res = self.open_columns_dev(cols)
            if res is None:
                return None
            # one device -> host copy each, into pinned staging buffers kept on the committer (a pageable destination costs
            # 20 ms for the 46 MB of a 309-column proof of the 4 GiB file; the arrays returned are copies of the staging)
            stage = getattr(self, "_open_stage", None)
            need = (res[0].numel(), res[1].numel())
            if stage is None or stage[0].numel() < need[0] or stage[1].numel() < need[1]:
                stage = (torch.empty(need[0], dtype=torch.int64).pin_memory(), torch.empty(need[1], dtype=torch.uint8).pin_memory())
                self._open_stage = stage
            hv, hp = stage[0][:need[0]], stage[1][:need[1]]
            hv.copy_(res[0].reshape(-1), non_blocking=True)
            hp.copy_(res[1].reshape(-1), non_blocking=True)
            torch.cuda.current_stream(res[0].device).synchronize()
            vals = hv.numpy().view(np.uint64).reshape(res[0].shape).copy()
            paths = hp.numpy().reshape(res[1].shape).copy()
            return [LcColumn(vals[k], paths[k]) for k in range(len(cols))]
        by_rows = self.hashing == "rows"  # the encoded matrix is row-sharded: values come from every rank, paths from the owner
        dev = (self.comm_rows if by_rows else self.comm_cols).device
        val_w = 0 if by_rows else self.n_rows * L  # words of column values in an owner's payload
        depth = log2(self.n_cols)
        depth_sub = min(depth, log2(cb))  # path levels inside a subtree
        owners = [c // cb for c in cols]
        mine = [c for c, o in zip(cols, owners) if o == self.rank]

        def my_payload():
            idx = torch.tensor([c - self.col0 for c in mine], dtype=torch.long, device=dev)
            if by_rows:
                vals = torch.empty(len(mine), 0, dtype=torch.int64, device=dev)
            else:
                m3 = self.comm_cols.view(self.n_rows, self.col_stride, L)
                vals = m3.index_select(1, idx).transpose(0, 1).contiguous().view(len(mine), -1)  # [k, n_rows*L]
            sub = self.subtree.view(-1, 32)
            parts, off, n, node = [], 0, cb, idx.clone()
            for _ in range(depth_sub):
                parts.append(sub.index_select(0, off + (node ^ 1)))
                off += n
                n //= 2
                node = node >> 1
            paths = torch.stack(parts, dim=1).reshape(len(mine), -1) if parts else torch.empty(len(mine), 0, dtype=torch.uint8, device=dev)
            # one int64 buffer per rank: values, then the path bytes (32-byte digests = 4 words each)
            return torch.cat([vals, paths.contiguous().view(torch.int64).view(len(mine), -1)], dim=1).contiguous()

        width = val_w + depth_sub * 4
        payload = my_payload() if mine else torch.empty(0, width, dtype=torch.int64, device=dev)
        gathered = None
        if by_rows:
            # every rank contributes its rows of the requested columns, padded to the longest row block
            max_rows = max(cnt for _, cnt in self.rows)
            part = torch.zeros(max_rows, len(cols), L, dtype=torch.int64, device=dev)
            if self.rows_local:
                idx_all = torch.tensor(list(cols), dtype=torch.long, device=dev)
                part[:self.rows_local] = self.comm_rows.view(self.rows_local, self.n_cols, L).index_select(1, idx_all)
            if W > 1:
                parts_all = [torch.empty_like(part) for _ in range(W)] if self.rank == 0 else None
                dist.gather(part, parts_all, dst=dist.get_global_rank(self.group, 0) if self.group else 0, group=self.group)
            else:
                parts_all = [part]
            if self.rank == 0:
                gathered = torch.cat([parts_all[q][:cnt] for q, (_, cnt) in enumerate(self.rows)], dim=0).cpu().numpy()
        counts = [sum(1 for o in owners if o == q) for q in range(W)]
        if self.rank == 0:
            bufs = [payload] + [torch.empty(counts[q], width, dtype=torch.int64, device=dev) for q in range(1, W)]
            ops = [dist.P2POp(dist.irecv, bufs[q], q, group=self.group) for q in range(1, W) if counts[q]]
        else:
            bufs, ops = None, ([dist.P2POp(dist.isend, payload, 0, group=self.group)] if mine else [])
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if self.rank != 0:
            return None
        host = [b.cpu().numpy() for b in bufs]
        top = self.top.cpu().numpy().reshape(-1, 32)
        out, seen = [], [0] * W
        for k, (c, q) in enumerate(zip(cols, owners)):
            row = host[q][seen[q]]
            seen[q] += 1
            if by_rows:
                col = np.ascontiguousarray(gathered[:, k]).view(np.uint64).reshape(self.n_rows, L)
            else:
                col = row[:val_w].view(np.uint64).reshape(self.n_rows, L)
            path = [row[val_w:].view(np.uint8).reshape(depth_sub, 32)] if depth_sub else []
            off, n, node = 0, W, q
            for _ in range(depth - depth_sub):  # siblings among / above the subtree roots
                path.append(top[off + (node ^ 1)][None])
                off += n
                n //= 2
                node >>= 1
            out.append(LcColumn(col, np.concatenate(path) if path else np.empty((0, 32), np.uint8)))
        return out


# The committer is encoding-agnostic: a Brakedown plan (non power-of-two n_cols) takes the all-to-all path over the padded
# leaf range; only the fused peer-store variant is specific to the NTT kernel.
ShardedCommitter = ShardedLigeroCommitter
