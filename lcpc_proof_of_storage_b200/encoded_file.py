"""Host-side mirror of proof-of-storage's streaming file encoder over the C ABI's lcpc_stream_* calls.

Reference (proof-of-storage/src/lcpc_online/):
  encoded_file_writer.rs   EncodedFileWriter::{new, push_bytes, finalize_to_merkle_tree, finalize_to_commit,
                           finalize_to_column_digest, convert_unencoded_file}
  encoded_file_metadata.rs EncodedFileMetadata (JSON: ulid, pre_encoded_size, encoded_size, rows_written,
                           row_capacity, bytes_of_data)
  merkle_tree.rs           MerkleTree::{new, root, get_path, to_bytes, from_bytes}
  file_handler.rs:279-335  FileHandler::edit_bytes (row re-encode + tree rebuild)

The encoded file is column major: element (row r, column c) is stored as its canonical little-endian `to_repr`
bytes at byte offset (c * row_capacity + r) * 8 (encoded_file_writer.rs:327-349), with row_capacity = 2 * rows at
creation (:83).  Rows are encoded on the GPU block by block; the column digests advance one BLAKE3 chunk at a
time, so the file never has to fit in device memory.  Only byte buffering and file plumbing live here.
"""
from __future__ import annotations

import ctypes as C
import json
import mmap
import os
import time
from dataclasses import asdict, dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import check
from .lcpc2d import FT63, Context, LcCommit, LigeroEncoding, log2, next_pow2

DATA_BYTE_CAPACITY = 7    # WriteableFt63::DATA_BYTE_CAPACITY (fields/writable_ft63.rs:27)
WRITTEN_BYTES_WIDTH = 8   # size_of::<WriteableFt63>() (fields/data_field.rs:24)
_CROCKFORD = "0123456789ABCDEFGHJKMNPQRSTVWXYZ"


def _new_ulid() -> str:
    """A ULID string (48-bit millisecond timestamp + 80 random bits, Crockford base 32), as ulid::Ulid serialises."""
    v = (int(time.time() * 1000) << 80) | int.from_bytes(os.urandom(10), "big")
    return "".join(_CROCKFORD[(v >> (5 * i)) & 31] for i in reversed(range(26)))


@dataclass
class EncodedFileMetadata:
    """encoded_file_metadata.rs:5-27."""
    ulid: str
    pre_encoded_size: int
    encoded_size: int
    rows_written: int
    row_capacity: int
    bytes_of_data: int

    def write_to_file(self, f) -> None:
        f.write(json.dumps(asdict(self), separators=(",", ":")).encode())

    @classmethod
    def read_from_file(cls, f) -> "EncodedFileMetadata":
        return cls(**json.loads(f.read().decode()))


class MerkleTree:
    """merkle_tree.rs:7-100: `digests` = [width leaves | width/2 | ... | root]."""

    def __init__(self, digests: np.ndarray):
        digests = np.ascontiguousarray(digests, dtype=np.uint8).reshape(-1, 32)
        n = digests.shape[0]
        if (n + 1) & n or n < 3:
            raise ValueError("input size must be a power of two")
        self.digests = digests
        self.width = (n + 1) // 2

    def root(self) -> bytes:
        return self.digests[-1].tobytes()

    def get_path(self, index: int) -> Optional[List[bytes]]:
        if index >= self.width:
            return None
        path, off, n = [], 0, self.width
        for _ in range(log2(self.width)):
            path.append(self.digests[off + (index ^ 1)].tobytes())
            off += n
            n //= 2
            index >>= 1
        return path

    def __len__(self) -> int:
        return self.digests.shape[0]

    def __getitem__(self, i: int) -> bytes:
        return self.digests[i].tobytes()

    def to_bytes(self) -> bytes:
        return self.digests.tobytes()

    @classmethod
    def from_bytes(cls, data: bytes) -> "MerkleTree":
        if len(data) % 32:
            raise ValueError("input size must be a power of two")
        return cls(np.frombuffer(data, dtype=np.uint8).reshape(-1, 32).copy())


class EncodedFileWriter:
    """EncodedFileWriter<WriteableFt63, Blake3, LigeroEncoding<_>> (encoded_file_writer.rs:33-508)."""

    def __init__(self, num_pre_encoded_columns: int, num_encoded_columns: int, original_file_size: int,
                 target_file: Optional[str] = None, ctx: Optional[Context] = None, block_rows: int = 0):
        # the asserts of EncodedFileWriter::new (:52-71)
        assert num_encoded_columns & (num_encoded_columns - 1) == 0, "num_encoded_columns must be a power of two"
        assert num_pre_encoded_columns < num_encoded_columns, "num_pre_encoded_columns must be less than num_encoded_columns"
        assert num_pre_encoded_columns > 0, "num_pre_encoded_columns must be > 0"
        self.pre_encoded_size, self.encoded_size = num_pre_encoded_columns, num_encoded_columns
        self.encoding = LigeroEncoding(FT63, num_pre_encoded_columns, num_encoded_columns, ctx=ctx)
        self.row_bytes = num_pre_encoded_columns * DATA_BYTE_CAPACITY
        num_rows = -(-(-(-original_file_size // DATA_BYTE_CAPACITY)) // num_pre_encoded_columns)  # two div_ceils (:79-81)
        self.num_rows = max(1, num_rows)
        self.row_capacity = self.num_rows * 2  # :83
        self.bytes_received = 0
        self.incoming = bytearray()
        self._file = self._map = None
        sink_ptr = None
        if target_file is not None:
            desired = self.row_capacity * num_encoded_columns * WRITTEN_BYTES_WIDTH  # :92-94 set_len
            self._file = open(target_file, "w+b")
            self._file.truncate(desired)
            self._map = mmap.mmap(self._file.fileno(), desired)
            self._sink_view = np.frombuffer(self._map, dtype=np.uint8)
            sink_ptr = self._sink_view.ctypes.data
        self._s = C.c_void_p()
        check(_lib.load().lcpc_stream_begin(self.encoding.plan, self.num_rows, block_rows, sink_ptr,
                                            self.row_capacity if sink_ptr else 0, C.byref(self._s)))
        self._tree: Optional[MerkleTree] = None
        self.rows_written = 0

    # :211-241 -- rows are cut from the incoming byte buffer whenever at least one full row is available
    def push_bytes(self, data: bytes) -> None:
        self.bytes_received += len(data)
        self.incoming += data
        n_full = len(self.incoming) // self.row_bytes
        if n_full:
            take = n_full * self.row_bytes
            buf = np.frombuffer(bytes(self.incoming[:take]), dtype=np.uint8)
            check(_lib.load().lcpc_stream_push_bytes_host(self._s, buf.ctypes.data, take))
            del self.incoming[:take]

    def _finish(self) -> None:
        if self._tree is not None:
            return
        if self.incoming:  # the last, partially filled row (process_current_row(finalize = true), :449-452)
            buf = np.frombuffer(bytes(self.incoming), dtype=np.uint8)
            check(_lib.load().lcpc_stream_push_bytes_host(self._s, buf.ctypes.data, len(buf)))
            self.incoming.clear()
        hashes = np.empty((2 * next_pow2(self.encoded_size) - 1, 32), dtype=np.uint8)
        rows = C.c_size_t()
        check(_lib.load().lcpc_stream_finish(self._s, hashes.ctypes.data, C.byref(rows)))
        self.rows_written = rows.value
        self._tree = MerkleTree(hashes)
        if self._map is not None:
            self._map.flush()

    def get_encoded_file_metadata(self) -> EncodedFileMetadata:
        return EncodedFileMetadata(_new_ulid(), self.pre_encoded_size, self.encoded_size, self.rows_written,
                                   self.row_capacity, self.bytes_received)

    def finalize_to_merkle_tree(self) -> Tuple[EncodedFileMetadata, MerkleTree]:
        self._finish()
        return self.get_encoded_file_metadata(), self._tree

    def finalize_to_commit(self) -> Tuple[EncodedFileMetadata, bytes]:
        self._finish()
        return self.get_encoded_file_metadata(), self._tree.root()

    def finalize_to_column_digest(self) -> Tuple[EncodedFileMetadata, np.ndarray]:
        self._finish()
        return self.get_encoded_file_metadata(), self._tree.digests[:self.encoded_size].copy()

    def close(self) -> None:
        if getattr(self, "_s", None):
            _lib.load().lcpc_stream_free(self._s)
            self._s = None
        if self._map is not None:
            self._sink_view = None
            self._map.close()
            self._file.close()
            self._map = self._file = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def convert_unencoded_file(cls, unencoded_file: str, target_encoded_file: str, target_digest_file: Optional[str],
                               target_metadata_file: Optional[str], num_pre_encoded_columns: int,
                               num_encoded_columns: int, ctx: Optional[Context] = None,
                               read_size: int = 1 << 24) -> Tuple[EncodedFileMetadata, MerkleTree]:
        """encoded_file_writer.rs:134-209."""
        if num_pre_encoded_columns < 1:
            raise ValueError(f"Number of pre-encoded columns must be greater than 0, instead got {num_pre_encoded_columns}")
        if num_encoded_columns < 2 or num_encoded_columns & (num_encoded_columns - 1):
            raise ValueError("Number of encoded columns must be a power of 2, instead got ratio of "
                             f"{num_pre_encoded_columns}/{num_encoded_columns}")
        if num_encoded_columns <= num_pre_encoded_columns:
            raise ValueError("Number of encoded columns must be greater than the number of columns")
        total_size = os.path.getsize(unencoded_file)
        w = cls(num_pre_encoded_columns, num_encoded_columns, total_size, target_encoded_file, ctx=ctx)
        try:
            with open(unencoded_file, "rb") as f:
                while True:
                    chunk = f.read(read_size)
                    if not chunk:
                        break
                    w.push_bytes(chunk)
            metadata, tree = w.finalize_to_merkle_tree()
        finally:
            w.close()
        assert metadata.bytes_of_data == total_size
        assert metadata.rows_written == -(-total_size // (DATA_BYTE_CAPACITY * num_pre_encoded_columns))
        if target_metadata_file:
            with open(target_metadata_file, "wb") as f:
                metadata.write_to_file(f)
        if target_digest_file:
            with open(target_digest_file, "wb") as f:
                f.write(tree.to_bytes())
        return metadata, tree


def read_encoded_column(encoded_file: str, metadata: EncodedFileMetadata, column: int) -> np.ndarray:
    """EncodedFileReader::get_encoded_column_without_path's file access (encoded_file_reader.rs:214-253): the
    canonical (to_repr) integers of one column, rows_written of them."""
    with open(encoded_file, "rb") as f:
        f.seek(column * metadata.row_capacity * WRITTEN_BYTES_WIDTH)
        return np.frombuffer(f.read(metadata.rows_written * WRITTEN_BYTES_WIDTH), dtype="<u8").copy()


def edit_bytes(commit: LcCommit, total_data_bytes: int, byte_start: int, new_bytes: bytes) -> Tuple[bytes, MerkleTree]:
    """FileHandler::edit_bytes (file_handler.rs:279-335) on a device-resident commitment of a file: overwrite
    `new_bytes` at `byte_start`, re-encode the touched rows, rebuild the tree.  Returns (original bytes, new tree)."""
    if byte_start + len(new_bytes) > total_data_bytes:
        raise ValueError("can't edit more bytes than there are in the file!")
    row_bytes = commit.n_per_row * DATA_BYTE_CAPACITY
    start_row = byte_start // row_bytes
    end_row = -(-(byte_start + len(new_bytes)) // row_bytes)
    rows = commit.coeffs[start_row:end_row, :, 0]
    raw = bytearray(np.ascontiguousarray(rows).view(np.uint8).reshape(-1, 8)[:, :7].tobytes())  # to_data_bytes (:42-46)
    lo = byte_start - start_row * row_bytes
    original = bytes(raw[lo:lo + len(new_bytes)])
    raw[lo:lo + len(new_bytes)] = new_bytes
    b = np.frombuffer(bytes(raw), dtype=np.uint8).reshape(-1, 7).astype(np.uint64)
    elems = np.zeros(b.shape[0], dtype=np.uint64)
    for k in range(7):
        elems |= b[:, k] << np.uint64(8 * k)
    tree = commit.update_rows(start_row, elems.reshape(end_row - start_row, commit.n_per_row, 1))
    return original, MerkleTree(tree)


def append_bytes(commit: LcCommit, total_data_bytes: int, bytes_to_add: bytes) -> MerkleTree:
    """FileHandler::append_bytes (file_handler.rs:336-402) on a device-resident commitment of a file of
    `total_data_bytes` bytes: the partially filled last row is completed and re-encoded, further rows are appended."""
    row_bytes = commit.n_per_row * DATA_BYTE_CAPACITY
    start_row = total_data_bytes // row_bytes
    head = b""
    if start_row < commit.n_rows:  # bytes already in the last, partially filled row
        row = commit.coeffs[start_row, :, 0]
        head = np.ascontiguousarray(row).view(np.uint8).reshape(-1, 8)[:, :7].tobytes()[:total_data_bytes - start_row * row_bytes]
    raw = head + bytes_to_add
    n_rows = -(-len(raw) // row_bytes)
    buf = np.zeros(n_rows * row_bytes, dtype=np.uint8)
    buf[:len(raw)] = np.frombuffer(raw, dtype=np.uint8)
    b = buf.reshape(-1, 7).astype(np.uint64)
    elems = np.zeros(b.shape[0], dtype=np.uint64)
    for k in range(7):
        elems |= b[:, k] << np.uint64(8 * k)
    return MerkleTree(commit.append_rows(start_row, elems.reshape(n_rows, commit.n_per_row, 1)))
