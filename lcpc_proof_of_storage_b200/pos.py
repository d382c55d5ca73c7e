"""Host-side mirror of proof-of-storage's commit layer (proof-of-storage/src/lcpc_online.rs and the
dimension helpers in networking/server.rs / client.rs) over the lcpc_b200 C ABI.

Only the commitment path is mirrored: file bytes -> WriteableFt63 elements -> Ligero commit, the
`Leaves` / `ColumnsWithPath` / `ColumnsWithoutPath` request kinds, and the client's retrievability
check.  Network protocol, database and on-disk formats are out of scope.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from ._lib import check
from .lcpc2d import (FIELD_LIMBS, FT63, FT253_192, Context, LcColumn, LcCommit, LigeroEncoding, ProverError, VerifierError,
                     default_context, log2, next_pow2)

DATA_BYTE_CAPACITY = 7   # WriteableFt63: CAPACITY / 8 (fields/data_field.rs:20, writable_ft63.rs:30)
WRITTEN_BYTES_WIDTH = 8  # size_of::<WriteableFt63>() (data_field.rs:22)


# ---- CommitDimensions / CommitRequestType (lcpc_online.rs:36-68) --------------------------------

@dataclass
class Specified:
    num_pre_encoded_columns: int
    num_encoded_columns: int


class Square:
    pass


@dataclass
class Commit:
    pass


@dataclass
class Leaves:
    cols: Sequence[int]


@dataclass
class ColumnsWithPath:
    cols: Sequence[int]


@dataclass
class ColumnsWithoutPath:
    cols: Sequence[int]


# ---- dimension helpers ----------------------------------------------------------------------------

def is_power_of_two(x: int) -> bool:
    return x > 0 and x & (x - 1) == 0


def dims_ok(num_pre_encoded_columns: int, num_encoded_columns: int) -> bool:
    """lcpc_online.rs:70-76."""
    return (is_power_of_two(num_encoded_columns) and num_pre_encoded_columns >= 1 and num_encoded_columns >= 2
            and num_encoded_columns >= 2 * num_pre_encoded_columns)


def get_soundness_from_matrix_dims(pre_encoded_cols: int, encoded_cols: int) -> int:
    """networking/server.rs:1160-1170: ceil(-128 / log2((1 + rho) / 2)), capped at the column count."""
    den = math.log2((1.0 + pre_encoded_cols / encoded_cols) / 2.0)
    return min(int(math.ceil(-128.0 / den)), encoded_cols)


def _sqrt_ceil_f32(n: int) -> int:
    """`(n as f32).sqrt().ceil() as usize` -- single precision, as the reference computes it."""
    return int(np.ceil(np.sqrt(np.float32(n))))


def get_aspect_ratio_default_from_field_len(field_len: int) -> Tuple[int, int, int]:
    """networking/server.rs:1139-1158 -> (num_pre_encoded_columns, num_encoded_columns, soundness)."""
    w = _sqrt_ceil_f32(field_len)
    pre = w if is_power_of_two(w) else next_pow2(w)
    enc = next_pow2(pre + 1)
    return pre, enc, get_soundness_from_matrix_dims(pre, enc)


def get_aspect_ratio_default_from_file_len(file_len: int) -> Tuple[int, int, int]:
    """networking/server.rs:1172-1182 (divides by WRITTEN_BYTES_WIDTH = 8, as the reference does)."""
    return get_aspect_ratio_default_from_field_len(-(-file_len // WRITTEN_BYTES_WIDTH))


def get_column_indicies_from_random_seed(random_seed: int, number_of_columns_to_extract: int,
                                         max_column_index: int) -> List[int]:
    """networking/client.rs:443-456: ChaCha8Rng::seed_from_u64 + choose_multiple (host-side)."""
    out = np.zeros(max(1, number_of_columns_to_extract), dtype=np.uint64)
    n = C.c_size_t()
    check(_lib.load().lcpc_pos_choose_columns(random_seed, number_of_columns_to_extract, max_column_index,
                                              out.ctypes.data, C.byref(n)))
    return [int(x) for x in out[:n.value]]


# ---- bytes <-> field elements ---------------------------------------------------------------------

def data_byte_capacity(field: int) -> int:
    """DataField::DATA_BYTE_CAPACITY (data_field.rs:22-24): bytes of file data one element carries."""
    if field == FT63:
        return DATA_BYTE_CAPACITY
    if field == FT253_192:
        return 31  # ft253_192.rs:15 `type DataBytes = [u8; 31]`
    raise ValueError("not a DataField: only WriteableFt63 and Ft253_192 carry file bytes")


def convert_byte_vec_to_field_elements_vec(data: bytes) -> np.ndarray:
    """DataField::from_byte_vec for WriteableFt63 (fields/data_field.rs:38-46, writable_ft63.rs:35-40):
    a pure byte shuffle -- the 7-byte little-endian integer IS the stored limb.  The commit path does
    this on the device (lcpc_commit_bytes_host); this host version exists for callers that want the Vec."""
    n = (len(data) + DATA_BYTE_CAPACITY - 1) // DATA_BYTE_CAPACITY
    buf = np.zeros(n * 8, dtype=np.uint8).reshape(n, 8)
    src = np.frombuffer(data, dtype=np.uint8)
    full = len(data) // 7
    buf[:full, :7] = src[:full * 7].reshape(full, 7)
    if full < n:
        tail = src[full * 7:]
        buf[full, :tail.shape[0]] = tail
    return buf.view(np.uint64).reshape(n, 1)


def field_vec_to_byte_vec(elems: np.ndarray) -> bytes:
    """DataField::field_vec_to_byte_vec (data_field.rs:54-60): the low 7 bytes of every limb."""
    e = np.ascontiguousarray(elems, dtype=np.uint64).reshape(-1, 1)
    return e.view(np.uint8).reshape(-1, 8)[:, :7].tobytes()


# ---- convert_file_data_to_commit (lcpc_online.rs:81-239) ---------------------------------------------

def _resolve_dims(data_len: int, dimensions) -> Tuple[int, int]:
    if isinstance(dimensions, Specified):
        pre, enc = dimensions.num_pre_encoded_columns, dimensions.num_encoded_columns
        if pre < 1:
            raise ValueError(f"Number of pre-encoded columns must be greater than 0, instead got {pre}")
        if enc < 2:
            raise ValueError(f"Number of pencoded columns must be greater than 0, instead got {enc}")
        if not is_power_of_two(enc):
            raise ValueError(f"Number of encoded columns must be a power of 2, instead got ratio of {pre}/{enc}")
        if not enc > pre:
            raise ValueError("Number of encoded columns must be greater than the number of columns")
        return pre, enc
    w = _sqrt_ceil_f32(data_len)  # CommitDimensions::Square (:120-129)
    pre = w if is_power_of_two(w) else next_pow2(w)
    return pre, next_pow2(pre + 1)


def convert_file_data_to_commit(data: Union[bytes, np.ndarray], what_to_extract, dimensions, ctx: Optional[Context] = None,
                                download: bool = True, field: int = FT63):
    """`data` is either the raw file bytes (packed on the device) or an (n, LIMBS) uint64 element array.
    Returns an LcCommit (Commit), an (n, 32) uint8 array (Leaves), a list of LcColumn (ColumnsWithPath)
    or a list of (n_rows, LIMBS) arrays (ColumnsWithoutPath).  `field` is the reference's `F: DataField` type parameter:
    FT63 (WriteableFt63, 7 data bytes per element: the server's PoSField) or FT253_192 (31 data bytes per element: the
    field of benches/commit_to_different_shapes_bench.rs)."""
    capacity = data_byte_capacity(field)
    if isinstance(data, (bytes, bytearray, memoryview)):
        data_len = (len(data) + capacity - 1) // capacity
    else:
        data = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, FIELD_LIMBS[field])
        data_len = data.shape[0]
    if data_len == 0:
        raise ValueError("Cannot convert empty file to commit")
    pre, enc_cols = _resolve_dims(data_len, dimensions)
    enc = LigeroEncoding(field, pre, enc_cols, ctx=ctx or default_context())
    want_full = isinstance(what_to_extract, Commit)
    if isinstance(data, np.ndarray):
        comm = LcCommit.commit(data, enc, download=download and want_full)
    else:
        comm = LcCommit.commit_bytes(bytes(data), enc, download=download and want_full)
    if want_full:
        return comm
    if isinstance(what_to_extract, Leaves):
        return comm.leaves(list(what_to_extract.cols))
    if isinstance(what_to_extract, ColumnsWithoutPath):
        return [c.col for c in comm.open_columns(list(what_to_extract.cols), with_path=False)]
    if isinstance(what_to_extract, ColumnsWithPath):
        return comm.open_columns(list(what_to_extract.cols))
    raise TypeError("unknown CommitRequestType")


def server_retreive_columns(comm: LcCommit, requested_columns: Sequence[int]) -> List[LcColumn]:
    """lcpc_online.rs:241-249."""
    return comm.open_columns(list(requested_columns))


# ---- client-side verification (lcpc_online.rs:251-452) -------------------------------------------------

def _verify_columns(ctx: Context, columns: Sequence[LcColumn], col_idx: Optional[Sequence[int]], root: Optional[bytes],
                    field: int = FT63):
    from .lcpc2d import FIELD_LIMBS

    n, L = len(columns), FIELD_LIMBS[field]
    flat = [np.ascontiguousarray(c.col, dtype=np.uint64).reshape(-1) for c in columns]
    n_rows = flat[0].shape[0] // L
    # ragged columns / paths cannot be passed flat: the reference fails the affected column (lib.rs:985-1030)
    if any(f.shape[0] != n_rows * L for f in flat) or flat[0].shape[0] % L:
        raise VerifierError("ColumnEval", "column eval invalid")
    if root is not None and any(np.asarray(c.path).shape != np.asarray(columns[0].path).shape for c in columns):
        raise VerifierError("ColumnPath", "column path invalid")
    cols = np.ascontiguousarray(np.stack([f.reshape(n_rows, L) for f in flat]))
    leaves = np.empty((n, 32), dtype=np.uint8)
    ok = np.zeros(n, dtype=np.uint32)
    paths = idx = rootb = None
    path_len = 0
    if root is not None:
        path_len = columns[0].path.shape[0]
        paths = np.ascontiguousarray(np.stack([c.path for c in columns])) if path_len else np.empty((n, 0, 32), np.uint8)
        idx = np.ascontiguousarray(np.asarray(col_idx, dtype=np.uint64))
        rootb = np.frombuffer(root, dtype=np.uint8).copy()
    p = lambda a: None if a is None else a.ctypes.data
    check(_lib.load().lcpc_verify_columns_host(ctx.handle, field, p(cols), n_rows, p(paths), path_len, p(idx), n, p(rootb),
                                               p(leaves), p(ok)))
    return leaves, ok


def hash_column_to_digest(column: LcColumn, ctx: Optional[Context] = None, field: int = FT63) -> bytes:
    """lcpc_online.rs:431-452."""
    leaves, _ = _verify_columns(ctx or default_context(), [column], None, None, field)
    return leaves[0].tobytes()


def client_online_verify_column_paths(commitment_root: bytes, requested_columns: Sequence[int],
                                      received_columns: Sequence[LcColumn], ctx: Optional[Context] = None,
                                      field: int = FT63) -> None:
    """lcpc_online.rs:251-281: every received column must hash up its path to the root."""
    if len(received_columns) != len(requested_columns):
        raise VerifierError("ColumnEval")
    if not received_columns:
        return
    _, ok = _verify_columns(ctx or default_context(), received_columns, requested_columns, commitment_root, field)
    if not ok.all():
        raise VerifierError("ColumnEval")


def client_online_verify_column_leaves(locally_derived_column_leaves: np.ndarray, requested_columns: Sequence[int],
                                       received_column_leaves: np.ndarray) -> None:
    """lcpc_online.rs:319-356 (the reference reports both failure kinds as NumColOpens)."""
    if len(locally_derived_column_leaves) != len(requested_columns) or len(received_column_leaves) != len(requested_columns):
        raise VerifierError("NumColOpens")
    if not np.array_equal(np.asarray(locally_derived_column_leaves), np.asarray(received_column_leaves)):
        raise VerifierError("NumColOpens")


def client_verify_commitment(commitment_root: bytes, locally_derived_column_leaves: np.ndarray,
                             requested_columns: Sequence[int], received_columns: Sequence[LcColumn],
                             required_columns_for_soundness: int, ctx: Optional[Context] = None,
                             field: int = FT63) -> None:
    """lcpc_online.rs:370-398."""
    if (required_columns_for_soundness < len(locally_derived_column_leaves)
            or required_columns_for_soundness < len(requested_columns)
            or required_columns_for_soundness < len(received_columns)):
        raise VerifierError("NumColOpens")
    ctx = ctx or default_context()
    if received_columns:
        received_leaves, _ = _verify_columns(ctx, received_columns, None, None, field)
    else:
        received_leaves = np.empty((0, 32), dtype=np.uint8)
    client_online_verify_column_leaves(locally_derived_column_leaves, requested_columns, received_leaves)
    client_online_verify_column_paths(commitment_root, requested_columns, received_columns, ctx, field)


def decode_row(row: np.ndarray, enc: LigeroEncoding) -> np.ndarray:
    """lcpc_online.rs:568-573: the coefficients (zero padded to n_cols) of one encoded row, or of a batch of rows."""
    out = np.array(row, dtype=np.uint64, copy=True, order="C")
    enc.decode(out)
    return out


def verifiable_polynomial_evaluation(commitment: LcCommit, left_evaluation_column: np.ndarray) -> np.ndarray:
    """lcpc_online.rs:454-484: result[j] = sum_r left[r] * comm[r][j] over the ENCODED matrix."""
    return commitment.fold(np.ascontiguousarray(left_evaluation_column, dtype=np.uint64).reshape(1, -1, 1), encoded=True)[0]
