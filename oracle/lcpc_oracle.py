"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes front-end to oracle/_build/liborc.so (the C restatement of the reference's
CPU algorithm) plus the host-side sequencing of the scheme (dimension selection,
prove, verify) restated in Python.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs may import this module; the
product package never does.

Field elements are numpy ``uint64`` arrays of shape ``(n, LIMBS)`` holding the
Montgomery residue, least-significant limb first -- bit-identical to a Rust
``&[F]`` (SURVEY.md section 8 a').  Paths cited are relative to /root/reference.

Parity status: blake3 / ChaCha20 / Keccak / merlin / field arithmetic are pinned
against independent implementations (tests/test_oracle_*.py).  PARITY UNPINNED (two
items, no reference-attested vector exists and the Rust reference cannot be built
here): the Ligero NTT convention (fffft, not in the tree) and the consumers of the
ChaCha streams as rand 0.8 / rand_chacha 0.3 / ff 0.13 define them (seed_from_u64,
set_stream, Uniform, choose_multiple, F::random), both restated as recalled;
rust/lcpc-b200/tests/parity.rs pins them where cargo exists.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass, field as dc_field
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liborc.so")

FT63, FT127, FT191, FT255, FT253_192 = 0, 1, 2, 3, 4
FIELD_NAMES = {FT63: "Ft63", FT127: "Ft127", FT191: "Ft191", FT255: "Ft255", FT253_192: "Ft253_192"}
LIMBS = {FT63: 1, FT127: 2, FT191: 3, FT255: 4, FT253_192: 4}
# lcpc-test-fields/src/lib.rs:19,42,54,66 and proof-of-storage/src/fields/ft253_192.rs:7 (PrimeFieldModulus)
MODULUS = {
    FT63: 5102708120182849537,
    FT127: 146823888364060453008360742206866194433,
    FT191: 1697146272512170708389931801544665676545308500647389167617,
    FT255: 46242760681095663677370860714659204618859642560429202607213929836750194081793,
    FT253_192: 14474011154664524421669271390699307717822958659997404088829842556525106692097,
}
GENERATOR = {FT63: 10, FT127: 3, FT191: 5, FT255: 5, FT253_192: 3}
NUM_BITS = {f: MODULUS[f].bit_length() for f in MODULUS}
TWO_ADICITY = {FT63: 41, FT127: 40, FT191: 41, FT255: 41, FT253_192: 192}
# PrimeFieldReprEndianness (ft253_192.rs:9): to_repr() of Ft253_192 is the canonical value BIG-endian
REPR_BIG_ENDIAN = {FT63: False, FT127: False, FT191: False, FT255: False, FT253_192: True}


def build(force: bool = False) -> str:
    """Compile the C oracle (make -C oracle).  Building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib = None
u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


class _Csc(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("cols", C.c_uint64), ("indptr", u64p),
                ("indices", u64p), ("data", u64p)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_sdig_dist.restype = C.c_double
        _lib.orc_sdig_get_dims.argtypes = [C.c_int, C.c_uint64, C.c_double, u64p, u64p, C.c_int]
    return _lib


def _p64(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def _p8(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u8p)


def set_threads(n: int) -> None:
    lib().orc_set_threads(C.c_int(n))


def max_threads() -> int:
    return int(lib().orc_get_max_threads())


# ----------------------------------------------------------------------------- field helpers

def to_limbs(fid: int, values: Sequence[int]) -> np.ndarray:
    """Python ints (raw limb integers, no conversion) -> (n, LIMBS) uint64."""
    L = LIMBS[fid]
    out = np.zeros((len(values), L), dtype=np.uint64)
    mask = (1 << 64) - 1
    for i, v in enumerate(values):
        for l in range(L):
            out[i, l] = (v >> (64 * l)) & mask
    return out


def from_limbs(a: np.ndarray) -> List[int]:
    a = np.asarray(a, dtype=np.uint64)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    return [sum(int(x) << (64 * l) for l, x in enumerate(row)) for row in a]


def mont_R(fid: int) -> int:
    return (1 << (64 * LIMBS[fid])) % MODULUS[fid]


def to_mont(fid: int, values: Sequence[int]) -> np.ndarray:
    p, R = MODULUS[fid], mont_R(fid)
    return to_limbs(fid, [(v % p) * R % p for v in values])


def from_mont(fid: int, a: np.ndarray) -> List[int]:
    p = MODULUS[fid]
    Rinv = pow(mont_R(fid), -1, p)
    return [v * Rinv % p for v in from_limbs(a)]


def _binop(fid, op, a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty_like(a)
    lib().orc_fe_binop(C.c_int(fid), C.c_int(op), _p64(out), _p64(a), _p64(b), C.c_size_t(a.size // LIMBS[fid]))
    return out


def fe_add(fid, a, b):
    return _binop(fid, 0, a, b)


def fe_sub(fid, a, b):
    return _binop(fid, 1, a, b)


def fe_mul(fid, a, b):
    return _binop(fid, 2, a, b)


def fe_to_canon(fid, a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    lib().orc_fe_to_canon(C.c_int(fid), _p64(out), _p64(a), C.c_size_t(a.size // LIMBS[fid]))
    return out


def fe_to_repr(fid, a) -> bytes:
    """PrimeField::to_repr() of every element, concatenated (8 * LIMBS bytes each; big-endian for Ft253_192)."""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty(a.size * 8, dtype=np.uint8)
    lib().orc_fe_to_repr(C.c_int(fid), out.ctypes.data_as(u8p), _p64(a), C.c_size_t(a.size // LIMBS[fid]))
    return out.tobytes()


def fe_from_canon(fid, a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    lib().orc_fe_from_canon(C.c_int(fid), _p64(out), _p64(a), C.c_size_t(a.size // LIMBS[fid]))
    return out


def fe_inv(fid, a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    lib().orc_fe_inv(C.c_int(fid), _p64(out), _p64(a), C.c_size_t(a.size // LIMBS[fid]))
    return out


def ntt_root(fid: int, log_n: int) -> np.ndarray:
    """The 2^log_n-th root of unity precomp_fft uses: ROOT_OF_UNITY^(2^(S-log_n)), Montgomery."""
    out = np.zeros((1, LIMBS[fid]), dtype=np.uint64)
    lib().orc_ntt_root(C.c_int(fid), C.c_int(log_n), _p64(out))
    return out


def fft_io(fid: int, rows: np.ndarray) -> np.ndarray:
    """fffft fft_io on each row (rows: (n_rows, n, LIMBS)); returns a new array."""
    rows = np.array(rows, dtype=np.uint64, copy=True, order="C")
    n_rows, n = rows.shape[0], rows.shape[1]
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    lib().orc_fft_io(C.c_int(fid), _p64(rows), C.c_int(log_n), C.c_size_t(n_rows))
    return rows


def ifft_oi(fid: int, rows: np.ndarray) -> np.ndarray:
    rows = np.array(rows, dtype=np.uint64, copy=True, order="C")
    n_rows, n = rows.shape[0], rows.shape[1]
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    lib().orc_ifft_oi(C.c_int(fid), _p64(rows), C.c_int(log_n), C.c_size_t(n_rows))
    return rows


# ----------------------------------------------------------------------------- hashing

def blake3(data: bytes) -> bytes:
    out = np.zeros(32, dtype=np.uint8)
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(0, dtype=np.uint8)
    buf = np.ascontiguousarray(buf)
    lib().orc_blake3(buf.ctypes.data_as(C.c_void_p), C.c_size_t(len(data)), _p8(out))
    return out.tobytes()


def hash_columns(fid: int, comm: np.ndarray) -> np.ndarray:
    """comm: (n_rows, n_cols, LIMBS) -> (n_cols, 32) uint8 leaves (lcpc-2d/src/lib.rs:736-775)."""
    comm = np.ascontiguousarray(comm, dtype=np.uint64)
    n_rows, n_cols = comm.shape[0], comm.shape[1]
    out = np.zeros((n_cols, 32), dtype=np.uint8)
    lib().orc_hash_columns(C.c_int(fid), _p64(comm), _p8(out), C.c_size_t(n_rows), C.c_size_t(n_cols),
                           C.c_size_t(n_cols))
    return out


def hash_column(fid: int, col: np.ndarray) -> bytes:
    col = np.ascontiguousarray(col, dtype=np.uint64)
    out = np.zeros(32, dtype=np.uint8)
    lib().orc_hash_column(C.c_int(fid), _p64(col), C.c_size_t(col.size // LIMBS[fid]), _p8(out))
    return out.tobytes()


def next_pow2(v: int) -> int:
    return 1 if v <= 1 else 1 << (v - 1).bit_length()


def log2(v: int) -> int:
    """lcpc-2d/src/lib.rs:857-859: log2 of next_power_of_two(v)."""
    return next_pow2(v).bit_length() - 1


def merkle_tree(leaves: np.ndarray) -> np.ndarray:
    """leaves: (np2, 32) -> full (2*np2-1, 32) array (lib.rs:777-815)."""
    np2 = leaves.shape[0]
    assert np2 & (np2 - 1) == 0
    out = np.zeros((2 * np2 - 1, 32), dtype=np.uint8)
    out[:np2] = leaves
    lib().orc_merkle_tree(_p8(out), C.c_size_t(np2))
    return out


# ----------------------------------------------------------------------------- encodings

def n_degree_tests(lam: int, length: int, flog2: int) -> int:
    """lcpc-2d/src/lib.rs:642-645."""
    den = flog2 - log2(length)
    return (lam + den - 1) // den


class LigeroEncoding:
    """lcpc-ligero-pc/src/lib.rs:31-186 LigeroEncodingRho<Ft, Rn, Rd> (default rho = 1/2, :189)."""

    LAMBDA = 128

    def __init__(self, fid: int, n_per_row: int, n_cols: int, rho_num: int = 1, rho_den: int = 2,
                 n_col_opens: Optional[int] = None, n_degree_tests_: Optional[int] = None):
        assert self._dims_ok(n_per_row, n_cols)
        self.fid, self.n_per_row, self.n_cols = fid, n_per_row, n_cols
        self.rho_num, self.rho_den = rho_num, rho_den
        self._n_col_opens_override = n_col_opens
        self._n_dt_override = n_degree_tests_

    # :61-64
    @classmethod
    def _n_col_opens(cls, rho_num: int, rho_den: int) -> int:
        rho = rho_num / rho_den
        den = math.log2((1.0 + rho) / 2.0)
        return int(math.ceil(-cls.LAMBDA / den))

    # :66-68
    @classmethod
    def _n_degree_tests(cls, fid: int, n_cols: int) -> int:
        return n_degree_tests(cls.LAMBDA, n_cols, NUM_BITS[fid] - 1)

    # :70-112
    @classmethod
    def get_dims_for_len(cls, fid: int, length: int, rho_num: int = 1, rho_den: int = 2) -> Tuple[int, int, int]:
        rho = rho_num / rho_den
        n_col_opens = cls._n_col_opens(rho_num, rho_den)
        lncf = float(n_col_opens * length)
        ndt = float(cls._n_degree_tests(fid, int(math.ceil(math.sqrt(lncf) / rho))))
        nc1 = next_pow2(int(math.ceil(math.sqrt(lncf / ndt) / rho)))
        assert nc1 <= (1 << TWO_ADICITY[fid])
        np1 = nc1 * rho_num // rho_den
        nr1 = (length + np1 - 1) // np1
        nd1 = cls._n_degree_tests(fid, nc1)
        nc2, np2 = nc1 // 2, np1 // 2
        nr2 = (length + np2 - 1) // np2
        nd2 = cls._n_degree_tests(fid, nc2)
        sz1 = n_col_opens * nr1 + (1 + nd1) * np1
        sz2 = n_col_opens * nr2 + (1 + nd2) * np2
        return (nr1, np1, nc1) if sz1 < sz2 else (nr2, np2, nc2)

    @classmethod
    def new(cls, fid: int, length: int, rho_num: int = 1, rho_den: int = 2) -> "LigeroEncoding":
        _, n_per_row, n_cols = cls.get_dims_for_len(fid, length, rho_num, rho_den)
        return cls(fid, n_per_row, n_cols, rho_num, rho_den)

    @staticmethod
    def _dims_ok(n_per_row: int, n_cols: int) -> bool:
        return n_per_row < n_cols and n_cols & (n_cols - 1) == 0 and n_cols > 0

    def get_dims(self, length: int) -> Tuple[int, int, int]:
        return ((length + self.n_per_row - 1) // self.n_per_row, self.n_per_row, self.n_cols)

    def dims_ok(self, n_per_row: int, n_cols: int) -> bool:
        return self._dims_ok(n_per_row, n_cols) and n_per_row == self.n_per_row and n_cols == self.n_cols

    def get_n_col_opens(self) -> int:
        if self._n_col_opens_override is not None:
            return self._n_col_opens_override
        return self._n_col_opens(self.rho_num, self.rho_den)

    def get_n_degree_tests(self) -> int:
        if self._n_dt_override is not None:
            return self._n_dt_override
        return self._n_degree_tests(self.fid, self.n_cols)

    def encode_rows(self, rows: np.ndarray) -> np.ndarray:
        return fft_io(self.fid, rows)


@dataclass
class CscMatrix:
    """sprs::CsMat in CSC storage (rows x cols)."""
    rows: int
    cols: int
    indptr: np.ndarray   # (cols+1,) uint64
    indices: np.ndarray  # (nnz,) uint64 row numbers
    data: np.ndarray     # (nnz, LIMBS) uint64 Montgomery

    def as_struct(self) -> _Csc:
        return _Csc(self.rows, self.cols, _p64(self.indptr), _p64(self.indices), _p64(self.data))


def sdig_get_dims(code: int, n: int, flog2: int):
    """matgen.rs:56-111 -> ([(ni, mi, cn)], [(nip, mip, dn)])."""
    pre = np.zeros(3 * 64, dtype=np.uint64)
    post = np.zeros(3 * 64, dtype=np.uint64)
    levels = lib().orc_sdig_get_dims(C.c_int(code), C.c_uint64(n), C.c_double(float(flog2)), _p64(pre), _p64(post), 64)
    assert levels > 0, "n must exceed the base-case length"
    pre = [tuple(int(x) for x in pre[3 * i:3 * i + 3]) for i in range(levels)]
    post = [tuple(int(x) for x in post[3 * i:3 * i + 3]) for i in range(levels)]
    return pre, post


def sdig_generate(fid: int, code: int, n: int, seed: int):
    """matgen.rs:28-53 generate -> (precodes, postcodes) as CscMatrix lists."""
    pre_dims, post_dims = sdig_get_dims(code, n, NUM_BITS[fid] - 1)
    L = LIMBS[fid]
    pres, posts = [], []
    for lvl, ((ni, mi, cn), (nip, mip, dn)) in enumerate(zip(pre_dims, post_dims)):
        a = CscMatrix(mi, ni, np.zeros(ni + 1, np.uint64), np.zeros(ni * cn, np.uint64), np.zeros((ni * cn, L), np.uint64))
        b = CscMatrix(mip, nip, np.zeros(nip + 1, np.uint64), np.zeros(nip * dn, np.uint64), np.zeros((nip * dn, L), np.uint64))
        pd = np.array([ni, mi, cn], dtype=np.uint64)
        qd = np.array([nip, mip, dn], dtype=np.uint64)
        lib().orc_sdig_gen_level(C.c_int(fid), C.c_uint64(seed), C.c_uint64(lvl), _p64(pd), _p64(qd),
                                 _p64(a.indptr), _p64(a.indices), _p64(a.data),
                                 _p64(b.indptr), _p64(b.indices), _p64(b.data))
        pres.append(a)
        posts.append(b)
    return pres, posts


def sdig_codeword_length(pre: List[CscMatrix], post: List[CscMatrix]) -> int:
    """encode.rs:18-33."""
    return pre[0].cols + post[-1].cols + sum(m.rows for m in pre[:-1]) + sum(m.rows for m in post)


class SdigEncoding:
    """lcpc-brakedown-pc/src/lib.rs:38-176 SdigEncodingS<Ft, S> (default S = SdigCode3, :19)."""

    LAMBDA = 128

    def __init__(self, fid: int, n_per_row: int, seed: int, code: int = 3, n_cols: Optional[int] = None):
        self.fid, self.code, self.seed = fid, code, seed
        self.precodes, self.postcodes = sdig_generate(fid, code, n_per_row, seed)
        self.n_per_row = n_per_row
        assert self.precodes[0].cols == n_per_row
        self.n_cols = sdig_codeword_length(self.precodes, self.postcodes)
        if n_cols is not None:
            assert n_cols == self.n_cols

    @classmethod
    def _n_col_opens(cls, code: int) -> int:  # :57-61
        dist = float(lib().orc_sdig_dist(C.c_int(code)))
        den = math.log2(1.0 - dist / 3.0)
        return int(math.ceil(-cls.LAMBDA / den))

    @classmethod
    def _n_degree_tests(cls, fid: int, n_cols: int) -> int:  # :64-66
        return n_degree_tests(cls.LAMBDA, n_cols, NUM_BITS[fid] - 1)

    @classmethod
    def n_per_row_for_len(cls, fid: int, length: int, code: int = 3, ml: bool = False) -> int:
        """:69-123 new / new_ml / _new_from_np1 dimension choice."""
        lncf = float(cls._n_col_opens(code) * length)
        ndt = float(cls._n_degree_tests(fid, int(math.ceil(math.sqrt(lncf))) * 2))
        np1 = int(math.ceil(math.sqrt(lncf / ndt)))
        if ml:
            np1 = next_pow2(np1)
        np1 = min(np1, length)
        n_col_opens = cls._n_col_opens(code)
        nr1 = (length + np1 - 1) // np1
        nd1 = cls._n_degree_tests(fid, np1 * 2)
        np2 = np1 // 2
        nr2 = (length + np2 - 1) // np2
        nd2 = cls._n_degree_tests(fid, np2 * 2)
        sz1 = n_col_opens * nr1 + (1 + nd1) * np1
        sz2 = n_col_opens * nr2 + (1 + nd2) * np2
        return np1 if sz1 < sz2 else np2

    @classmethod
    def new(cls, fid: int, length: int, seed: int, code: int = 3) -> "SdigEncoding":
        return cls(fid, cls.n_per_row_for_len(fid, length, code), seed, code)

    def get_dims(self, length: int):
        return ((length + self.n_per_row - 1) // self.n_per_row, self.n_per_row, self.n_cols)

    def dims_ok(self, n_per_row: int, n_cols: int) -> bool:
        return n_per_row < n_cols and n_per_row == self.n_per_row and n_cols == self.n_cols

    def get_n_col_opens(self) -> int:
        return self._n_col_opens(self.code)

    def get_n_degree_tests(self) -> int:
        return self._n_degree_tests(self.fid, self.n_cols)

    def _structs(self):
        n = len(self.precodes)
        pre = (_Csc * n)(*[m.as_struct() for m in self.precodes])
        post = (_Csc * n)(*[m.as_struct() for m in self.postcodes])
        return n, pre, post

    def encode_rows(self, rows: np.ndarray) -> np.ndarray:
        rows = np.array(rows, dtype=np.uint64, copy=True, order="C")
        n, pre, post = self._structs()
        lib().orc_sdig_encode_rows(C.c_int(self.fid), _p64(rows), C.c_size_t(rows.shape[0]),
                                   C.c_size_t(self.n_cols), C.c_size_t(n), pre, post)
        return rows


# ----------------------------------------------------------------------------- commit / prove / verify

@dataclass
class LcCommit:
    """lcpc-2d/src/lib.rs:174-191."""
    fid: int
    comm: np.ndarray    # (n_rows, n_cols, LIMBS)
    coeffs: np.ndarray  # (n_rows, n_per_row, LIMBS)
    n_rows: int
    n_cols: int
    n_per_row: int
    hashes: np.ndarray  # (2*np2-1, 32)

    def get_root(self) -> bytes:  # :291-296
        return self.hashes[-1].tobytes()


@dataclass
class LcColumn:
    col: np.ndarray          # (n_rows, LIMBS)
    path: np.ndarray         # (path_len, 32)


@dataclass
class LcEvalProof:
    n_cols: int
    p_eval: np.ndarray                  # (n_per_row, LIMBS)
    p_random_vec: List[np.ndarray]
    columns: List[LcColumn]


class ProverError(Exception):
    pass


class VerifierError(Exception):
    pass


def commit(coeffs_in: np.ndarray, enc) -> LcCommit:
    """lib.rs:651-700."""
    fid = enc.fid
    L = LIMBS[fid]
    coeffs_in = np.ascontiguousarray(coeffs_in, dtype=np.uint64).reshape(-1, L)
    length = coeffs_in.shape[0]
    n_rows, n_per_row, n_cols = enc.get_dims(length)
    assert n_rows * n_per_row >= length and (n_rows - 1) * n_per_row < length
    assert enc.dims_ok(n_per_row, n_cols)
    np2 = next_pow2(n_cols)
    coeffs = np.zeros((n_rows, n_per_row, L), dtype=np.uint64)
    comm = np.zeros((n_rows, n_cols, L), dtype=np.uint64)
    hashes = np.zeros((2 * np2 - 1, 32), dtype=np.uint8)
    if isinstance(enc, LigeroEncoding):
        rc = lib().orc_commit_ligero(C.c_int(fid), _p64(coeffs_in), C.c_size_t(length), C.c_size_t(n_per_row),
                                     C.c_size_t(n_cols), _p64(coeffs), _p64(comm), _p8(hashes))
    else:
        n, pre, post = enc._structs()
        rc = lib().orc_commit_sdig(C.c_int(fid), _p64(coeffs_in), C.c_size_t(length), C.c_size_t(n_per_row),
                                   C.c_size_t(n_cols), C.c_size_t(n), pre, post, _p64(coeffs), _p64(comm),
                                   _p8(hashes))
    if rc != 0:
        raise ProverError("Commit")
    return LcCommit(fid, comm, coeffs, n_rows, n_cols, n_per_row, hashes)


def collapse_columns(fid: int, coeffs: np.ndarray, tensor: np.ndarray) -> np.ndarray:
    """lib.rs:1126-1154: poly[j] = sum_r coeffs[r][j] * tensor[r]."""
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64)
    tensor = np.ascontiguousarray(tensor, dtype=np.uint64)
    n_rows, width = coeffs.shape[0], coeffs.shape[1]
    poly = np.zeros((width, LIMBS[fid]), dtype=np.uint64)
    lib().orc_collapse_columns(C.c_int(fid), _p64(coeffs), _p64(tensor), _p64(poly), C.c_size_t(n_rows),
                               C.c_size_t(width))
    return poly


def open_column(comm: LcCommit, column: int) -> LcColumn:
    """lib.rs:818-855."""
    if column >= comm.n_cols or column < 0:
        raise ProverError("ColumnNumber")
    L = LIMBS[comm.fid]
    col = np.zeros((comm.n_rows, L), dtype=np.uint64)
    path = np.zeros((log2(comm.n_cols), 32), dtype=np.uint8)
    rc = lib().orc_open_column(C.c_int(comm.fid), _p64(comm.comm), _p8(comm.hashes), C.c_size_t(comm.n_rows),
                               C.c_size_t(comm.n_cols), C.c_size_t(column), _p64(col), _p8(path))
    assert rc == 0
    return LcColumn(col, path)


def verify_column_path(fid: int, column: LcColumn, col_num: int, root: bytes) -> bool:
    """lib.rs:985-1012."""
    r = np.frombuffer(root, dtype=np.uint8).copy()
    col = np.ascontiguousarray(column.col)
    path = np.ascontiguousarray(column.path)
    return bool(lib().orc_verify_column_path(C.c_int(fid), _p64(col), C.c_size_t(col.shape[0]), _p8(path),
                                             C.c_size_t(path.shape[0]), C.c_size_t(col_num), _p8(r)))


def verify_column_value(fid: int, column: LcColumn, tensor: np.ndarray, poly_eval: np.ndarray) -> bool:
    """lib.rs:1015-1030."""
    col = np.ascontiguousarray(column.col)
    tensor = np.ascontiguousarray(tensor, dtype=np.uint64)
    pe = np.ascontiguousarray(poly_eval, dtype=np.uint64)
    return bool(lib().orc_verify_column_value(C.c_int(fid), _p64(col), _p64(tensor), C.c_size_t(col.shape[0]), _p64(pe)))


class _OrcTranscript(C.Structure):
    _fields_ = [("state", C.c_uint8 * 200), ("pos", C.c_uint8), ("pos_begin", C.c_uint8), ("cur_flags", C.c_uint8)]


class Transcript:
    """merlin::Transcript (new / append_message / challenge_bytes)."""

    def __init__(self, label: bytes):
        self._t = _OrcTranscript()
        lib().orc_transcript_new(C.byref(self._t), label, C.c_size_t(len(label)))

    def append_message(self, label: bytes, msg: bytes) -> None:
        lib().orc_transcript_append_message(C.byref(self._t), label, C.c_size_t(len(label)), msg, C.c_size_t(len(msg)))

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        out = (C.c_uint8 * n)()
        lib().orc_transcript_challenge_bytes(C.byref(self._t), label, C.c_size_t(len(label)), out, C.c_size_t(n))
        return bytes(out)


# lcpc-2d/src/macros.rs:28-36: `$l` is not substituted inside a byte-string
# literal, so every encoding uses these literal six bytes.
LABEL_DT = b"$l//DT"
LABEL_PR = b"$l//PR"
LABEL_PE = b"$l//PE"
LABEL_CO = b"$l//CO"


def random_field_vec(fid: int, key: bytes, n: int) -> np.ndarray:
    out = np.zeros((n, LIMBS[fid]), dtype=np.uint64)
    k = np.frombuffer(key, dtype=np.uint8).copy()
    lib().orc_random_field_vec(C.c_int(fid), _p8(k), _p64(out), C.c_size_t(n))
    return out


def random_columns(key: bytes, n_cols: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint64)
    k = np.frombuffer(key, dtype=np.uint8).copy()
    lib().orc_random_columns(_p8(k), C.c_uint64(n_cols), _p64(out), C.c_size_t(n))
    return out


def _transcript_update(tr: Transcript, label: bytes, fid: int, elems: np.ndarray) -> None:
    """FieldHash::transcript_update per element (lib.rs:48-50): to_repr bytes, one message each."""
    w = 8 * LIMBS[fid]
    raw = fe_to_repr(fid, elems)
    for i in range(len(raw) // w):
        tr.append_message(label, raw[i * w:(i + 1) * w])


def prove(comm: LcCommit, outer_tensor: np.ndarray, enc, tr: Transcript) -> LcEvalProof:
    """lib.rs:1034-1123."""
    fid = comm.fid
    if outer_tensor.shape[0] != comm.n_rows:
        raise ProverError("OuterTensor")
    p_random_vec = []
    for _ in range(enc.get_n_degree_tests()):
        key = tr.challenge_bytes(LABEL_DT, 32)
        rand_tensor = random_field_vec(fid, key, comm.n_rows)
        p_random = collapse_columns(fid, comm.coeffs, rand_tensor)
        _transcript_update(tr, LABEL_PR, fid, p_random)
        p_random_vec.append(p_random)
    p_eval = collapse_columns(fid, comm.coeffs, outer_tensor)
    _transcript_update(tr, LABEL_PE, fid, p_eval)
    key = tr.challenge_bytes(LABEL_CO, 32)
    cols = random_columns(key, comm.n_cols, enc.get_n_col_opens())
    columns = [open_column(comm, int(c)) for c in cols]
    return LcEvalProof(comm.n_cols, p_eval, p_random_vec, columns)


def verify(root: bytes, outer_tensor: np.ndarray, inner_tensor: np.ndarray, proof: LcEvalProof, enc,
           tr: Transcript) -> np.ndarray:
    """lib.rs:862-982; raises VerifierError(<variant name>) like the Rust enum (:139-167)."""
    fid = enc.fid
    L = LIMBS[fid]
    n_col_opens = enc.get_n_col_opens()
    if n_col_opens != len(proof.columns) or n_col_opens == 0:
        raise VerifierError("NumColOpens")
    n_rows = proof.columns[0].col.shape[0]
    n_cols = proof.n_cols
    n_per_row = proof.p_eval.shape[0]
    if inner_tensor.shape[0] != n_per_row:
        raise VerifierError("InnerTensor")
    if outer_tensor.shape[0] != n_rows:
        raise VerifierError("OuterTensor")
    if not enc.dims_ok(n_per_row, n_cols):
        raise VerifierError("EncodingDims")

    def encode_one(v: np.ndarray) -> np.ndarray:
        tmp = np.zeros((1, n_cols, L), dtype=np.uint64)
        tmp[0, :n_per_row] = v
        return enc.encode_rows(tmp)[0]

    rand_tensors, p_random_fft = [], []
    n_dt = enc.get_n_degree_tests()
    for i in range(n_dt):
        key = tr.challenge_bytes(LABEL_DT, 32)
        rand_tensors.append(random_field_vec(fid, key, n_rows))
        p_random_fft.append(encode_one(proof.p_random_vec[i]))
        _transcript_update(tr, LABEL_PR, fid, proof.p_random_vec[i])
    _transcript_update(tr, LABEL_PE, fid, proof.p_eval)
    key = tr.challenge_bytes(LABEL_CO, 32)
    cols = random_columns(key, n_cols, n_col_opens)
    p_eval_fft = encode_one(proof.p_eval)
    for col_num, column in zip(cols, proof.columns):
        col_num = int(col_num)
        rand = all(verify_column_value(fid, column, rand_tensors[i], p_random_fft[i][col_num]) for i in range(n_dt))
        ev = verify_column_value(fid, column, outer_tensor, p_eval_fft[col_num])
        path = verify_column_path(fid, column, col_num, root)
        if not rand:
            raise VerifierError("ColumnDegree")
        if not ev:
            raise VerifierError("ColumnEval")
        if not path:
            raise VerifierError("ColumnPath")
    # :977-981 sum_j inner[j] * p_eval[j]
    prod = fe_mul(fid, inner_tensor, proof.p_eval)
    acc = 0
    p = MODULUS[fid]
    for v in from_limbs(prod):
        acc = (acc + v) % p
    return to_limbs(fid, [acc])


# ----------------------------------------------------------------------------- synthetic inputs

def splitmix64_stream(seed: int, n: int) -> np.ndarray:
    """n draws of splitmix64 from `seed` (SURVEY.md section 8 d: the bench/test input generator)."""
    mask = (1 << 64) - 1
    out = np.empty(n, dtype=np.uint64)
    idx = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed & mask))
    z = idx
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    out[:] = z ^ (z >> np.uint64(31))
    return out


def random_field_elements(fid: int, seed: int, n: int) -> np.ndarray:
    """Seeded field elements mirroring ff_derive's `random`: draw LIMBS words, mask the top limb to
    NUM_BITS bits, and (instead of rejecting) reduce the rare values >= p by one subtraction so the
    generator stays vectorisable.  Values are used directly as Montgomery residues."""
    L = LIMBS[fid]
    with np.errstate(over="ignore"):
        raw = splitmix64_stream(seed, n * L).reshape(n, L).copy()
    top_bits = NUM_BITS[fid] - 64 * (L - 1)
    raw[:, L - 1] &= np.uint64((1 << top_bits) - 1)
    p = MODULUS[fid]
    plimbs = [(p >> (64 * l)) & ((1 << 64) - 1) for l in range(L)]
    # find rows >= p (compare from the top limb down)
    ge = np.ones(n, dtype=bool)
    decided = np.zeros(n, dtype=bool)
    for l in range(L - 1, -1, -1):
        gt = raw[:, l] > np.uint64(plimbs[l])
        lt = raw[:, l] < np.uint64(plimbs[l])
        ge = np.where(~decided & lt, False, ge)
        decided |= gt | lt
    for i in np.nonzero(ge)[0]:
        v = sum(int(raw[i, l]) << (64 * l) for l in range(L)) - p
        for l in range(L):
            raw[i, l] = (v >> (64 * l)) & ((1 << 64) - 1)
    return raw


def pos_choose_columns(seed: int, amount: int, max_index: int) -> List[int]:
    """proof-of-storage get_column_indicies_from_random_seed (networking/client.rs:443-456):
    ChaCha8Rng::seed_from_u64 + IteratorRandom::choose_multiple over 0..max_index."""
    out = np.zeros(max(1, amount), dtype=np.uint64)
    lib().orc_pos_choose_columns.restype = C.c_size_t
    n = lib().orc_pos_choose_columns(C.c_uint64(seed), C.c_size_t(amount), C.c_size_t(max_index), _p64(out))
    return [int(x) for x in out[:n]]


def pack_bytes7(data: bytes) -> np.ndarray:
    """DataField::from_byte_vec for WriteableFt63 (fields/data_field.rs:38-46, writable_ft63.rs:35-40)."""
    n = (len(data) + 6) // 7
    buf = np.zeros(n * 7, dtype=np.uint8)
    buf[:len(data)] = np.frombuffer(data, dtype=np.uint8)
    b = buf.reshape(n, 7).astype(np.uint64)
    out = np.zeros(n, dtype=np.uint64)
    for k in range(7):
        out |= b[:, k] << np.uint64(8 * k)
    return out.reshape(n, 1)


def pack_bytes31(data: bytes) -> np.ndarray:
    """DataField::from_byte_vec for Ft253_192 (fields/data_field.rs:38-46, ft253_192.rs:18-30): each 31-byte group is
    zero-padded to 32 and limb i = u64::from_be_bytes(group[8i : 8i+8]); the limbs are stored as they are (Montgomery
    limbs).  Raises ValueError when a group is not below the modulus: the reference goes on computing with unreduced
    limbs there, which is not a field computation (ff_derive's add drops the carry out of 2^256)."""
    n = (len(data) + 30) // 31
    buf = np.zeros((n, 32), dtype=np.uint8)
    flat = np.zeros(n * 31, dtype=np.uint8)
    flat[:len(data)] = np.frombuffer(data, dtype=np.uint8)
    buf[:, :31] = flat.reshape(n, 31)
    out = buf.reshape(n, 4, 8)[:, :, ::-1].copy().view(np.uint64).reshape(n, 4)  # big-endian bytes -> u64 per limb
    p = MODULUS[FT253_192]
    for v in from_limbs(out):
        if v >= p:
            raise ValueError("Ft253_192::from_data_bytes: group not below the modulus")
    return out


# ---- pieces of hash_columns for a row-sharded commit (test double of lcpc_dev_hash_chunk_range / lcpc_dev_hash_merge) ----

def leaf_chunks(fid: int, n_rows: int) -> int:
    """Number of 1024-byte BLAKE3 chunks of a leaf: 32 zero bytes + n_rows reprs (lib.rs:749-764)."""
    return (32 + n_rows * 8 * LIMBS[fid] + 1023) // 1024


def hash_chunk_cvs(fid: int, rows: np.ndarray, row_base: int, n_rows_total: int, chunk0: int, chunk_end: int) -> np.ndarray:
    """Chaining values of chunks [chunk0, chunk_end) of every column's leaf stream, computed from the row window
    `rows` ([n_local, n_cols, LIMBS], first row = global row `row_base`).  Returns [chunk_end - chunk0, n_cols, 32]."""
    w = 8 * LIMBS[fid]
    n_local, n_cols = rows.shape[0], rows.shape[1]
    total = 32 + n_rows_total * w
    out = np.zeros((chunk_end - chunk0, n_cols, 32), dtype=np.uint8)
    buf = (C.c_uint8 * 32)()
    for j in range(n_cols):
        col = fe_to_repr(fid, np.ascontiguousarray(rows[:, j]))  # bytes of the local rows of this column
        for c in range(chunk0, chunk_end):
            lo, hi = 1024 * c, min(1024 * (c + 1), total)
            piece = bytearray()
            if lo < 32:
                piece += bytes(32 - lo)
                lo = 32
            a, b = lo - 32 - row_base * w, hi - 32 - row_base * w  # offsets into `col`
            assert 0 <= a <= b <= n_local * w, "chunk range reaches outside the row window"
            piece += col[a:b]
            lib().orc_b3_chunk_cv(bytes(piece), C.c_size_t(len(piece)), C.c_uint64(c), buf)
            out[c - chunk0, j] = np.frombuffer(bytes(buf), dtype=np.uint8)
    return out


def hash_merge(cvs: np.ndarray) -> np.ndarray:
    """Leaves from chunk chaining values [n_chunks >= 2, n_cols, 32] -> [n_cols, 32]."""
    n_chunks, n_cols = cvs.shape[0], cvs.shape[1]
    out = np.zeros((n_cols, 32), dtype=np.uint8)
    buf = (C.c_uint8 * 32)()
    for j in range(n_cols):
        col = np.ascontiguousarray(cvs[:, j])
        lib().orc_b3_merge_cvs(col.ctypes.data_as(u8p), C.c_size_t(n_chunks), buf)
        out[j] = np.frombuffer(bytes(buf), dtype=np.uint8)
    return out
