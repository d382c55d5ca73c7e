"""Host-side mirror of lcpc-2d's commitment interface over the lcpc_b200 C ABI.

Names, argument meaning and error behaviour follow the reference
(lcpc-2d/src/lib.rs; lcpc-ligero-pc/src/lib.rs; lcpc-brakedown-pc/src/lib.rs) so that
the parity tests read like the reference's own tests.  Field elements are numpy
``uint64`` arrays of shape ``(n, LIMBS)`` holding the Montgomery residue, least-
significant limb first -- the memory image of a Rust ``&[F]``.

Everything numerical happens on the GPU behind the C ABI; this module only moves
buffers and mirrors the reference's dimension logic.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import LcpcCsc, LcpcError, check, u64p

FT63, FT127, FT191, FT255, FT253_192 = 0, 1, 2, 3, 4
FIELD_NAMES = {FT63: "Ft63", FT127: "Ft127", FT191: "Ft191", FT255: "Ft255", FT253_192: "Ft253_192"}
FIELD_LIMBS = {FT63: 1, FT127: 2, FT191: 3, FT255: 4, FT253_192: 4}
FIELD_NUM_BITS = {FT63: 63, FT127: 127, FT191: 191, FT255: 255, FT253_192: 253}
FIELD_TWO_ADICITY = {FT63: 41, FT127: 40, FT191: 41, FT255: 41, FT253_192: 192}
# PrimeFieldReprEndianness: "big" for Ft253_192 only (proof-of-storage/src/fields/ft253_192.rs:9)
FIELD_REPR_BIG_ENDIAN = {FT63: False, FT127: False, FT191: False, FT255: False, FT253_192: True}


class ProverError(Exception):
    """lcpc_2d::ProverError (lib.rs:113-132); `.variant` is the Rust variant name."""

    def __init__(self, variant: str, message: str = ""):
        self.variant = variant
        super().__init__(f"{variant}: {message}" if message else variant)


class VerifierError(Exception):
    """lcpc_2d::VerifierError (lib.rs:139-167)."""

    def __init__(self, variant: str, message: str = ""):
        self.variant = variant
        super().__init__(f"{variant}: {message}" if message else variant)


def _prover_call(status: int) -> None:
    try:
        check(status)
    except LcpcError as e:
        if e.variant in ("TooBig", "Encode", "Commit", "ColumnNumber", "OuterTensor"):
            raise ProverError(e.variant, str(e)) from None
        raise


def next_pow2(v: int) -> int:
    return 1 if v <= 1 else 1 << (v - 1).bit_length()


def log2(v: int) -> int:
    """lib.rs:857-859."""
    return next_pow2(v).bit_length() - 1


def n_degree_tests(lam: int, length: int, flog2: int) -> int:
    """lib.rs:642-645."""
    den = flog2 - log2(length)
    return (lam + den - 1) // den


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _elems(a, limbs: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a.reshape(-1, limbs)


class Context:
    """One device + stream (lcpc_ctx)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        lib = _lib.load()
        h = C.c_void_p()
        if stream is None:
            check(lib.lcpc_ctx_create(device, C.byref(h)))
        else:
            check(lib.lcpc_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h)))
        self._h = h
        self.device = device

    @classmethod
    def multi(cls, devices: Sequence[int]) -> "Context":
        """lcpc_ctx_create_multi: one context over several devices of this process; commitments made through it are
        sharded inside the library (rows for encoding, column blocks of the leaf range for the Merkle subtrees)."""
        lib = _lib.load()
        h = C.c_void_p()
        arr = (C.c_int32 * len(devices))(*devices)
        check(lib.lcpc_ctx_create_multi(arr, len(devices), C.byref(h)))
        self = cls.__new__(cls)
        self._h = h
        self.device = devices[0]
        return self

    @property
    def n_devices(self) -> int:
        return int(_lib.load().lcpc_ctx_device_count(self._h))

    @property
    def handle(self):
        return self._h

    def synchronize(self) -> None:
        check(_lib.load().lcpc_ctx_synchronize(self._h))

    @property
    def stream(self) -> int:
        """The cudaStream_t (as an integer) every call on this context enqueues on."""
        out = C.c_void_p()
        check(_lib.load().lcpc_ctx_stream(self._h, C.byref(out)))
        return int(out.value or 0)

    def measure_int_pipes(self) -> dict:
        """Integer-pipe issue rates of this device (lcpc_ctx_measure_int_pipes): SM sub-partition cycles per warp
        instruction for IMAD.WIDE / IMAD / LOP3 / an IMAD + LOP3 mix, and the SM clock seen during the measurement."""
        cyc = (C.c_double * 4)()
        ghz = C.c_double()
        check(_lib.load().lcpc_ctx_measure_int_pipes(self._h, cyc, C.byref(ghz)))
        return {"cycles_per_warp_instr": {"imad_wide": cyc[0], "imad": cyc[1], "lop3": cyc[2], "imad_lop3_mix": cyc[3]},
                "sm_ghz": ghz.value}

    def launch_count(self) -> int:
        return int(_lib.load().lcpc_ctx_launch_count(self._h))

    def kernel_timing(self, enable: bool) -> None:
        check(_lib.load().lcpc_ctx_kernel_timing(self._h, 1 if enable else 0))

    def kernel_timing_report(self) -> dict:
        """{kernel name: (launches, total_ms)} since the last report."""
        txt = _lib.load().lcpc_ctx_kernel_timing_report(self._h).decode()
        out = {}
        for line in txt.splitlines():
            name, n, ms = line.split()
            out[name] = (int(n), float(ms))
        return out

    def close(self) -> None:
        if self._h:
            _lib.load().lcpc_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class Transcript:
    """merlin::Transcript (host-side; lcpc_transcript_* of the C ABI)."""

    def __init__(self, label: bytes, _handle=None):
        if _handle is not None:
            self._h = _handle
            return
        h = C.c_void_p()
        check(_lib.load().lcpc_transcript_new(label, len(label), C.byref(h)))
        self._h = h

    def append_message(self, label: bytes, message: bytes) -> None:
        check(_lib.load().lcpc_transcript_append_message(self._h, label, len(label), message, len(message)))

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        out = (C.c_uint8 * n)()
        check(_lib.load().lcpc_transcript_challenge_bytes(self._h, label, len(label), out, n))
        return bytes(out)

    def clone(self) -> "Transcript":
        h = C.c_void_p()
        check(_lib.load().lcpc_transcript_clone(self._h, C.byref(h)))
        return Transcript(b"", _handle=h)

    def __del__(self):
        try:
            if self._h:
                _lib.load().lcpc_transcript_free(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


class _Encoding:
    """Common part of the `LcEncoding` implementations (lib.rs:75-105)."""

    fid: int
    n_per_row: int
    n_cols: int
    _plan: C.c_void_p
    ctx: Context

    # def_labels! (macros.rs:28-36): `$l` is not substituted inside a byte string
    LABEL_DT = b"$l//DT"
    LABEL_PR = b"$l//PR"
    LABEL_PE = b"$l//PE"
    LABEL_CO = b"$l//CO"

    @property
    def limbs(self) -> int:
        return FIELD_LIMBS[self.fid]

    @property
    def plan(self):
        return self._plan

    def get_dims(self, length: int) -> Tuple[int, int, int]:
        return ((length + self.n_per_row - 1) // self.n_per_row, self.n_per_row, self.n_cols)

    def encode(self, inp: np.ndarray) -> None:
        """LcEncoding::encode: one row (or a batch of rows) of n_cols elements, in place."""
        assert inp.dtype == np.uint64 and inp.flags["C_CONTIGUOUS"]
        rows = inp.size // (self.n_cols * self.limbs)
        assert rows * self.n_cols * self.limbs == inp.size, "row length must be n_cols"
        _prover_call(_lib.load().lcpc_encode_rows(self._plan, _ptr(inp), rows))

    def decode(self, inp: np.ndarray) -> None:
        """fffft ifft_oi on one row (or a batch of rows) of n_cols encoded elements, in place: decode(encode(x)) == x
        (proof-of-storage decode_row, lcpc_online.rs:568-573).  Ligero encodings only."""
        assert inp.dtype == np.uint64 and inp.flags["C_CONTIGUOUS"]
        rows = inp.size // (self.n_cols * self.limbs)
        assert rows * self.n_cols * self.limbs == inp.size, "row length must be n_cols"
        _prover_call(_lib.load().lcpc_decode_rows(self._plan, _ptr(inp), rows))

    def close(self) -> None:
        if getattr(self, "_plan", None):
            _lib.load().lcpc_plan_destroy(self._plan)
            self._plan = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LigeroEncoding(_Encoding):
    """LigeroEncodingRho<Ft, Rn, Rd> (lcpc-ligero-pc/src/lib.rs:31-186); default rho = 1/2 (:189)."""

    LAMBDA = 128

    def __init__(self, fid: int, n_per_row: int, n_cols: int, rho_num: int = 1, rho_den: int = 2,
                 ctx: Optional[Context] = None, root_of_unity: Optional[np.ndarray] = None):
        """new_from_dims (:138-148)."""
        if not self._dims_ok(n_per_row, n_cols):
            raise AssertionError("assertion failed: Self::_dims_ok(n_per_row, n_cols)")
        self.fid, self.n_per_row, self.n_cols = fid, n_per_row, n_cols
        self.rho_num, self.rho_den = rho_num, rho_den
        self.ctx = ctx or default_context()
        h = C.c_void_p()
        root = None
        if root_of_unity is not None:
            root = np.ascontiguousarray(root_of_unity, dtype=np.uint64).ctypes.data_as(u64p)
        check(_lib.load().lcpc_plan_ligero(self.ctx.handle, fid, n_per_row, n_cols, root, C.byref(h)))
        self._plan = h

    @classmethod
    def _n_col_opens(cls, rho_num: int, rho_den: int) -> int:  # :61-64
        rho = rho_num / rho_den
        return int(math.ceil(-cls.LAMBDA / math.log2((1.0 + rho) / 2.0)))

    @classmethod
    def _n_degree_tests(cls, fid: int, n_cols: int) -> int:  # :66-68
        return n_degree_tests(cls.LAMBDA, n_cols, FIELD_NUM_BITS[fid] - 1)

    @classmethod
    def _get_dims(cls, fid: int, length: int, rho_num: int = 1, rho_den: int = 2) -> Optional[Tuple[int, int, int]]:
        """:70-112."""
        rho = rho_num / rho_den
        n_col_opens = cls._n_col_opens(rho_num, rho_den)
        lncf = float(n_col_opens * length)
        ndt = float(cls._n_degree_tests(fid, int(math.ceil(math.sqrt(lncf) / rho))))
        nc1 = next_pow2(int(math.ceil(math.sqrt(lncf / ndt) / rho)))
        if nc1 > (1 << FIELD_TWO_ADICITY[fid]):
            return None
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
    def new(cls, fid: int, length: int, rho_num: int = 1, rho_den: int = 2,
            ctx: Optional[Context] = None) -> "LigeroEncoding":
        """:121-124."""
        dims = cls._get_dims(fid, length, rho_num, rho_den)
        if dims is None:
            raise ValueError("called `Option::unwrap()` on a `None` value")
        return cls(fid, dims[1], dims[2], rho_num, rho_den, ctx)

    @classmethod
    def new_ml(cls, fid: int, n_vars: int, ctx: Optional[Context] = None) -> "LigeroEncoding":
        """:128-135."""
        n_monomials = 1 << n_vars
        n_rows, n_per_row, n_cols = cls._get_dims(fid, n_monomials)
        assert n_rows & (n_rows - 1) == 0 and n_per_row & (n_per_row - 1) == 0
        assert n_rows * n_per_row == n_monomials
        return cls(fid, n_per_row, n_cols, ctx=ctx)

    @staticmethod
    def _dims_ok(n_per_row: int, n_cols: int) -> bool:  # :114-118
        return n_per_row < n_cols and n_cols > 0 and n_cols & (n_cols - 1) == 0

    def dims_ok(self, n_per_row: int, n_cols: int) -> bool:  # :171-177
        return self._dims_ok(n_per_row, n_cols) and n_per_row == self.n_per_row and n_cols == self.n_cols

    def get_n_col_opens(self) -> int:
        return self._n_col_opens(self.rho_num, self.rho_den)

    def get_n_degree_tests(self) -> int:
        return self._n_degree_tests(self.fid, self.n_cols)


@dataclass
class CscMatrix:
    """sprs::CsMat<F> in CSC storage (rows x cols), host arrays."""
    rows: int
    cols: int
    indptr: np.ndarray   # (cols+1,) uint64
    indices: np.ndarray  # (nnz,) uint64
    data: np.ndarray     # (nnz, LIMBS) uint64, Montgomery

    def as_struct(self) -> LcpcCsc:
        for a in (self.indptr, self.indices, self.data):
            assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
        return LcpcCsc(self.rows, self.cols, self.indptr.ctypes.data_as(u64p),
                       self.indices.ctypes.data_as(u64p), self.data.ctypes.data_as(u64p))


def sdig_codeword_length(pre: Sequence[CscMatrix], post: Sequence[CscMatrix]) -> int:
    """encode.rs:18-33."""
    return pre[0].cols + post[-1].cols + sum(m.rows for m in pre[:-1]) + sum(m.rows for m in post)


class SdigEncoding(_Encoding):
    """SdigEncodingS<Ft, S> (lcpc-brakedown-pc/src/lib.rs:38-176) built from generated code matrices
    (matgen.rs:28-53; generation is host-side and stays with the caller)."""

    LAMBDA = 128

    def __init__(self, fid: int, precodes: Sequence[CscMatrix], postcodes: Sequence[CscMatrix],
                 n_col_opens: int = 6593, ctx: Optional[Context] = None):
        assert len(precodes) == len(postcodes) and len(precodes) > 0
        self.fid = fid
        self.precodes, self.postcodes = list(precodes), list(postcodes)
        self.n_per_row = precodes[0].cols
        self.n_cols = sdig_codeword_length(precodes, postcodes)
        self._n_col_opens_value = n_col_opens
        self.ctx = ctx or default_context()
        n = len(precodes)
        pre = (LcpcCsc * n)(*[m.as_struct() for m in precodes])
        post = (LcpcCsc * n)(*[m.as_struct() for m in postcodes])
        h = C.c_void_p()
        check(_lib.load().lcpc_plan_brakedown(self.ctx.handle, fid, self.n_per_row, self.n_cols, n, pre, post, C.byref(h)))
        self._plan = h

    # -- construction from a seed, as the reference does (host-side generation) ---------------
    @staticmethod
    def generate(fid: int, n_per_row: int, seed: int, code: int = 3):
        """matgen::generate::<Ft, SdigCode<code>>(n_per_row, seed) (matgen.rs:28-53)."""
        lib = _lib.load()
        L = FIELD_LIMBS[fid]
        pre_d = np.zeros(3 * 64, dtype=np.uint64)
        post_d = np.zeros(3 * 64, dtype=np.uint64)
        n_levels = C.c_int32()
        check(lib.lcpc_sdig_get_dims(code, n_per_row, fid, pre_d.ctypes.data_as(u64p), post_d.ctypes.data_as(u64p), 64,
                                     C.byref(n_levels)))
        pres, posts = [], []
        for lvl in range(n_levels.value):
            ni, mi, cn = (int(x) for x in pre_d[3 * lvl:3 * lvl + 3])
            nip, mip, dn = (int(x) for x in post_d[3 * lvl:3 * lvl + 3])
            a = CscMatrix(mi, ni, np.zeros(ni + 1, np.uint64), np.zeros(ni * cn, np.uint64), np.zeros((ni * cn, L), np.uint64))
            b = CscMatrix(mip, nip, np.zeros(nip + 1, np.uint64), np.zeros(nip * dn, np.uint64), np.zeros((nip * dn, L), np.uint64))
            pd = np.array([ni, mi, cn], dtype=np.uint64)
            qd = np.array([nip, mip, dn], dtype=np.uint64)
            p64 = lambda x: x.ctypes.data_as(u64p)
            check(lib.lcpc_sdig_gen_level(fid, seed, lvl, p64(pd), p64(qd), p64(a.indptr), p64(a.indices), p64(a.data),
                                          p64(b.indptr), p64(b.indices), p64(b.data)))
            pres.append(a)
            posts.append(b)
        return pres, posts

    @classmethod
    def _n_col_opens(cls, code: int = 3) -> int:  # lib.rs:57-61
        dist = float(_lib.load().lcpc_sdig_dist(code))
        return int(math.ceil(-cls.LAMBDA / math.log2(1.0 - dist / 3.0)))

    @classmethod
    def _n_degree_tests(cls, fid: int, n_cols: int) -> int:  # lib.rs:64-66
        return n_degree_tests(cls.LAMBDA, n_cols, FIELD_NUM_BITS[fid] - 1)

    @classmethod
    def _n_per_row_for_len(cls, fid: int, length: int, code: int = 3, ml: bool = False) -> int:
        """lib.rs:69-123: new / new_ml / _new_from_np1."""
        n_col_opens = cls._n_col_opens(code)
        lncf = float(n_col_opens * length)
        ndt = float(cls._n_degree_tests(fid, int(math.ceil(math.sqrt(lncf))) * 2))
        np1 = int(math.ceil(math.sqrt(lncf / ndt)))
        if ml:
            np1 = next_pow2(np1)
        np1 = min(np1, length)
        nr1 = (length + np1 - 1) // np1
        nd1 = cls._n_degree_tests(fid, np1 * 2)
        np2 = np1 // 2
        nr2 = (length + np2 - 1) // np2
        nd2 = cls._n_degree_tests(fid, np2 * 2)
        sz1 = n_col_opens * nr1 + (1 + nd1) * np1
        sz2 = n_col_opens * nr2 + (1 + nd2) * np2
        return np1 if sz1 < sz2 else np2

    @classmethod
    def new(cls, fid: int, length: int, seed: int, code: int = 3, ctx: Optional[Context] = None) -> "SdigEncoding":
        """SdigEncodingS::new(len, seed) (lib.rs:93-100)."""
        return cls.new_from_dims(fid, cls._n_per_row_for_len(fid, length, code), None, seed, code, ctx)

    @classmethod
    def new_ml(cls, fid: int, n_vars: int, seed: int, code: int = 3, ctx: Optional[Context] = None) -> "SdigEncoding":
        """lib.rs:104-113."""
        return cls.new_from_dims(fid, cls._n_per_row_for_len(fid, 1 << n_vars, code, ml=True), None, seed, code, ctx)

    @classmethod
    def new_from_dims(cls, fid: int, n_per_row: int, n_cols: Optional[int], seed: int, code: int = 3,
                      ctx: Optional[Context] = None) -> "SdigEncoding":
        """lib.rs:126-137."""
        pre, post = cls.generate(fid, n_per_row, seed, code)
        enc = cls(fid, pre, post, n_col_opens=cls._n_col_opens(code), ctx=ctx)
        if n_cols is not None and n_cols != enc.n_cols:
            raise AssertionError("assertion failed: n_cols == codeword_length(&precodes, &postcodes)")
        return enc

    def dims_ok(self, n_per_row: int, n_cols: int) -> bool:  # :160-167
        return n_per_row < n_cols and n_per_row == self.n_per_row and n_cols == self.n_cols

    def get_n_col_opens(self) -> int:
        return self._n_col_opens_value

    def get_n_degree_tests(self) -> int:
        return n_degree_tests(self.LAMBDA, self.n_cols, FIELD_NUM_BITS[self.fid] - 1)


@dataclass
class LcColumn:
    """lib.rs:422-439."""
    col: np.ndarray   # (n_rows, LIMBS)
    path: np.ndarray  # (log2(n_cols), 32) uint8


@dataclass
class LcEvalProof:
    """lib.rs:516-529."""
    n_cols: int
    p_eval: np.ndarray
    p_random_vec: List[np.ndarray]
    columns: List[LcColumn]

    def get_n_cols(self) -> int:
        return self.n_cols

    def get_n_per_row(self) -> int:
        return self.p_eval.shape[0]

    def verify(self, root: bytes, outer_tensor: np.ndarray, inner_tensor: np.ndarray, enc: "_Encoding",
               tr: "Transcript") -> np.ndarray:
        """LcEvalProof::verify (lib.rs:547-556 -> :862-982): returns the evaluation (1, LIMBS) or raises
        VerifierError with the reference's variant name."""
        return verify(root, outer_tensor, inner_tensor, self, enc, tr)


class LcCommit:
    """lib.rs:174-191.  `comm`, `coeffs`, `hashes` are host arrays like the Rust struct's Vecs; they
    are fetched from the device-resident handle on first access and cached."""

    def __init__(self, enc: _Encoding, handle: C.c_void_p, n_rows: int, comm=None, coeffs=None, hashes=None):
        self.enc = enc
        self.fid = enc.fid
        self._h = handle
        self.n_rows, self.n_per_row, self.n_cols = n_rows, enc.n_per_row, enc.n_cols
        self._comm, self._coeffs, self._hashes = comm, coeffs, hashes

    # -- construction ------------------------------------------------------------------
    @classmethod
    def commit(cls, coeffs_in: np.ndarray, enc: _Encoding, download: bool = True) -> "LcCommit":
        """LcCommit::commit (lib.rs:314 -> :651-700)."""
        L = enc.limbs
        coeffs_in = _elems(coeffs_in, L)
        n = coeffs_in.shape[0]
        n_rows, n_per_row, n_cols = enc.get_dims(n)
        if n == 0:
            raise AssertionError("assertion failed: (n_rows - 1) * n_per_row < coeffs_in.len()")
        np2 = next_pow2(n_cols)
        comm = coeffs = hashes = None
        if download:
            coeffs = np.empty((n_rows, n_per_row, L), dtype=np.uint64)
            comm = np.empty((n_rows, n_cols, L), dtype=np.uint64)
            hashes = np.empty((2 * np2 - 1, 32), dtype=np.uint8)
        h = C.c_void_p()
        _prover_call(_lib.load().lcpc_commit_host(enc.plan, _ptr(coeffs_in), n, _ptr(coeffs), _ptr(comm), _ptr(hashes),
                                                  C.byref(h)))
        return cls(enc, h, n_rows, comm, coeffs, hashes)

    @classmethod
    def commit_bytes(cls, data: bytes, enc: _Encoding, download: bool = True) -> "LcCommit":
        """proof-of-storage: DataField::from_byte_vec + commit (lcpc_online.rs:81-143).  WriteableFt63 packs 7 bytes per
        element (writable_ft63.rs:35-40), Ft253_192 packs 31 (ft253_192.rs:18-30; groups that are not below the modulus
        are refused, see include/lcpc_b200.h)."""
        buf = np.frombuffer(data, dtype=np.uint8)
        per = 31 if enc.fid == FT253_192 else 7
        L = FIELD_LIMBS[enc.fid]
        n = (len(data) + per - 1) // per
        if n == 0:
            raise ValueError("Cannot convert empty file to commit")  # lcpc_online.rs:91
        n_rows, n_per_row, n_cols = enc.get_dims(n)
        np2 = next_pow2(n_cols)
        comm = coeffs = hashes = None
        if download:
            coeffs = np.empty((n_rows, n_per_row, L), dtype=np.uint64)
            comm = np.empty((n_rows, n_cols, L), dtype=np.uint64)
            hashes = np.empty((2 * np2 - 1, 32), dtype=np.uint8)
        h = C.c_void_p()
        _prover_call(_lib.load().lcpc_commit_bytes_host(enc.plan, _ptr(buf), len(data), _ptr(coeffs), _ptr(comm),
                                                        _ptr(hashes), C.byref(h)))
        return cls(enc, h, n_rows, comm, coeffs, hashes)

    # -- the Rust struct's public fields ------------------------------------------------
    def _download(self) -> None:
        L = self.enc.limbs
        np2 = next_pow2(self.n_cols)
        self._coeffs = np.empty((self.n_rows, self.n_per_row, L), dtype=np.uint64)
        self._comm = np.empty((self.n_rows, self.n_cols, L), dtype=np.uint64)
        self._hashes = np.empty((2 * np2 - 1, 32), dtype=np.uint8)
        check(_lib.load().lcpc_commit_download(self._h, _ptr(self._coeffs), _ptr(self._comm), _ptr(self._hashes)))

    @property
    def comm(self) -> np.ndarray:
        if self._comm is None:
            self._download()
        return self._comm

    @property
    def coeffs(self) -> np.ndarray:
        if self._coeffs is None:
            self._download()
        return self._coeffs

    @property
    def hashes(self) -> np.ndarray:
        if self._hashes is None:
            self._download()
        return self._hashes

    def get_root(self) -> bytes:
        """lib.rs:291-296."""
        out = np.empty(32, dtype=np.uint8)
        check(_lib.load().lcpc_commit_root(self._h, _ptr(out)))
        return out.tobytes()

    def update_rows(self, row0: int, coeff_rows: np.ndarray) -> np.ndarray:
        """Replace coefficient rows [row0, row0 + k) and bring the commitment up to date: only those rows are
        re-encoded and only the BLAKE3 chunks of the column leaves that contain them are re-hashed
        (proof-of-storage FileHandler::edit_bytes -> reencode_row -> recalculate_merkle_tree,
        lcpc_online/file_handler.rs:279-402, 474-481).  Returns the new flat tree."""
        L = self.enc.limbs
        rows = np.ascontiguousarray(coeff_rows, dtype=np.uint64).reshape(-1, self.n_per_row, L)
        k = rows.shape[0]
        np2 = next_pow2(self.n_cols)
        hashes = np.empty((2 * np2 - 1, 32), dtype=np.uint8)
        comm_rows = np.empty((k, self.n_cols, L), dtype=np.uint64)
        _prover_call(_lib.load().lcpc_commit_update_rows_host(self._h, row0, k, _ptr(rows), _ptr(comm_rows), _ptr(hashes)))
        if self._coeffs is not None:
            self._coeffs[row0:row0 + k] = rows
        if self._comm is not None:
            self._comm[row0:row0 + k] = comm_rows
        self._hashes = hashes
        return hashes

    def append_rows(self, row0: int, coeff_rows: np.ndarray) -> np.ndarray:
        """Write rows [row0, row0 + k) with row0 <= n_rows <= row0 + k: the last row may be replaced and new rows follow
        (proof-of-storage FileHandler::append_bytes, lcpc_online/file_handler.rs:336-402).  Returns the new flat tree."""
        L = self.enc.limbs
        rows = np.ascontiguousarray(coeff_rows, dtype=np.uint64).reshape(-1, self.n_per_row, L)
        k = rows.shape[0]
        hashes = np.empty((2 * next_pow2(self.n_cols) - 1, 32), dtype=np.uint8)
        _prover_call(_lib.load().lcpc_commit_append_rows_host(self._h, row0, k, _ptr(rows), None, _ptr(hashes)))
        self.n_rows = max(self.n_rows, row0 + k)
        self._coeffs = self._comm = None  # re-read lazily from the (grown) device matrices
        self._hashes = hashes
        return hashes

    # -- folds / openings ----------------------------------------------------------------
    def fold(self, tensors: np.ndarray, encoded: bool = False) -> np.ndarray:
        """collapse_columns for a batch of tensors: (n_tensors, n_rows, L) -> (n_tensors, width, L)."""
        L = self.enc.limbs
        tensors = np.ascontiguousarray(tensors, dtype=np.uint64).reshape(-1, self.n_rows, L)
        width = self.n_cols if encoded else self.n_per_row
        out = np.empty((tensors.shape[0], width, L), dtype=np.uint64)
        _prover_call(_lib.load().lcpc_fold_host(self._h, 1 if encoded else 0, _ptr(tensors), tensors.shape[0], _ptr(out)))
        return out

    def prove(self, outer_tensor: np.ndarray, enc: "_Encoding", tr: "Transcript") -> LcEvalProof:
        """LcCommit::prove (lib.rs:319-326 -> :1034-1123)."""
        return prove(self, outer_tensor, enc, tr)

    def open_columns(self, cols: Sequence[int], with_path: bool = True) -> List[LcColumn]:
        L = self.enc.limbs
        idx = np.ascontiguousarray(np.asarray(cols, dtype=np.uint64))
        n = idx.shape[0]
        depth = log2(self.n_cols)
        out = np.empty((n, self.n_rows, L), dtype=np.uint64)
        paths = np.empty((n, depth, 32), dtype=np.uint8) if with_path else None
        _prover_call(_lib.load().lcpc_open_columns_host(self._h, _ptr(idx), n, _ptr(out), _ptr(paths)))
        return [LcColumn(out[i], paths[i] if with_path else np.empty((0, 32), np.uint8)) for i in range(n)]

    def leaves(self, cols: Sequence[int]) -> np.ndarray:
        """CommitRequestType::Leaves (lcpc_online.rs:144-190): leaf digests of selected columns."""
        idx = np.ascontiguousarray(np.asarray(cols, dtype=np.uint64))
        out = np.empty((idx.shape[0], 32), dtype=np.uint8)
        _prover_call(_lib.load().lcpc_leaves_host(self._h, _ptr(idx), idx.shape[0], _ptr(out)))
        return out

    def close(self) -> None:
        if self._h:
            _lib.load().lcpc_commit_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def commit(coeffs_in: np.ndarray, enc: _Encoding) -> LcCommit:
    """lib.rs:651."""
    return LcCommit.commit(coeffs_in, enc)


def open_column(comm: LcCommit, column: int) -> LcColumn:
    """lib.rs:818-855; ProverError::ColumnNumber for column >= n_cols."""
    if column < 0:
        raise ProverError("ColumnNumber", "bad column number")
    return comm.open_columns([column])[0]


def collapse_columns(comm: LcCommit, tensor: np.ndarray) -> np.ndarray:
    """lib.rs:1126-1154 over the unencoded coefficients."""
    t = _elems(tensor, comm.enc.limbs)
    if t.shape[0] != comm.n_rows:
        raise ProverError("OuterTensor", "bad outer tensor size")
    return comm.fold(t[None])[0]


def prove(comm: LcCommit, outer_tensor: np.ndarray, enc: _Encoding, tr: Transcript) -> LcEvalProof:
    """lib.rs:1034-1123.  check_comm (:1045) is structural here: the handle cannot be malformed."""
    L = enc.limbs
    if not enc.dims_ok(comm.n_per_row, comm.n_cols):
        raise ProverError("Commit", "bad dimensions for commitment")
    outer = _elems(outer_tensor, L)
    n_dt, n_open = enc.get_n_degree_tests(), enc.get_n_col_opens()
    depth = log2(comm.n_cols)
    p_eval = np.empty((comm.n_per_row, L), dtype=np.uint64)
    p_random = np.empty((n_dt, comm.n_per_row, L), dtype=np.uint64)
    cols = np.empty((n_open, comm.n_rows, L), dtype=np.uint64)
    paths = np.empty((n_open, depth, 32), dtype=np.uint8)
    _prover_call(_lib.load().lcpc_prove(comm._h, _ptr(outer), outer.shape[0], n_dt, n_open, tr._h, _ptr(p_eval),
                                        _ptr(p_random), None, _ptr(cols), _ptr(paths)))
    return LcEvalProof(comm.n_cols, p_eval, [p_random[i] for i in range(n_dt)],
                       [LcColumn(cols[i], paths[i]) for i in range(n_open)])


def verify(root: bytes, outer_tensor: np.ndarray, inner_tensor: np.ndarray, proof: LcEvalProof, enc: _Encoding,
           tr: Transcript) -> np.ndarray:
    """lib.rs:862-982."""
    from ._lib import VERIFIER_ERRORS

    L = enc.limbs
    outer, inner = _elems(outer_tensor, L), _elems(inner_tensor, L)
    n_columns = len(proof.columns)
    n_rows = proof.columns[0].col.shape[0] if n_columns else 0
    path_len = proof.columns[0].path.shape[0] if n_columns else 0
    for c in proof.columns:  # ragged proofs cannot be passed flat; the reference fails the affected column
        if c.col.shape[0] != n_rows:
            raise VerifierError("ColumnEval", "column eval invalid")
        if c.path.shape[0] != path_len:
            raise VerifierError("ColumnPath", "column path invalid")
    columns = np.ascontiguousarray(np.stack([c.col for c in proof.columns])) if n_columns else np.empty((0, 0, L), np.uint64)
    paths = np.ascontiguousarray(np.stack([c.path for c in proof.columns])) if n_columns else np.empty((0, 0, 32), np.uint8)
    p_eval = _elems(proof.p_eval, L)
    n_pr = len(proof.p_random_vec)
    # the C ABI takes the p_random vectors flat, n_per_row elements each.  The reference copies each vector into a
    # zero row of n_cols (lib.rs:913-916) and feeds ALL its elements to the transcript, so a vector of another length
    # changes the challenges and fails the degree test; report it as that instead of reading a misaligned buffer
    for v in proof.p_random_vec:
        if _elems(v, L).shape[0] != p_eval.shape[0]:
            raise VerifierError("ColumnDegree", "column degree check failed")
    p_random = (np.ascontiguousarray(np.stack([_elems(v, L) for v in proof.p_random_vec]))
                if n_pr else np.empty((0, p_eval.shape[0], L), np.uint64))
    rootb = np.frombuffer(root, dtype=np.uint8).copy()
    res = np.empty((1, L), dtype=np.uint64)
    rc = _lib.load().lcpc_verify(enc.plan, _ptr(rootb), _ptr(outer), outer.shape[0], _ptr(inner), inner.shape[0],
                                 proof.n_cols, _ptr(p_eval), p_eval.shape[0], _ptr(p_random), n_pr, _ptr(columns), n_rows,
                                 _ptr(paths), path_len, n_columns, enc.get_n_col_opens(), enc.get_n_degree_tests(),
                                 tr._h, _ptr(res))
    if rc in VERIFIER_ERRORS:
        raise VerifierError(VERIFIER_ERRORS[rc], _lib.load().lcpc_last_error().decode())
    check(rc)
    return res
