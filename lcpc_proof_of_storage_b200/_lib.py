"""ctypes binding of liblcpc_b200.so (the C ABI declared in include/lcpc_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc, and if that
fails, or no CUDA device is present when a context is created, the error propagates.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "lcpc_b200.h")

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
szp = C.POINTER(C.c_size_t)
vpp = C.POINTER(C.c_void_p)

LCPC_OK = 0
STATUS_NAMES = {
    0: "Ok", -1: "TooBig", -2: "Encode", -3: "Commit", -4: "ColumnNumber", -5: "OuterTensor",
    -6: "Dims", -7: "InvalidArg", -8: "Cuda", -9: "NoMem",
}
VERIFIER_ERRORS = {
    -20: "NumColOpens", -21: "ColumnPath", -22: "ColumnEval", -23: "ColumnDegree", -24: "OuterTensor",
    -25: "InnerTensor", -26: "EncodingDims", -27: "Encode",
}
STATUS_NAMES.update({k: "Verifier" + v for k, v in VERIFIER_ERRORS.items()})


class LcpcCsc(C.Structure):
    _fields_ = [("rows", C.c_uint64), ("cols", C.c_uint64), ("indptr", u64p), ("indices", u64p), ("data", u64p)]


class LcpcError(RuntimeError):
    """A non-zero lcpc_status; `.variant` names the ProverError-style variant."""

    def __init__(self, code: int, message: str):
        self.code = code
        self.variant = STATUS_NAMES.get(code, str(code))
        super().__init__(f"lcpc_b200: {self.variant}: {message}")


_SIGNATURES = {
    "lcpc_abi_version": (C.c_uint32, []),
    "lcpc_last_error": (C.c_char_p, []),
    "lcpc_field_limbs": (C.c_int32, [C.c_int32]),
    "lcpc_field_constants": (C.c_int32, [C.c_int32, u64p, u64p, u64p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "lcpc_ctx_create": (C.c_int32, [C.c_int32, vpp]),
    "lcpc_ctx_create_on_stream": (C.c_int32, [C.c_int32, C.c_void_p, vpp]),
    "lcpc_ctx_create_multi": (C.c_int32, [C.POINTER(C.c_int32), C.c_int32, vpp]),
    "lcpc_ctx_device_count": (C.c_int32, [C.c_void_p]),
    "lcpc_ctx_synchronize": (C.c_int32, [C.c_void_p]),
    "lcpc_ctx_stream": (C.c_int32, [C.c_void_p, vpp]),
    "lcpc_ctx_measure_int_pipes": (C.c_int32, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "lcpc_ctx_destroy": (None, [C.c_void_p]),
    "lcpc_ctx_launch_count": (C.c_uint64, [C.c_void_p]),
    "lcpc_ctx_kernel_timing": (C.c_int32, [C.c_void_p, C.c_int32]),
    "lcpc_ctx_kernel_timing_report": (C.c_char_p, [C.c_void_p]),
    "lcpc_stream_begin": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, vpp]),
    "lcpc_stream_push_elems_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "lcpc_stream_push_bytes_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "lcpc_stream_finish": (C.c_int32, [C.c_void_p, C.c_void_p, szp]),
    "lcpc_stream_free": (None, [C.c_void_p]),
    "lcpc_commit_update_rows_host": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcpc_commit_append_rows_host": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcpc_plan_ligero": (C.c_int32, [C.c_void_p, C.c_int32, C.c_size_t, C.c_size_t, u64p, vpp]),
    "lcpc_plan_brakedown": (C.c_int32, [C.c_void_p, C.c_int32, C.c_size_t, C.c_size_t, C.c_size_t,
                                        C.POINTER(LcpcCsc), C.POINTER(LcpcCsc), vpp]),
    "lcpc_plan_get_dims": (C.c_int32, [C.c_void_p, C.c_size_t, szp, szp, szp]),
    "lcpc_plan_destroy": (None, [C.c_void_p]),
    "lcpc_encode_rows": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "lcpc_decode_rows": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "lcpc_dev_decode": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "lcpc_commit_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, vpp]),
    "lcpc_commit_bytes_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, vpp]),
    "lcpc_commit_dev": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, vpp]),
    "lcpc_commit_get_dims": (C.c_int32, [C.c_void_p, szp, szp, szp]),
    "lcpc_commit_root": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "lcpc_commit_download": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcpc_commit_device_ptrs": (C.c_int32, [C.c_void_p, vpp, vpp, vpp]),
    "lcpc_commit_free": (None, [C.c_void_p]),
    "lcpc_fold_host": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lcpc_open_columns_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "lcpc_leaves_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lcpc_transcript_new": (C.c_int32, [C.c_char_p, C.c_size_t, vpp]),
    "lcpc_transcript_clone": (C.c_int32, [C.c_void_p, vpp]),
    "lcpc_transcript_append_message": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]),
    "lcpc_transcript_challenge_bytes": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "lcpc_transcript_free": (None, [C.c_void_p]),
    "lcpc_random_field_vec": (C.c_int32, [C.c_int32, C.c_char_p, C.c_void_p, C.c_size_t]),
    "lcpc_random_columns": (C.c_int32, [C.c_char_p, C.c_uint64, C.c_void_p, C.c_size_t]),
    "lcpc_prove": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcpc_verify": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t,
                                C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]),
    "lcpc_pos_choose_columns": (C.c_int32, [C.c_uint64, C.c_size_t, C.c_size_t, C.c_void_p, szp]),
    "lcpc_verify_columns_host": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                             C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcpc_sdig_get_dims": (C.c_int32, [C.c_int32, C.c_uint64, C.c_int32, u64p, u64p, C.c_int32, C.POINTER(C.c_int32)]),
    "lcpc_sdig_gen_level": (C.c_int32, [C.c_int32, C.c_uint64, C.c_uint64, u64p, u64p, u64p, u64p, u64p, u64p, u64p, u64p]),
    "lcpc_sdig_dist": (C.c_double, [C.c_int32]),
    "lcpc_dev_encode": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lcpc_dev_encode_scatter": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p,
                                            C.POINTER(C.c_void_p), C.c_size_t]),
    "lcpc_dev_hash_columns": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]),
    "lcpc_dev_merkle_tree": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "lcpc_dev_hash_merge_tree": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p, C.c_size_t, C.c_size_t]),
    "lcpc_dev_merkleize": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]),
    "lcpc_dev_fold": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p,
                                  C.c_size_t, C.c_void_p]),
    "lcpc_dev_add_partials": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "lcpc_dev_gather_columns": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                            C.c_size_t, C.c_void_p]),
    "lcpc_dev_gather_paths": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lcpc_dev_pack_bytes7": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lcpc_dev_hash_chunk_range": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.c_size_t, C.c_size_t,
                                              C.c_size_t, C.c_uint64, C.c_uint64, C.c_void_p]),
    "lcpc_dev_hash_chunk_range_scatter": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.c_size_t, C.c_size_t,
                                                      C.c_size_t, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t]),
    "lcpc_dev_hash_merge": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p]),
}

_lib = None


def library_path() -> str:
    from . import build as _build

    return _build.LIB


def declared_symbols() -> List[str]:
    """Every function include/lcpc_b200.h declares (parsed from the header text)."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lcpc_[a-z0-9_]+)\s*\(", text)))


def load():
    """Load (building first if needed) the CUDA library.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build

    # LCPC_B200_LIB: load another build of the same library (kernel experiments); still no fallback
    path = os.environ.get("LCPC_B200_LIB") or _build.build()
    lib = C.CDLL(path)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != LCPC_OK:
        msg = load().lcpc_last_error()
        raise LcpcError(status, msg.decode() if msg else "")
