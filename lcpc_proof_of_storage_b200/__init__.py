"""lcpc_proof_of_storage_b200 -- B200-native lcpc commitment hot path.

Host-side mirror of the reference's interface for this path (lcpc-2d's `LcEncoding`,
`LcCommit::commit`, `prove`, `LcEvalProof::verify`, the Ligero and Brakedown encodings,
and proof-of-storage's file commit) over the C ABI in include/lcpc_b200.h.  All
arithmetic runs in hand-written CUDA kernels for sm_100a (csrc/); there is no CPU
fallback.
"""
from . import _lib
from ._lib import LcpcError
from .lcpc2d import (
    FT63, FT127, FT191, FT255, FT253_192, FIELD_LIMBS, FIELD_NAMES,
    Context, LigeroEncoding, SdigEncoding, CscMatrix, LcCommit, LcColumn, LcEvalProof,
    ProverError, VerifierError, Transcript, commit, prove, verify, open_column, collapse_columns, log2, next_pow2,
    n_degree_tests, sdig_codeword_length,
)

__all__ = [
    "FT63", "FT127", "FT191", "FT255", "FT253_192", "FIELD_LIMBS", "FIELD_NAMES", "Context", "LigeroEncoding",
    "SdigEncoding", "CscMatrix", "LcCommit", "LcColumn", "LcEvalProof", "ProverError", "VerifierError",
    "LcpcError", "Transcript", "commit", "prove", "verify", "sdig_codeword_length", "open_column", "collapse_columns", "log2", "next_pow2", "n_degree_tests",
]
