"""In-tree build of the CUDA library (sm_100a only).

    python -m lcpc_proof_of_storage_b200.build [--force] [--verbose]

Produces lcpc_proof_of_storage_b200/_lib/liblcpc_b200.so with plain `nvcc`; the .so is
git-ignored but travels to the GPU box with the source tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "liblcpc_b200.so")
SOURCES = ["lcpc_ntt_f0.cu", "lcpc_ntt_f1.cu", "lcpc_ntt_f2.cu", "lcpc_ntt_f3.cu", "lcpc_ntt_f4.cu", "lcpc_ntt.cu", "lcpc_hash.cu", "lcpc_linalg.cu", "lcpc_api.cu", "lcpc_scheme.cu", "lcpc_hostrand.cu", "lcpc_stream.cu", "lcpc_ubench.cu", "lcpc_multi.cu"]
HEADERS = ["lcpc_ntt_impl.cuh", "lcpc_field.cuh", "lcpc_mont32.cuh", "lcpc_blake3.cuh", "lcpc_kernels.h", "lcpc_handles.h", "lcpc_hostrand.h", os.path.join("..", "..", "include", "lcpc_b200.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
    "-ccbin", "/usr/bin/g++",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: lcpc_proof_of_storage_b200 needs the CUDA toolkit to build")


def _digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS + [os.path.basename(__file__)]:
        path = os.path.join(CSRC, name) if name != os.path.basename(__file__) else os.path.abspath(__file__)
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "build.stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    objs = []

    def compile_one(src: str) -> str:
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}")
        if verbose:
            sys.stderr.write(res.stdout)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
