#!/usr/bin/env python
"""Headline benchmark: lcpc commit throughput (field elements / s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Ligero commit of a 2^24-coefficient polynomial over
the 63-bit field, rho = 1/2, BLAKE3: 512 rows x 32768 -> 65536 columns.  One "step" is one
commit: encode every row (batched NTT), hash every column, build the Merkle tree.
At N > 1 the matrix has N x 512 rows of the same width (weak scaling); rows are sharded
for encoding (chunk-aligned row blocks), every rank hashes the BLAKE3 chunks of ALL columns
over its own rows and stores the 32-byte chaining values straight into the owning rank's
store over NVLink (3 % of the encoded matrix), the owner merges them into leaves and builds
its Merkle subtree, and rank 0 joins the N subtree roots.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs in HBM,
CUDA events, max over ranks); `e2e` goes through the host-buffer path (pinned host
coefficients in, encoded matrix + Merkle tree out; `e2e.root_only` is the same with only the
32-byte root coming back).  At N = 1 the line also carries `configs`: the other BASELINE.json
configurations (Ligero 2^28, Brakedown Ft255 2^24, the (n_dt+1)-tensor folds), each timed and
checked against a committed known answer, and `integer_peak`: the integer-pipe issue rates
measured on this device, the denominators of the second roofline.  At N > 1 small sharded cases
(Brakedown, both hashing modes, folds, openings, the byte commit) are checked against committed
known answers BEFORE the timed region (`parity_checks`); a mismatch is a non-zero exit.
`--impl reference` times the CPU restatement of the reference's algorithm (oracle/) with
all host threads on the same workload -- the Rust reference itself cannot be built here.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

FID = 0  # Ft63
LOG_N = 24
N_PER_ROW, N_COLS = 32768, 65536
ROWS_PER_GPU = (1 << LOG_N) // N_PER_ROW  # 512
METRIC = "ligero_commit_throughput_ft63"  # the size is the workload's (config.workload): N x 2^24 coefficients at N GPUs
UNIT = "field elements/s"


# The other BASELINE.json configurations, timed at N = 1 after the headline workload (`configs` in the line); inputs are
# the seeded streams of lcpc_proof_of_storage_b200/synth.py, known answers in tests/golden/bench_roots.json
CONFIG_CASES = {
    "ligero_ft63_2^28": {"seed": 28, "n_per_row": 131072, "n_cols": 262144},
    "brakedown_ft255_2^24": {"seed": 3, "code_seed": 0, "n_per_row": 166292, "n_cols": 252931},
    "brakedown_ft63_2^24": {"seed": 4, "code_seed": 0, "n_per_row": 166293, "n_cols": 252932},   # configs[4]'s Brakedown point
    "ligero_ft63_2^20": {"seed": 20, "n_per_row": 8192, "n_cols": 16384},                       # configs[4]'s small end
    "fold_ft63_2^24": {"tensor_seed": 41, "n_tensors": 4},   # n_degree_tests + 1 = 4 at these widths
    "fold_ft63_2^28": {"tensor_seed": 42, "n_tensors": 4},
}
# Small sharded cases checked at N > 1 before the timed region (`parity_checks` in the line); the row counts leave every
# rank of 2, 4 or 8 at least one BLAKE3 chunk of rows
PARITY_CASES = {
    "ligero": {"seed": 7, "tensor_seed": 8, "n_rows": 1024, "n_per_row": 512, "n_cols": 1024, "open": [0, 5, 511, 512, 1023, 700, 5]},
    "bytes": {"seed": 9, "n_rows": 1024, "n_per_row": 512, "n_cols": 1024, "n_bytes": 7 * 512 * 1024 - 13},
    "brakedown": {"seed": 10, "code_seed": 0, "n_rows": 1024, "n_per_row": 3000, "n_cols": 4563},
}


def workload_name(n_gpus: int) -> str:
    rows = ROWS_PER_GPU * n_gpus
    return (f"Ligero commit, Ft63, rho=1/2, BLAKE3, {rows} rows x {N_PER_ROW} -> {N_COLS} cols "
            f"({rows * N_PER_ROW} coefficients)")


def workload_config(n_gpus: int) -> dict:
    """`config` of the result line: the workload and nothing else, identical in both arms (ours and --impl reference)."""
    return {"workload": workload_name(n_gpus), "field": "Ft63", "n_rows": ROWS_PER_GPU * n_gpus, "n_per_row": N_PER_ROW,
            "n_cols": N_COLS, "digest": "BLAKE3", "l2": "inputs larger than L2 (128 MiB in, 256 MiB out per GPU)"}


def algorithmic_bytes(n_coeffs: int, n_rows: int) -> int:
    """read coeffs once + write encoded matrix once + write the tree."""
    return n_coeffs * 8 + n_rows * N_COLS * 8 + (2 * N_COLS - 1) * 32


def make_coeffs(seed: int, n: int) -> np.ndarray:
    """Seeded Ft63 elements (Montgomery limbs): splitmix64 stream masked to 63 bits, reduced once."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)
        z = (idx ^ (idx >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    z &= np.uint64((1 << 63) - 1)
    p = np.uint64(5102708120182849537)
    z = np.where(z >= p, z - p, z)
    return z.reshape(n, 1)


def bind_to_gpu_numa_node(index: int) -> str:
    """Pin this rank to the CPUs NVML reports as local to its GPU before any pinned host buffer is allocated, so that
    first touch puts the buffers on the GPU's NUMA node (with N ranks copying 128 MiB each, remote-socket buffers are what
    limits the end-to-end step).  Best effort: returns what was done."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to gpu {index}"
        return "no local cpus reported"
    except Exception as e:  # NVML missing, cgroup restrictions, ...
        return f"unbound ({type(e).__name__})"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, enabled: bool = True):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.nv = None
        if not enabled:  # ranks other than 0: their samples are never reported, and concurrent NVML queries slow each other
            return
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._sample()  # pays NVML's first-call costs outside any timed region; discarded (the GPU may be idle)
            self.samples, self.reasons = [], set()
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        try:
            self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            names = {
                0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting",
                0x100: "display_clock_setting",
            }
            for bit, name in names.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self):
        if not self.nv:
            return
        self._stop.clear()

        def loop():
            while not self._stop.is_set():
                self._sample()
                time.sleep(0.01)

        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self):
        if self._thread:
            self._sample()
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def expected_root(n_gpus: int):
    """Merkle root of the N-GPU workload from the committed fixture (None when N is not in it)."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "bench_roots.json")) as f:
            return json.load(f)["roots_by_n_gpus"].get(str(n_gpus))
    except OSError:
        return None


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"




# per-kernel compulsory HBM bytes for one launch at this workload (DESIGN.md "Kernels")
def kernel_algorithmic_bytes(name: str, n_rows: int) -> int:
    enc = n_rows * N_COLS * 8
    return {
        "k_ntt_strided": n_rows * N_PER_ROW * 8 + enc,  # reads the coefficients, writes the widened rows
        "k_ntt_block": 2 * enc,                          # in place over the encoded matrix
        "k_ntt_block_scatter": 2 * enc,                  # same pass, stores go to the owning ranks' column blocks
        "k_hash_chunks": enc + ((32 + n_rows * 8 + 1023) // 1024) * N_COLS * 32,
        "k_hash_chunks_scatter": enc + ((32 + n_rows * 8 + 1023) // 1024) * N_COLS * 32,
        # leaf hashing + chunk-value merge + tree in one launch: reads the encoded matrix, writes the tree
        "k_hash_tree": enc + (2 * N_COLS - 1) * 32,
        "k_merge_tree": ((32 + n_rows * 8 + 1023) // 1024) * N_COLS * 32 + (2 * N_COLS - 1) * 32,
        "k_hash_merge": ((32 + n_rows * 8 + 1023) // 1024) * N_COLS * 32 + N_COLS * 32,
        "k_merkle_levels": 2 * N_COLS * 32,
    }.get(name, 0)


# Integer-pipe model of the kernels ("Integer-pipe model"): per SM sub-partition a warp instruction
# of the IMAD.WIDE class costs 4 issue cycles on the FMA pipe and holds the ALU side for 2; IMAD / IADD3 / LOP3 / SHF / PRMT
# cost 2 on their pipe.  The table gives the ESSENTIAL instruction counts per unit of work (W = IMAD.WIDE, I = other FMA-
# pipe integer ops, A = ALU-pipe ops), derived in DESIGN.md from the arithmetic alone -- no addressing, no loop control.
# unit = one field element through the kernel (NTT) or one BLAKE3 compression (hashing).
INT_PIPE_MODEL = {
    # 12 of the 16 butterfly levels: 5.06 Montgomery products (6 W + 2 I + 5 A each since the digit steps are single
    # three-operand sums; 6 W + 3 I + 8 A before) and 12 modular add/sub (1 I + 4 A each)
    "k_ntt_block": {"unit": "element", "W": 30.4, "I": 22.1, "A": 73.3},
    "k_ntt_block_scatter": {"unit": "element", "W": 30.4, "I": 22.1, "A": 73.3},
    # top 4 levels of a rate-1/2 row: 2.0 products per element, 3 add/sub levels
    "k_ntt_strided": {"unit": "element", "W": 12.0, "I": 7.0, "A": 22.0},
    # 56 G functions: 6 adds on the FMA pipe, 4 XOR + 2 PRMT + 2 funnel shifts on the ALU pipe each; 8 final XORs;
    # 8 de-Montgomery reductions (2 W + 2 I + 5 A each)
    "k_hash_chunks": {"unit": "compression", "W": 16, "I": 352, "A": 496},
    "k_hash_chunks_scatter": {"unit": "compression", "W": 16, "I": 352, "A": 496},
    # + per column 4 parent compressions (5 chunk values) and 1 Merkle node: 5 / 65 more compressions, no reductions
    "k_hash_tree": {"unit": "compression", "W": 16, "I": 352 * 70 / 65, "A": 496 * 70 / 65},
}


def int_pipe_floor_ms(name: str, units: float, n_sm: int, sm_mhz: float, cyc: dict = None):
    """Time the kernel would take if its binding integer pipe never idled: max(cW W + cI I, cW/2 W + cA A) issue cycles per
    warp instruction group, 32 units per warp, 4 sub-partitions per SM.  cyc = measured cycles per warp instruction
    (lcpc_ctx_measure_int_pipes); without it the round-1 figures 4 / 2 / 2 (profiles/r01c_int_pipes.txt)."""
    m = INT_PIPE_MODEL.get(name)
    if not m or not sm_mhz:
        return None
    cw, ci, ca = (cyc["imad_wide"], cyc["imad"], cyc["lop3"]) if cyc else (4.0, 2.0, 2.0)
    cycles = max(cw * m["W"] + ci * m["I"], cw / 2 * m["W"] + ca * m["A"])
    return cycles * (units / 32.0) / (4 * n_sm) / (sm_mhz * 1e3)


def integer_pipe_report(kernels_ms_per_step: dict, n_sm: int, sm_mhz, cyc: dict = None) -> dict:
    """{kernel: {floor_ms, measured_ms, frac}} for the kernels the model covers (one launch of each per step)."""
    units = {"element": ROWS_PER_GPU * N_COLS, "compression": N_COLS * ((32 + ROWS_PER_GPU * 8 + 63) // 64)}
    out = {}
    for k, ms in kernels_ms_per_step.items():
        if k not in INT_PIPE_MODEL or not ms:
            continue
        f = int_pipe_floor_ms(k, units[INT_PIPE_MODEL[k]["unit"]], n_sm, sm_mhz, cyc)
        if f:
            out[k] = {"floor_ms": round(f, 4), "measured_ms": round(ms, 4), "frac": round(f / ms, 3)}
    return out


def golden(section: str) -> dict:
    try:
        with open(os.path.join(ROOT, "tests", "golden", "bench_roots.json")) as f:
            return json.load(f).get(section, {})
    except OSError:
        return {}


def dram_traffic(name: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `name` from the round's `ncu --set full` capture of this
    workload on the shipped build (profiles/r02_dram_traffic.json, written by tools/ncu_traffic.py); None when the kernel
    is not in it."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_dram_traffic.json")) as f:
            d = json.load(f)
        v = d["kernels"].get(name)
        return (v["dram_read_bytes"] + v["dram_write_bytes"], d.get("source")) if v else (None, None)
    except Exception:
        return None, None


def _sha(t) -> str:
    import hashlib

    return hashlib.sha256(t.contiguous().cpu().numpy().tobytes()).hexdigest()


def _time_steps(torch, stream, fn, steps: int, warmup: int) -> float:
    for _ in range(warmup):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(steps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def run_configs(torch, P, lib, _lib, ctx, stream, peak_gbs: float, d_coeffs_24) -> dict:
    """The other BASELINE.json configurations on one GPU, device-resident, CUDA-event timed, each checked against its
    committed known answer (tests/golden/bench_roots.json `configs`, computed by the CPU oracle on the same seeded
    inputs).  Per case: ms, field elements / s, SURVEY 8(d) algorithmic GB/s and its fraction of the measured HBM peak."""
    from lcpc_proof_of_storage_b200 import synth as S

    want = golden("configs")
    dev = torch.device("cuda", ctx.device)
    out = {}

    def entry(ms, n_elems, alg_bytes, ok, **extra):
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        e = {"ms": ms, "elements_per_s": n_elems / (ms * 1e-3), "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak_gbs,
             "matches_known_answer": ok}
        e.update(extra)
        return e

    def fold_case(name, d_coeffs, n_rows, n_per_row):
        k = CONFIG_CASES[name]
        nt = k["n_tensors"]
        tens = S.ft63_torch(k["tensor_seed"], nt * n_rows, dev)
        d_out = torch.empty(nt * n_per_row, dtype=torch.int64, device=dev)

        def fn():
            _lib.check(lib.lcpc_dev_fold(ctx.handle, FID, d_coeffs.data_ptr(), n_rows, n_per_row, n_per_row, tens.data_ptr(),
                                         nt, d_out.data_ptr()))

        ms = _time_steps(torch, stream, fn, 20, 3)
        ok = (_sha(d_out) == want.get(name, {}).get("sha256")) if name in want else None
        # collapse_columns: the matrix is streamed once for all n_dt + 1 tensors (8 B per coefficient per pass)
        return entry(ms, n_rows * n_per_row, n_rows * n_per_row * 8 + nt * (n_rows + n_per_row) * 8, ok,
                     workload=f"collapse_columns, Ft63, {nt} tensors (n_degree_tests + 1) in one pass over {n_rows} x {n_per_row}")

    out["fold_ft63_2^24"] = fold_case("fold_ft63_2^24", d_coeffs_24, ROWS_PER_GPU, N_PER_ROW)

    # ---- Ligero Ft63, 2^28 coefficients: 2048 x 131072 -> 262144 (2 GiB in, 4 GiB encoded)
    k = CONFIG_CASES["ligero_ft63_2^28"]
    n, npr, nc = 1 << 28, k["n_per_row"], k["n_cols"]
    n_rows = n // npr
    enc = P.LigeroEncoding(FID, npr, nc, ctx=ctx)
    d_c = S.ft63_torch(k["seed"], n, dev)
    d_m = torch.empty(n_rows * nc, dtype=torch.int64, device=dev)
    d_h = torch.zeros((2 * nc - 1) * 32, dtype=torch.uint8, device=dev)

    def commit28():
        _lib.check(lib.lcpc_dev_encode(enc.plan, d_c.data_ptr(), n_rows, d_m.data_ptr()))
        _lib.check(lib.lcpc_dev_merkleize(ctx.handle, FID, d_m.data_ptr(), n_rows, nc, nc, d_h.data_ptr()))

    ms = _time_steps(torch, stream, commit28, 5, 2)
    root = bytes(d_h[-32:].cpu().numpy()).hex()
    ok = (root == want.get("ligero_ft63_2^28", {}).get("root")) if "ligero_ft63_2^28" in want else None
    out["ligero_ft63_2^28"] = entry(ms, n, n * 8 + n_rows * nc * 8 + (2 * nc - 1) * 32, ok, root=root,
                                    workload=f"Ligero commit, Ft63, rho=1/2, BLAKE3, {n_rows} rows x {npr} -> {nc} cols")
    del d_m, d_h
    out["fold_ft63_2^28"] = fold_case("fold_ft63_2^28", d_c, n_rows, npr)
    del d_c, enc
    torch.cuda.empty_cache()

    # ---- Brakedown code 3 over Ft255, 2^24 coefficients: 101 x 166292 -> 252931
    k = CONFIG_CASES["brakedown_ft255_2^24"]
    n = 1 << 24
    enc = P.SdigEncoding.new(P.FT255, n, seed=k["code_seed"], ctx=ctx)
    assert (enc.n_per_row, enc.n_cols) == (k["n_per_row"], k["n_cols"])
    npr, nc = enc.n_per_row, enc.n_cols
    n_rows = (n + npr - 1) // npr
    np2 = P.next_pow2(nc)
    d_c = torch.zeros(n_rows * npr * 4, dtype=torch.int64, device=dev)   # zero-padded last row (lib.rs:665-674)
    d_c[:n * 4] = S.ft255_torch(k["seed"], n, dev)
    d_m = torch.empty(n_rows * nc * 4, dtype=torch.int64, device=dev)
    d_h = torch.zeros((2 * np2 - 1) * 32, dtype=torch.uint8, device=dev)

    def commit_bd():
        _lib.check(lib.lcpc_dev_encode(enc.plan, d_c.data_ptr(), n_rows, d_m.data_ptr()))
        _lib.check(lib.lcpc_dev_merkleize(ctx.handle, P.FT255, d_m.data_ptr(), n_rows, nc, nc, d_h.data_ptr()))

    ms = _time_steps(torch, stream, commit_bd, 5, 2)
    root = bytes(d_h[-32:].cpu().numpy()).hex()
    ok = (root == want.get("brakedown_ft255_2^24", {}).get("root")) if "brakedown_ft255_2^24" in want else None
    out["brakedown_ft255_2^24"] = entry(ms, n, n * 32 + n_rows * nc * 32 + (2 * np2 - 1) * 32, ok, root=root,
                                        workload=f"Brakedown (code 3) commit, Ft255, BLAKE3, {n_rows} rows x {npr} -> {nc} cols")
    del d_c, d_m, d_h, enc
    torch.cuda.empty_cache()

    # ---- Brakedown code 3 over Ft63, 2^24 coefficients: 101 x 166293 -> 252932 (configs[4])
    k = CONFIG_CASES["brakedown_ft63_2^24"]
    enc = P.SdigEncoding.new(FID, n, seed=k["code_seed"], ctx=ctx)
    assert (enc.n_per_row, enc.n_cols) == (k["n_per_row"], k["n_cols"])
    npr, nc = enc.n_per_row, enc.n_cols
    n_rows = (n + npr - 1) // npr
    np2 = P.next_pow2(nc)
    d_c = torch.zeros(n_rows * npr, dtype=torch.int64, device=dev)
    d_c[:n] = S.ft63_torch(k["seed"], n, dev)
    d_m = torch.empty(n_rows * nc, dtype=torch.int64, device=dev)
    d_h = torch.zeros((2 * np2 - 1) * 32, dtype=torch.uint8, device=dev)

    def commit_bd63():
        _lib.check(lib.lcpc_dev_encode(enc.plan, d_c.data_ptr(), n_rows, d_m.data_ptr()))
        _lib.check(lib.lcpc_dev_merkleize(ctx.handle, FID, d_m.data_ptr(), n_rows, nc, nc, d_h.data_ptr()))

    ms = _time_steps(torch, stream, commit_bd63, 10, 3)
    root = bytes(d_h[-32:].cpu().numpy()).hex()
    ok = (root == want.get("brakedown_ft63_2^24", {}).get("root")) if "brakedown_ft63_2^24" in want else None
    out["brakedown_ft63_2^24"] = entry(ms, n, n * 8 + n_rows * nc * 8 + (2 * np2 - 1) * 32, ok, root=root,
                                       workload=f"Brakedown (code 3) commit, Ft63, BLAKE3, {n_rows} rows x {npr} -> {nc} cols")
    del d_c, d_m, d_h, enc

    # ---- Ligero Ft63, 2^20 coefficients: 128 x 8192 -> 16384 (configs[4]'s small end: launches and dependent tails)
    k = CONFIG_CASES["ligero_ft63_2^20"]
    n, npr, nc = 1 << 20, k["n_per_row"], k["n_cols"]
    n_rows = n // npr
    enc = P.LigeroEncoding(FID, npr, nc, ctx=ctx)
    d_c = S.ft63_torch(k["seed"], n, dev)
    d_m = torch.empty(n_rows * nc, dtype=torch.int64, device=dev)
    d_h = torch.zeros((2 * nc - 1) * 32, dtype=torch.uint8, device=dev)

    def commit20():
        _lib.check(lib.lcpc_dev_encode(enc.plan, d_c.data_ptr(), n_rows, d_m.data_ptr()))
        _lib.check(lib.lcpc_dev_merkleize(ctx.handle, FID, d_m.data_ptr(), n_rows, nc, nc, d_h.data_ptr()))

    ms = _time_steps(torch, stream, commit20, 50, 5)
    root = bytes(d_h[-32:].cpu().numpy()).hex()
    ok = (root == want.get("ligero_ft63_2^20", {}).get("root")) if "ligero_ft63_2^20" in want else None
    out["ligero_ft63_2^20"] = entry(ms, n, n * 8 + n_rows * nc * 8 + (2 * nc - 1) * 32, ok, root=root,
                                    workload=f"Ligero commit, Ft63, rho=1/2, BLAKE3, {n_rows} rows x {npr} -> {nc} cols")
    return out


def run_pos_config(torch, dist, P, ctx, stream, rank: int, world: int, peak_gbs: float):
    """BASELINE configs[3] at N > 1: proof-of-storage commit of a 4 GiB synthetic file sharded over the ranks (7-byte packing
    on the device, default aspect of networking/server.rs:1139-1182: 18725 rows x 32768 -> 65536), then the retrievability
    proof -- the 309 columns of get_column_indicies_from_random_seed(1337, ..) opened with their Merkle paths -- and its
    verification on rank 0.  Root, column indices, opened values, paths and leaves are checked against the CPU oracle's
    (tests/golden/pos_4gib.json).  Returns the `configs` entry on rank 0, None elsewhere."""
    import hashlib

    from lcpc_proof_of_storage_b200 import pos
    from lcpc_proof_of_storage_b200 import synth as S
    from lcpc_proof_of_storage_b200.sharded import ShardedLigeroCommitter

    n_bytes = 1 << 32
    pre, enc_cols, soundness = pos.get_aspect_ratio_default_from_file_len(n_bytes)
    n_elems = (n_bytes + 6) // 7
    n_rows = (n_elems + pre - 1) // pre
    dev = torch.device("cuda", ctx.device)
    enc = P.LigeroEncoding(P.FT63, pre, enc_cols, ctx=ctx)
    sc = ShardedLigeroCommitter(enc, n_rows, dist.group.WORLD, hashing="auto")
    lo, hi = sc.byte_range(n_bytes)
    lo8 = lo - lo % 8
    data = S.bytes_torch(4, hi - lo8, dev, lo8)[lo - lo8:].clone()
    for _ in range(2):
        sc.commit_bytes(data)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        sc.commit_bytes(data)
    e1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_commit = float(t.item())
    cols = pos.get_column_indicies_from_random_seed(1337, soundness, enc_cols)
    sc.open_columns(cols)  # first use: staging buffers, NCCL channels
    ms_open_dev = None
    if sc.hashing == "rows":
        dist.barrier()
        torch.cuda.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(5):
            sc.open_columns_dev(cols)
        e3.record(stream)
        torch.cuda.synchronize()
        ms_open_dev = e2.elapsed_time(e3) / 5
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    opened = sc.open_columns(cols)
    torch.cuda.synchronize()
    ms_open_host = (time.perf_counter() - t0) * 1e3
    if rank != 0:
        return None
    root = sc.root()
    t0 = time.perf_counter()
    try:
        pos.client_online_verify_column_paths(root, cols, opened, ctx)
        verified = True
    except Exception:
        verified = False
    ms_verify = (time.perf_counter() - t0) * 1e3
    try:
        with open(os.path.join(ROOT, "tests", "golden", "pos_4gib.json")) as f:
            want = json.load(f)
    except OSError:
        want = None
    known = None
    if want:
        vals = np.stack([o.col for o in opened])
        paths = np.stack([o.path for o in opened])
        leaves, _ = pos._verify_columns(ctx, opened, None, None)
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
        known = {"root": root.hex() == want["root"], "columns": cols == want["columns"],
                 "column_values": sha(vals) == want["column_values_sha256"], "paths": sha(paths) == want["paths_sha256"],
                 "leaves": sha(leaves) == want["leaves_sha256"]}
    alg = n_bytes + n_rows * enc_cols * 8 + (2 * enc_cols - 1) * 32   # PoS commit from bytes
    gbs = alg / (ms_commit * 1e-3) / 1e9
    return {"workload": f"proof-of-storage commit, 4 GiB file, WriteableFt63, {n_rows} rows x {pre} -> {enc_cols}, sharded over "
                        f"{world} GPUs; retrievability proof of {len(cols)} columns (seed 1337)",
            "ms": ms_commit, "elements_per_s": n_elems / (ms_commit * 1e-3), "file_GBps": n_bytes / (ms_commit * 1e-3) / 1e9,
            "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / (peak_gbs * world), "hashing": sc.hashing, "cv_fused": sc.cv_fused,
            "open_columns": len(cols), "open_on_device_ms": ms_open_dev, "open_to_host_ms": ms_open_host,
            "open_bytes": len(cols) * (n_rows * 8 + int(opened[0].path.size)), "verify_paths_ms": ms_verify,
            "verified": verified, "root": root.hex(), "matches_known_answer": None if known is None else all(known.values()),
            "known_answers": known}


def sharded_parity_checks(torch, dist, P, ctx, rank: int, world: int) -> dict:
    """Small sharded cases against committed known answers (tests/golden/bench_roots.json `parity`), run on every rank
    before the timed region: both hashing modes and the fused exchange on a Ligero commit, the coefficient fold, the fold
    over the encoded matrix, openings (values + paths, checked again by lcpc_verify_columns_host), the proof-of-storage
    byte commit and a Brakedown commit (lcpc-brakedown-pc/src/encode.rs:36-94, lcpc-2d/src/lib.rs:736-775).
    Returns {check: bool} on rank 0 ({} elsewhere)."""
    import numpy as np

    from lcpc_proof_of_storage_b200 import pos
    from lcpc_proof_of_storage_b200 import synth as S
    from lcpc_proof_of_storage_b200.sharded import ShardedCommitter

    want = golden("parity")
    dev = torch.device("cuda", ctx.device)
    res = {}

    def put(name, value):
        if rank == 0:
            res[name] = bool(value)

    k = PARITY_CASES["ligero"]
    enc = P.LigeroEncoding(FID, k["n_per_row"], k["n_cols"], ctx=ctx)
    tens = S.ft63_torch(k["tensor_seed"], 2 * k["n_rows"], dev)
    for mode in ("columns", "rows", "auto"):
        sc = ShardedCommitter(enc, k["n_rows"], dist.group.WORLD, hashing=mode)
        coeffs = S.ft63_torch(k["seed"], sc.rows_local * k["n_per_row"], dev, start=sc.row0 * k["n_per_row"])
        sc.commit(coeffs)
        label = {"columns": "ligero_columns", "rows": "ligero_rows_nccl", "auto": "ligero_rows_fused"}[mode]
        put(label + "_root", rank != 0 or sc.root().hex() == want.get("ligero_root"))
        f = sc.fold(tens)
        fe = sc.fold_encoded(tens)
        put(label + "_fold", _sha(f) == want.get("fold_sha256"))
        put(label + "_fold_encoded", _sha(fe) == want.get("fold_encoded_sha256"))
        opened = sc.open_columns(k["open"])
        if rank == 0:
            cols = np.stack([o.col for o in opened])
            paths = np.stack([o.path for o in opened])
            import hashlib

            put(label + "_open_columns", hashlib.sha256(cols.tobytes()).hexdigest() == want.get("open_cols_sha256"))
            put(label + "_open_paths", hashlib.sha256(np.ascontiguousarray(paths).tobytes()).hexdigest() == want.get("open_paths_sha256"))
            try:
                pos.client_online_verify_column_paths(sc.root(), k["open"], opened, ctx)
                put(label + "_verify_columns", True)
            except Exception:
                put(label + "_verify_columns", False)
        if mode == "auto":
            put("auto_picks_fused_row_hashing", sc.hashing == "rows" and sc.cv_fused)
        del sc
    k = PARITY_CASES["bytes"]
    enc_b = P.LigeroEncoding(FID, k["n_per_row"], k["n_cols"], ctx=ctx)
    sc = ShardedCommitter(enc_b, k["n_rows"], dist.group.WORLD, hashing="auto")
    lo, hi = sc.byte_range(k["n_bytes"])
    lo8 = lo - lo % 8
    data = S.bytes_torch(k["seed"], hi - lo8, dev, lo8)[lo - lo8:].clone()
    sc.commit_bytes(data)
    put("commit_bytes_root", rank != 0 or sc.root().hex() == want.get("bytes_root"))
    del sc
    k = PARITY_CASES["brakedown"]
    enc_s = P.SdigEncoding.new_from_dims(FID, k["n_per_row"], k["n_cols"], seed=k["code_seed"], ctx=ctx)
    n_total = k["n_rows"] * k["n_per_row"] - 5
    # columns = the transposing passes store into the owners' column blocks over NVLink; columns_nccl = the same blocks by
    # pack + NCCL all-to-all; auto = row hashing with the chaining values stored into the owners' stores
    for mode in ("columns", "columns_nccl", "rows", "auto"):
        sc = ShardedCommitter(enc_s, k["n_rows"], dist.group.WORLD, hashing=mode.split("_")[0],
                              fused=False if mode == "columns_nccl" else None)
        if mode.startswith("columns"):
            put(f"brakedown_{mode}_exchange_as_named", sc.fused == (mode == "columns"))
        e0, e1 = sc.row0 * k["n_per_row"], (sc.row0 + sc.rows_local) * k["n_per_row"]
        coeffs = torch.zeros(sc.rows_local * k["n_per_row"], dtype=torch.int64, device=dev)
        have = max(0, min(n_total, e1) - e0)
        if have:
            coeffs[:have] = S.ft63_torch(k["seed"], have, dev, start=e0)
        sc.commit(coeffs)
        put(f"brakedown_{'rows_fused' if mode == 'auto' else mode}_root", rank != 0 or sc.root().hex() == want.get("brakedown_root"))
        if mode == "auto":
            put("brakedown_auto_picks_fused_row_hashing", sc.hashing == "rows" and sc.cv_fused)
            opened = sc.open_columns([0, k["n_cols"] - 1, 3000, 4095])
            if rank == 0:
                try:
                    pos.client_online_verify_column_paths(sc.root(), [0, k["n_cols"] - 1, 3000, 4095], opened, ctx)
                    put("brakedown_rows_fused_open_verify", True)
                except Exception:
                    put("brakedown_rows_fused_open_verify", False)
        del sc
    return res


def run_reference(args, rank: int, world: int) -> None:
    """CPU arm: the oracle's restatement of the reference algorithm, all host threads."""
    if rank != 0:
        return
    from oracle import lcpc_oracle as O

    O.build()
    O.set_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: use every host core
    cores = O.max_threads()
    n_rows = ROWS_PER_GPU * max(1, args.gpus)
    # bounded sample: at most 512 rows of the same width per step (throughput per coefficient does
    # not depend on the row count: rows are independent and leaves are hashed row-block by row-block)
    sample_rows = min(n_rows, 512)
    n = sample_rows * N_PER_ROW
    coeffs = make_coeffs(2, n)
    enc = O.LigeroEncoding(FID, N_PER_ROW, N_COLS)
    steps, warmup = args.steps, args.warmup
    t_probe = time.perf_counter()
    O.commit(coeffs, enc)
    t_probe = time.perf_counter() - t_probe
    if t_probe * (steps + warmup) > 240:  # keep the run within a few minutes
        steps = max(1, int(240 / t_probe) - 1)
        warmup = 0
    for _ in range(max(0, warmup - 1)):
        O.commit(coeffs, enc)
    t0 = time.perf_counter()
    for _ in range(steps):
        c = O.commit(coeffs, enc)
    dt = time.perf_counter() - t0
    value = n * steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 (63-bit prime field, Montgomery)", "data": "synthetic",
        "config": workload_config(max(1, args.gpus)), "sample": f"{sample_rows} of {n_rows} rows per step",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample_rows} rows x {N_PER_ROW} -> {N_COLS} (full width), C restatement "
                                   f"of the reference algorithm, OpenMP over rows / 32-column blocks"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "root": c.get_root().hex(),
    }
    _emit(line)


_REAL_STDOUT = None


def _claim_stdout() -> None:
    """stdout carries exactly ONE line, the result JSON: libraries that write to fd 1 while the run is set up (NCCL prints its
    version there) are pointed at stderr, and the JSON goes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main() -> None:
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hashing", default="rows-fused", choices=["columns", "rows", "rows-fused"],
                    help="N > 1 only.  'rows-fused' (default, the measured winner from 2 GPUs up): BLAKE3 chunks are hashed "
                         "where the rows are and the hash kernel stores the 32-byte chaining values into the owners' stores "
                         "over NVLink; 'rows': the same values through an NCCL all-to-all; 'columns': the encoded matrix is "
                         "re-sharded to column blocks (the NTT's last pass stores into peer HBM)")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the `configs` block (N = 1: Ligero 2^28, Brakedown Ft255 2^24, folds; N > 1: the 4 GiB proof-of-storage file)")
    ap.add_argument("--no-parity-checks", action="store_true", help="N > 1: skip the sharded known-answer checks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch

    import lcpc_proof_of_storage_b200 as P
    from lcpc_proof_of_storage_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lcpc_proof_of_storage_b200 has no CPU fallback")
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single process: not bound"
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    stream = torch.cuda.current_stream()
    ctx = P.Context(local_rank, stream=stream.cuda_stream)
    parity_checks = None
    if world > 1 and not args.no_parity_checks:
        parity_checks = sharded_parity_checks(torch, dist, P, ctx, rank, world)
        bad = torch.tensor([0 if all(parity_checks.values()) else 1], dtype=torch.int32, device="cuda")
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()):
            if rank == 0:
                sys.stderr.write("sharded parity checks FAILED: " + json.dumps(parity_checks) + "\n")
            dist.destroy_process_group()
            raise SystemExit(3)
        torch.cuda.empty_cache()
    enc = P.LigeroEncoding(FID, N_PER_ROW, N_COLS, ctx=ctx)

    n_rows_total = ROWS_PER_GPU * world
    n_total = n_rows_total * N_PER_ROW
    np2 = N_COLS
    # this rank's rows of the coefficient matrix (seed 2 stream, sliced by row block)
    row0, rows_local = rank * ROWS_PER_GPU, ROWS_PER_GPU
    if world > 1 and args.hashing != "columns":  # chunk-aligned row blocks: 508 / 512 / ... / 516 rows at 8 GPUs
        from lcpc_proof_of_storage_b200.sharded import chunk_row_partition

        row0, rows_local = chunk_row_partition(1, n_rows_total, world)[0][rank]
    from lcpc_proof_of_storage_b200 import synth as S

    h_coeffs_np = S.ft63_np(2, rows_local * N_PER_ROW, start=row0 * N_PER_ROW)  # = make_coeffs(2, n_total)[my rows]
    h_coeffs = torch.from_numpy(h_coeffs_np.view(np.int64).reshape(-1)).pin_memory()
    d_coeffs = h_coeffs.cuda(non_blocking=True)
    torch.cuda.synchronize()

    if world == 1:
        d_comm = torch.empty(ROWS_PER_GPU * N_COLS, dtype=torch.int64, device="cuda")
        d_hashes = torch.zeros((2 * np2 - 1) * 32, dtype=torch.uint8, device="cuda")

        def step():
            _lib.check(lib.lcpc_dev_encode(enc.plan, d_coeffs.data_ptr(), ROWS_PER_GPU, d_comm.data_ptr()))
            # merkleize: chunk hashing (k_hash_chunks), then leaf merge + the whole Merkle tree in one launch (k_merge_tree)
            _lib.check(lib.lcpc_dev_merkleize(ctx.handle, FID, d_comm.data_ptr(), ROWS_PER_GPU, N_COLS, N_COLS,
                                              d_hashes.data_ptr()))

        def root_hex():
            return bytes(d_hashes[-32:].cpu().numpy()).hex()

        def finish():
            pass
    else:
        from lcpc_proof_of_storage_b200.sharded import ShardedLigeroCommitter

        sc = ShardedLigeroCommitter(enc, n_rows_total, dist.group.WORLD, hashing=args.hashing.split("-")[0],
                                    fused=True if args.hashing == "rows-fused" else None)
        assert (sc.row0, sc.rows_local) == (row0, rows_local)

        def step():
            # commit k's column hashing is issued after commit k+1's encode (three symmetric buffers): its exchange drains
            # over NVLink behind the next encode.  finish() below completes the last one INSIDE the timed region.
            sc.commit(d_coeffs, defer=sc.fused)

        def root_hex():
            return sc.root().hex() if rank == 0 else ""

        def finish():
            sc.flush()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # NVML is set up and the sampling thread started BEFORE the barrier: its first calls take milliseconds, more when
    # eight processes make them at once, and a different time on every rank.  Between the barrier and the first event
    # that skew lands inside the timed region of the fast ranks (they wait for the slowest rank in the first commit's
    # exchange barrier): at 8 GPUs it cost 2 - 20 ms of a 15 ms region (profiles/r01c_scaling_and_configs.md).
    sampler = ClockSampler(local_rank, enabled=(rank == 0))
    for _ in range(warmup):
        step()
    finish()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = ctx.launch_count()
    # the timed region: exactly `steps` steps between two events on the launching stream, nothing else enqueued
    ev0.record(stream)
    for _ in range(steps):
        step()
    finish()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    # the same `steps` steps once more with a CUDA event before and after every launch (the library's own timer):
    # per-kernel durations for the roofline.  The events serialise launches a little (about 6 % on this step), which is
    # why the headline comes from the region above and this pass reports its own total next to the kernel times.
    ctx.kernel_timing(True)
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    for _ in range(steps):
        step()
    finish()
    ev3.record(stream)
    barrier()
    sampler.stop()
    ms_total_timed = ev2.elapsed_time(ev3)
    kt = ctx.kernel_timing_report()
    ctx.kernel_timing(False)
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / steps
    value = n_total / (ms_per_step * 1e-3)
    gpu_root = root_hex()

    # single-commit latency: one commit at a time, nothing deferred or pipelined, every rank synchronised before and after
    # (host clock around a device-synchronised region; median over the commits, max over ranks)
    lat = []
    for _ in range(min(steps, 20)):
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            step()
        else:
            sc.commit(d_coeffs)
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    latency_ms = sorted(lat)[len(lat) // 2]
    if dist is not None:
        t = torch.tensor([latency_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        latency_ms = float(t.item())

    # ---- end to end through the host-buffer C-ABI call (N = 1 path per rank) ----------------------
    e2e = None
    if world == 1:  # host destinations of the full LcCommit (the sharded leg reads back the root only)
        h_comm = torch.empty(ROWS_PER_GPU * N_COLS, dtype=torch.int64).pin_memory()
        h_hashes = torch.empty((2 * np2 - 1) * 32, dtype=torch.uint8).pin_memory()
    n_local = rows_local * N_PER_ROW
    e2e_steps = max(1, min(steps, 20))

    def e2e_step():
        _lib.check(lib.lcpc_commit_host(enc.plan, h_coeffs.data_ptr(), n_local, None, h_comm.data_ptr(),
                                        h_hashes.data_ptr(), None))

    if world == 1:
        for _ in range(3):
            e2e_step()
        barrier()
        sampler.start()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        sampler.stop()
        e2e = {"value": n_local * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": n_local * 8,
               "d2h_bytes_per_step": ROWS_PER_GPU * N_COLS * 8 + (2 * np2 - 1) * 32, "ms_per_step": 1e3 * dt / e2e_steps,
               "steps": e2e_steps, "api": "lcpc_commit_host (pinned host coeffs in; LcCommit.comm + LcCommit.hashes out)"}
        assert bytes(h_hashes[-32:].numpy()).hex() == gpu_root, "e2e root differs from the device-resident root"
        # informational: the same host call keeping the commitment resident in HBM (what a server that answers
        # openings / folds from the handle needs): coefficients in, 32-byte root out
        keep = C.c_void_p()
        root_buf = torch.empty(32, dtype=torch.uint8).pin_memory()

        def resident_step():
            _lib.check(lib.lcpc_commit_host(enc.plan, h_coeffs.data_ptr(), n_local, None, None, None, C.byref(keep)))
            _lib.check(lib.lcpc_commit_root(keep, root_buf.data_ptr()))
            lib.lcpc_commit_free(keep)

        for _ in range(2):
            resident_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            resident_step()
        torch.cuda.synchronize()
        dt2 = time.perf_counter() - t0
        e2e["shape"] = "full LcCommit out (comm + hashes)"
        e2e["root_only"] = {"value": n_local * e2e_steps / dt2, "unit": UNIT, "ms_per_step": 1e3 * dt2 / e2e_steps,
                            "h2d_bytes_per_step": n_local * 8, "d2h_bytes_per_step": 32,
                            "api": "lcpc_commit_host(keep=handle) + lcpc_commit_root: the commitment stays resident in HBM"}
        assert bytes(root_buf.numpy()).hex() == gpu_root
    else:
        # sharded end to end, both shapes: pinned host row shards in; (a) the full LcCommit out -- every rank's encoded rows
        # and Merkle subtree to its own pinned host buffers, the top of the tree on rank 0 -- and (b) the root only
        by_rows = sc.hashing == "rows"
        h_comm = torch.empty((rows_local if by_rows else n_rows_total) * (N_COLS if by_rows else sc.cb), dtype=torch.int64).pin_memory()
        h_sub = torch.empty((2 * sc.cb - 1) * 32, dtype=torch.uint8).pin_memory()
        h_top = torch.empty((2 * world - 1) * 32, dtype=torch.uint8).pin_memory()

        def e2e_sharded_step(full: bool):
            if by_rows:  # PCIe in, encode and PCIe out overlapped over row chunks (ShardedLigeroCommitter.commit_host)
                sc.commit_host(h_coeffs, h_comm if full else None)
            else:
                sc.commit(h_coeffs.cuda(non_blocking=True))
                if full:
                    h_comm.copy_(sc.comm_cols, non_blocking=True)
            if full:
                h_sub.copy_(sc.subtree, non_blocking=True)
                if rank == 0:
                    h_top.copy_(sc.top, non_blocking=True)
                if by_rows:
                    sc.wait_host_copies()
                torch.cuda.synchronize()
            elif rank == 0:
                sc.root()  # device -> host read of the result

        def time_e2e(full: bool) -> float:
            for _ in range(3):  # untimed, like the one-GPU leg: first-use allocations of the staging tensors
                e2e_sharded_step(full)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_sharded_step(full)
            barrier()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        dt_full = time_e2e(True)
        if rank == 0:
            assert bytes(h_top[-32:].numpy()).hex() == gpu_root, "e2e root differs from the device-resident root"
        dt_root = time_e2e(False)
        d2h_full = n_rows_total * N_COLS * 8 + world * (2 * sc.cb - 1) * 32 + (2 * world - 1) * 32
        e2e = {"value": n_total * e2e_steps / dt_full, "unit": UNIT, "h2d_bytes_per_step": n_local * 8 * world,
               "d2h_bytes_per_step": d2h_full, "ms_per_step": 1e3 * dt_full / e2e_steps, "steps": e2e_steps,
               "shape": "full LcCommit out (comm + hashes)",
               "api": "ShardedLigeroCommitter.commit_host (pinned host row shards in; every rank's encoded rows and Merkle subtree "
                      "to its pinned host buffers, top of the tree on rank 0; PCIe in / encode / PCIe out overlapped over row chunks)",
               "root_only": {"value": n_total * e2e_steps / dt_root, "unit": UNIT, "ms_per_step": 1e3 * dt_root / e2e_steps,
                             "h2d_bytes_per_step": n_local * 8 * world, "d2h_bytes_per_step": 32,
                             "api": "ShardedLigeroCommitter.commit_host (pinned host row shards in; Merkle root out on rank 0)"}}

    # ---- BASELINE configs[3] on the N GPUs (N > 1): 4 GiB proof-of-storage file, commit + retrievability proof + verify
    pos_cfg = None
    if world > 1 and not args.no_configs:
        del h_comm
        sc_fused, sc_cv_fused, sc_hashing = sc.fused, sc.cv_fused, sc.hashing
        sc = None  # release the headline workload's buffers (symmetric memory stays mapped until exit)
        torch.cuda.empty_cache()
        pos_cfg = run_pos_config(torch, dist, P, ctx, stream, rank, world, measured_peak_gbs()[0])

    elif world > 1:
        sc_fused, sc_cv_fused, sc_hashing = sc.fused, sc.cv_fused, sc.hashing

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event timings from the timed region) ----------
    peak, peak_src = measured_peak_gbs()
    roofline = None
    if kt:
        dom = max(kt.items(), key=lambda kv: kv[1][1])
        name, (count, total_ms) = dom
        per_launch_ms = total_ms / count
        alg = kernel_algorithmic_bytes(name, rows_local)
        ach = alg / (per_launch_ms * 1e-3) / 1e9 if alg else None
        traffic, traffic_src = dram_traffic(name) if world == 1 else (None, None)
        roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": (ach / peak) if ach else None,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg, "ms_per_launch": per_launch_ms,
                    "share_of_step": total_ms / ms_total_timed,
                    "ms_per_step_with_launch_events": ms_total_timed / steps,
                    "kernels_ms_per_step": {k: v[1] / steps for k, v in kt.items()},
                    "note": "integer-pipe bound (64-bit Montgomery + BLAKE3 ARX on 32-bit IMAD/ALU), see DESIGN.md"}
        # second roofline: the integer pipes (what actually binds these kernels): essential instruction counts of the
        # arithmetic against the issue rates MEASURED on this device just now (lcpc_ctx_measure_int_pipes)
        try:
            n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
            integer_peak = ctx.measure_int_pipes()
            integer_peak["n_sm"] = n_sm
            integer_peak["note"] = ("SM sub-partition cycles per warp instruction, all resident warps issuing: pure IMAD.WIDE / "
                                    "IMAD / LOP3 streams and an IMAD + LOP3 mix (two pipes at once)")
            roofline["integer_pipe"] = integer_pipe_report(roofline["kernels_ms_per_step"], n_sm, sampler.summary()["sm_mhz"],
                                                           integer_peak["cycles_per_warp_instr"])
            roofline["integer_pipe_note"] = ("floor = essential IMAD.WIDE / IMAD / ALU instruction counts of the arithmetic "
                                             "(DESIGN.md section 3) x the measured cycles per instruction (integer_peak) at the "
                                             "sampled SM clock; frac = floor / measured")
        except Exception as e:  # the model is commentary: it must never cost the line
            integer_peak = None
            roofline["integer_pipe"] = {"error": repr(e)}
    step_gbs = algorithmic_bytes(n_total, n_rows_total) / (ms_per_step * 1e-3) / 1e9

    # ---- the other BASELINE.json configurations (N = 1) ----------------------------------------------
    configs = {"pos_4gib": pos_cfg} if pos_cfg is not None else None
    if pos_cfg is not None:
        assert pos_cfg["matches_known_answer"] is not False and pos_cfg["verified"], f"configs[3] differs from its known answer: {pos_cfg}"
    if world == 1 and not args.no_configs:
        del d_comm, d_hashes, h_comm, h_hashes
        torch.cuda.empty_cache()
        configs = run_configs(torch, P, lib, _lib, ctx, stream, peak, d_coeffs)
        bad = [k for k, v in configs.items() if v["matches_known_answer"] is False]
        assert not bad, f"configs differ from tests/golden/bench_roots.json: {bad}"

    # ---- CPU baseline: oracle on the same input, all host threads; also the parity gate ------------
    # (rank 0 at N = 1 only: at N > 1 the other ranks have left and the line carries no CPU leg)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import lcpc_oracle as O

        O.build()
        O.set_threads(os.cpu_count() or 1)
        oenc = O.LigeroEncoding(FID, N_PER_ROW, N_COLS)
        sample_rows = ROWS_PER_GPU * (world if world <= 2 else 1)
        if world == 1 or world == 2:
            cpu_in = make_coeffs(2, n_total)
        else:
            cpu_in = h_coeffs_np
        oc = O.commit(cpu_in, oenc)  # warm-up (page faults of the 256 MiB encoded matrix, thread pool start)
        reps, t0 = 0, time.perf_counter()
        while reps < 20 and (reps < 3 or time.perf_counter() - t0 < 10.0):  # a bounded ~10 s sample, at least 3 commits
            oc = O.commit(cpu_in, oenc)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": cpu_in.shape[0] / dt, "unit": UNIT, "cores": O.max_threads(), "kind": "port",
               "sample": f"{reps} commits of {sample_rows} rows x {N_PER_ROW} -> {N_COLS} after one warm-up ({dt:.3f} s each), "
                         f"C restatement of the reference algorithm with its rayon decomposition as OpenMP",
               "root_matches_gpu": (oc.get_root().hex() == gpu_root) if world <= 2 else None}
        if world <= 2:
            assert oc.get_root().hex() == gpu_root, "GPU Merkle root differs from the CPU oracle"

    # known-answer root of this workload at N GPUs (tests/golden/bench_roots.json, computed once by the CPU oracle):
    # the parity gate of the multi-GPU lines, where no CPU leg runs
    kat = expected_root(world)
    if kat is not None:
        assert gpu_root == kat, f"GPU Merkle root at {world} GPUs differs from tests/golden/bench_roots.json"

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 (63-bit prime field, Montgomery)", "data": "synthetic",
        "config": workload_config(world),
        "parallelism": "single GPU" if world == 1 else (
                       f"row shards x{world}; " + ("encode kernel stores into peer column blocks over NVLink (symmetric memory); commit k hashed "
                                                    "after commit k+1's encode is issued, all K finished inside the timed region"
                                                   if sc_fused else ("BLAKE3 chunk chaining values hashed where the rows are, "
                                                                     + ("stored by the hash kernel into the owners' stores over NVLink"
                                                                        if sc_cv_fused else "NCCL all-to-all of 32 B per chunk and column")
                                                                     if sc_hashing == "rows" else "NCCL all-to-all"))
                       + "; per-rank Merkle subtrees, roots all-gathered"),
        "algorithmic_GBps": step_gbs, "frac_of_hbm_peak_whole_commit": step_gbs / peak,
        "single_commit_latency_ms": latency_ms,
        "host_binding": numa,
        "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
        "root": gpu_root, "root_matches_golden": (gpu_root == kat) if kat is not None else None,
        "integer_peak": integer_peak if kt else None, "configs": configs, "parity_checks": parity_checks,
    }
    _emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
