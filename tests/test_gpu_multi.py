"""One process, several devices behind the C ABI (lcpc_ctx_create_multi, csrc/lcpc_multi.cu): commitments made through a
multi-device context are bit-identical to single-device ones -- root, tree, encoded matrix, folds, openings, leaves,
proofs -- for Ligero and Brakedown plans, field elements and file bytes.

A device may be listed more than once, so the sharded path (chunk-aligned row blocks, chaining values scattered into the
owners' stores, per-device subtrees, the join) also runs on a one-GPU box: [0, 0] and [0, 0, 0, 0].  With two or more GPUs
the same cases run across real peers as well."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import lcpc_proof_of_storage_b200 as pkg

    return pkg


def _device_sets():
    import torch

    sets = [[0, 0], [0, 0, 0, 0]]
    n = torch.cuda.device_count()
    if n >= 2:
        sets.append([0, 1])
    if n >= 4:
        sets.append([0, 1, 2, 3])
    if n >= 8:
        sets.append(list(range(8)))
    return sets


def _same_commit(a, b):
    assert a.get_root() == b.get_root()
    assert np.array_equal(a.hashes, b.hashes)
    assert np.array_equal(a.comm, b.comm)
    assert np.array_equal(a.coeffs, b.coeffs)


@pytest.mark.parametrize("fid,n_rows,n_per_row,n_cols", [
    (0, 1024, 512, 1024),       # 9 chunks per leaf
    (0, 700, 256, 512),         # ragged: 6 chunks, uneven row blocks
    (3, 300, 64, 128),          # 32-byte elements, 10 chunks
    (4, 260, 64, 128),          # Ft253_192: big-endian repr
    (1, 500, 128, 256),         # 16-byte elements
])
def test_multi_device_ligero_commit_fold_open(P, oracle, fid, n_rows, n_per_row, n_cols):
    O = oracle
    L = O.LIMBS[fid]
    coeffs = O.random_field_elements(fid, 123, n_rows * n_per_row - 7)
    ref = P.LcCommit.commit(coeffs, P.LigeroEncoding(fid, n_per_row, n_cols))
    oref = O.commit(coeffs, O.LigeroEncoding(fid, n_per_row, n_cols))
    assert ref.get_root() == oref.get_root()
    tensors = O.random_field_elements(fid, 5, 3 * n_rows).reshape(3, n_rows, L)
    cols = [0, 1, n_cols - 1, n_cols // 2, n_cols // 2 - 1, 77 % n_cols, 0]
    for devices in _device_sets():
        ctx = P.Context.multi(devices)
        assert ctx.n_devices == len(devices)
        enc = P.LigeroEncoding(fid, n_per_row, n_cols, ctx=ctx)
        got = P.LcCommit.commit(coeffs, enc)
        _same_commit(got, ref)
        # the resident (sharded) handle: root, download, folds over both matrices, openings, leaves
        res = P.LcCommit.commit(coeffs, enc, download=False)
        assert res.get_root() == ref.get_root()
        assert np.array_equal(res.fold(tensors), ref.fold(tensors))
        assert np.array_equal(res.fold(tensors, encoded=True), ref.fold(tensors, encoded=True))
        for a, b in zip(res.open_columns(cols), ref.open_columns(cols)):
            assert np.array_equal(a.col, b.col) and np.array_equal(a.path, b.path)
        for a, b in zip(res.open_columns(cols, with_path=False), ref.open_columns(cols, with_path=False)):
            assert np.array_equal(a.col, b.col)
        assert np.array_equal(res.leaves(cols), ref.leaves(cols))
        assert np.array_equal(res.hashes, ref.hashes) and np.array_equal(res.comm, ref.comm)
        with pytest.raises(P.ProverError):
            res.open_columns([n_cols])
        res.close()


def test_multi_device_prove_verify(P, oracle):
    """prove on a sharded commitment gives the proof of the single-device commitment, and it verifies."""
    O = oracle
    fid, n_rows, n_per_row, n_cols = 0, 600, 256, 512
    coeffs = O.random_field_elements(fid, 321, n_rows * n_per_row)
    p = O.MODULUS[fid]
    x = 987654321
    outer = O.to_mont(fid, [pow(x, n_per_row * i, p) for i in range(n_rows)])
    inner = O.to_mont(fid, [pow(x, j, p) for j in range(n_per_row)])
    enc1 = P.LigeroEncoding(fid, n_per_row, n_cols)
    c1 = P.LcCommit.commit(coeffs, enc1)
    root = c1.get_root()

    def start():
        tr = P.Transcript(b"multi")
        tr.append_message(b"polycommit", root)
        return tr

    pf1 = c1.prove(outer, enc1, start())
    for devices in _device_sets()[:3]:
        enc = P.LigeroEncoding(fid, n_per_row, n_cols, ctx=P.Context.multi(devices))
        c = P.LcCommit.commit(coeffs, enc, download=False)
        pf = c.prove(outer, enc, start())
        assert np.array_equal(pf.p_eval, pf1.p_eval)
        assert all(np.array_equal(a, b) for a, b in zip(pf.p_random_vec, pf1.p_random_vec))
        assert all(np.array_equal(a.col, b.col) and np.array_equal(a.path, b.path) for a, b in zip(pf.columns, pf1.columns))
        res = pf.verify(root, outer, inner, enc, start())     # the verifier runs on the context's first device
        assert np.array_equal(res, pf1.verify(root, outer, inner, enc1, start()))


def test_multi_device_brakedown_and_bytes(P, oracle):
    O = oracle
    # Brakedown: non power-of-two n_cols, padding leaves in the last column block(s)
    fid, n_per_row, n_rows = 0, 3000, 520
    coeffs = O.random_field_elements(fid, 11, n_rows * n_per_row - 100)
    ref = P.LcCommit.commit(coeffs, P.SdigEncoding.new_from_dims(fid, n_per_row, None, seed=0))
    data = bytes(np.random.default_rng(3).integers(0, 256, size=7 * 256 * 700 - 5, dtype=np.uint8))
    ref_b = P.LcCommit.commit_bytes(data, P.LigeroEncoding(0, 256, 512))
    for devices in _device_sets():
        ctx = P.Context.multi(devices)
        got = P.LcCommit.commit(coeffs, P.SdigEncoding.new_from_dims(fid, n_per_row, None, seed=0, ctx=ctx))
        _same_commit(got, ref)
        got_b = P.LcCommit.commit_bytes(data, P.LigeroEncoding(0, 256, 512, ctx=ctx))
        _same_commit(got_b, ref_b)


def test_multi_device_small_and_unsupported_shapes_run_on_the_first_device(P, oracle):
    """Single-chunk leaves, fewer chunks than devices and 24-byte elements: same results, through the first device."""
    O = oracle
    ctx = P.Context.multi([0, 0, 0, 0])
    for fid, n_rows, n_per_row, n_cols in [(0, 20, 64, 128), (0, 300, 64, 128), (2, 400, 64, 128)]:
        coeffs = O.random_field_elements(fid, 9, n_rows * n_per_row)
        ref = P.LcCommit.commit(coeffs, P.LigeroEncoding(fid, n_per_row, n_cols))
        got = P.LcCommit.commit(coeffs, P.LigeroEncoding(fid, n_per_row, n_cols, ctx=ctx))
        _same_commit(got, ref)
    enc = P.LigeroEncoding(0, 64, 128, ctx=ctx)
    rows = O.random_field_elements(0, 4, 2 * 128).reshape(2, 128, 1)
    got = rows.copy()
    enc.encode(got)
    assert np.array_equal(got, O.fft_io(0, rows))
    with pytest.raises(P.LcpcError):
        P.Context.multi([0, 0, 0])        # not a power of two


def test_multi_device_commit_from_a_c_client(tmp_path, oracle):
    """The boundary itself: a C11 program linked against the library makes a two-shard context, commits, opens columns and
    folds through include/lcpc_b200.h, and prints the root; compared with the oracle's."""
    O = oracle
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    import torch

    root_dir = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root_dir, "lcpc_proof_of_storage_b200", "_lib")
    second = 1 if torch.cuda.device_count() >= 2 else 0
    n_rows, n_per_row, n_cols = 600, 256, 512
    coeffs = O.random_field_elements(0, 77, n_rows * n_per_row)
    coeffs.tofile(tmp_path / "coeffs.bin")
    src = tmp_path / "client.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include "lcpc_b200.h"
#define CHECK(x) do { int32_t rc_ = (x); if (rc_ != LCPC_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, (int)rc_, lcpc_last_error()); return 1; } } while (0)
int main(int argc, char **argv) {
    const size_t n_rows = 600, npr = 256, n_cols = 512, n = n_rows * npr;
    if (argc < 3) return 9;
    int32_t devices[2] = {0, atoi(argv[2])};
    uint64_t *coeffs = malloc(n * 8), *col = malloc(2 * n_rows * 8), *fold = malloc(npr * 8), *tensor = calloc(n_rows, 8);
    uint8_t root[32], paths[2 * 9 * 32];
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(coeffs, 8, n, f) != n) return 2;
    fclose(f);
    lcpc_ctx *ctx; lcpc_plan *plan; lcpc_commit *c;
    CHECK(lcpc_ctx_create_multi(devices, 2, &ctx));
    if (lcpc_ctx_device_count(ctx) != 2) return 3;
    CHECK(lcpc_plan_ligero(ctx, LCPC_FT63, npr, n_cols, NULL, &plan));
    CHECK(lcpc_commit_host(plan, coeffs, n, NULL, NULL, NULL, &c));
    CHECK(lcpc_commit_root(c, root));
    uint64_t cols[2] = {5, 300};
    CHECK(lcpc_open_columns_host(c, cols, 2, col, paths));
    uint64_t one[1];
    CHECK(lcpc_field_constants(LCPC_FT63, NULL, one, NULL, NULL, NULL));
    tensor[3] = one[0];                       /* unit tensor: the fold is coefficient row 3 */
    CHECK(lcpc_fold_host(c, 0, tensor, 1, fold));
    for (size_t j = 0; j < npr; j++) if (fold[j] != coeffs[3 * npr + j]) return 4;
    for (int i = 0; i < 32; i++) printf("%02x", root[i]);
    printf(" %llu %llu\n", (unsigned long long)col[0], (unsigned long long)col[n_rows]);
    lcpc_commit_free(c); lcpc_plan_destroy(plan); lcpc_ctx_destroy(ctx);
    return 0;
}
''')
    exe = tmp_path / "client"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root_dir, "include"), str(src), "-o",
                    str(exe), "-L", lib_dir, "-llcpc_b200", f"-Wl,-rpath,{lib_dir}"], check=True)
    res = subprocess.run([str(exe), str(tmp_path / "coeffs.bin"), str(second)], capture_output=True, text=True)
    assert res.returncode == 0, (res.returncode, res.stdout, res.stderr)
    exp = O.commit(coeffs, O.LigeroEncoding(0, n_per_row, n_cols))
    root_hex, c0, c1 = res.stdout.split()
    assert root_hex == exp.get_root().hex()
    assert int(c0) == int(exp.comm[0, 5, 0]) and int(c1) == int(exp.comm[0, 300, 0])
