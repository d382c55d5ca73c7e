// One process, N devices behind the same C ABI: a commitment made through a context from
// lcpc_ctx_create_multi is sharded over the context's devices inside the library --
//
//   rows      device g owns a block of whole rows of the coefficient matrix, cut on BLAKE3 chunk boundaries of the column
//             leaves, and encodes them (lcpc-2d/src/lib.rs:677-682: rows are independent; no communication);
//   hashing   every device hashes the chunks of ALL columns that its rows make up (k_hash_chunks_scatter) and stores each
//             32-byte chaining value straight into the chaining-value store of the device that owns that column's block
//             of the padded leaf range -- its own HBM or a peer's over NVLink (peer access is enabled at context
//             creation): 3 % of the encoded matrix for 8-byte elements; the encoded matrix never moves;
//   tree      each device merges its columns' chaining values into leaves and builds its power-of-two Merkle subtree
//             (k_merge_tree, one launch); the subtree roots are copied to the first device, which computes the top
//             log2(N) levels (lib.rs:777-815);
//   fold      partial collapse_columns over the local rows on every device, partials copied to the first device and summed
//             mod p there (lib.rs:1126-1154);
//   open      column values gathered from every device's rows straight into the caller's buffer (2-D copies), the path
//             from the owner's subtree plus the siblings among the subtree roots (lib.rs:818-855).
//
// Cross-device ordering is by CUDA events between the devices' streams; nothing is synchronised through the host until
// the results are copied out.  The result is bit-identical to the single-device commitment (tests/test_gpu_multi.py).
// A device may be listed more than once: its shards then share the device (how the path is tested on a one-GPU box).
#include <algorithm>
#include <cmath>
#include <thread>

#include "lcpc_handles.h"

using namespace lcpc;
using namespace lcpc::abi;

namespace {

struct Partition {
    std::vector<size_t> row0, rows;
    std::vector<uint64_t> c0, c1;
};

// Row blocks whose boundaries are chunk boundaries of the column leaves: the leaf stream is 32 zero bytes followed by w
// bytes per row, cut into 1024-byte chunks, so chunk 0 holds the first (1024 - 32) / w rows and every later chunk
// 1024 / w.  Device q owns the chunks between the boundaries nearest to the even split q * n_rows / n.  Empty when
// elements straddle chunk boundaries (24-byte elements) or the leaf is a single chunk.
bool chunk_row_partition(int limbs, size_t n_rows, size_t n_dev, Partition &out) {
    const uint64_t w = 8ull * (uint64_t)limbs;
    if (1024 % w) return false;
    const uint64_t n_chunks = (32 + (uint64_t)n_rows * w + 1023) / 1024;
    if (n_chunks < 2) return false;
    const uint64_t first = (1024 - 32) / w, per = 1024 / w;
    auto start = [&](uint64_t c) -> uint64_t { return c == 0 ? 0 : std::min<uint64_t>(n_rows, first + (c - 1) * per); };
    std::vector<uint64_t> bounds{0};
    for (size_t q = 1; q < n_dev; q++) {
        const double target = (double)q * (double)n_rows / (double)n_dev;
        uint64_t best = bounds.back();
        double best_d = std::fabs((double)start(best) - target);
        for (uint64_t k = bounds.back() + 1; k <= n_chunks; k++) {
            const double d = std::fabs((double)start(k) - target);
            if (d < best_d) {
                best_d = d;
                best = k;
            } else if ((double)start(k) > target) {
                break;
            }
        }
        bounds.push_back(best);
    }
    bounds.push_back(n_chunks);
    out = Partition{};
    for (size_t q = 0; q < n_dev; q++) {
        out.c0.push_back(bounds[q]);
        out.c1.push_back(bounds[q + 1]);
        out.row0.push_back((size_t)start(bounds[q]));
        out.rows.push_back((size_t)(start(bounds[q + 1]) - start(bounds[q])));
    }
    return true;
}

// run fn(g) for every device on its own host thread (pageable host copies block the issuing thread: one thread per
// device keeps all PCIe links busy); returns the first failure, with its message re-posted on the calling thread
template <class Fn>
int32_t for_each_device(size_t n, Fn fn) {
    std::vector<int32_t> rc(n, LCPC_OK);
    std::vector<std::string> msg(n);
    std::vector<std::thread> th;
    th.reserve(n);
    for (size_t g = 0; g < n; g++)
        th.emplace_back([&, g]() {
            rc[g] = fn(g);
            if (rc[g] != LCPC_OK) msg[g] = last_error();
        });
    for (auto &t : th) t.join();
    for (size_t g = 0; g < n; g++)
        if (rc[g] != LCPC_OK) return fail(rc[g], msg[g]);
    return LCPC_OK;
}

int log2_exact(size_t v) {
    int l = 0;
    while (((size_t)1 << l) < v) l++;
    return l;
}

size_t level_off(size_t n_leaves, int l) { return l == 0 ? 0 : 2 * n_leaves - (n_leaves >> (l - 1)); }

int32_t sync_all(lcpc_ctx *ctx) {
    for (lcpc_ctx *s : ctx->subs) {
        CU(cudaSetDevice(s->device));
        CU(cudaStreamSynchronize(s->stream));
    }
    return LCPC_OK;
}

// flat tree [np2 | np2/2 | ... | 1] from the devices' subtrees and the top levels, into host memory
int32_t assemble_hashes(lcpc_commit *c, uint8_t *hashes_out) {
    lcpc_ctx *ctx = c->plan->ctx;
    const size_t W = c->shards.size(), cb = c->cb, np2 = c->np2;
    const int depth_sub = log2_exact(cb);
    for (size_t h = 0; h < W; h++) {
        lcpc_ctx *s = ctx->subs[h];
        CU(cudaSetDevice(s->device));
        for (int l = 0; l <= depth_sub; l++) {
            const size_t n = cb >> l;
            CU(cudaMemcpyAsync(hashes_out + (level_off(np2, l) + h * n) * 32, c->shards[h].d_subtree + level_off(cb, l) * 32, n * 32,
                               cudaMemcpyDeviceToHost, s->stream));
        }
    }
    if (W > 1) {  // levels above the subtree roots (the roots themselves are level depth_sub, already copied)
        lcpc_ctx *s0 = ctx->subs[0];
        CU(cudaSetDevice(s0->device));
        CU(cudaMemcpyAsync(hashes_out + level_off(np2, depth_sub + 1) * 32, c->d_top + W * 32, (W - 1) * 32, cudaMemcpyDeviceToHost,
                           s0->stream));
    }
    return LCPC_OK;
}

}  // namespace

namespace lcpc {
namespace abi {
namespace multi {

bool usable(const lcpc_plan *plan, size_t n_rows) {
    if (!plan || plan->subs.size() < 2) return false;
    Partition p;
    if (!chunk_row_partition(limbs_of(plan->fid), n_rows, plan->subs.size(), p)) return false;
    for (size_t r : p.rows)
        if (r == 0) return false;  // fewer chunks than devices: not worth sharding, the first device takes it
    return true;
}

void release(lcpc_commit *c) {
    if (!c || c->shards.empty()) return;
    lcpc_ctx *ctx = c->plan->ctx;
    for (size_t g = 0; g < c->shards.size(); g++) {
        lcpc_ctx *s = ctx->subs[g];
        lcpc_shard &sh = c->shards[g];
        cudaSetDevice(s->device);
        if (sh.d_coeffs) cudaFreeAsync(sh.d_coeffs, s->stream);
        if (sh.d_comm) cudaFreeAsync(sh.d_comm, s->stream);
        if (sh.d_cvs) cudaFreeAsync(sh.d_cvs, s->stream);
        if (sh.d_subtree) cudaFreeAsync(sh.d_subtree, s->stream);
        if (sh.ev) cudaEventDestroy(sh.ev);
        if (sh.ev_tree) cudaEventDestroy(sh.ev_tree);
    }
    if (c->d_top) {
        cudaSetDevice(ctx->subs[0]->device);
        cudaFreeAsync(c->d_top, ctx->subs[0]->stream);
    }
    c->shards.clear();
    c->d_top = nullptr;
}

// coeffs != null: field elements (n_coeffs of them); else file_bytes (n_bytes, WriteableFt63 7-byte packing)
int32_t commit_host(lcpc_plan *plan, lcpc_commit *c, const uint64_t *coeffs, size_t n_coeffs, const uint8_t *file_bytes,
                    size_t n_bytes, uint64_t *coeffs_out, uint64_t *comm_out, uint8_t *hashes_out) {
    lcpc_ctx *ctx = plan->ctx;
    const size_t W = plan->subs.size();
    const int fid = plan->fid, L = limbs_of(fid);
    const size_t w = (size_t)L * 8, npr = c->n_per_row, n_cols = c->n_cols, n_rows = c->n_rows, np2 = c->np2;
    Partition part;
    if (!chunk_row_partition(L, n_rows, W, part)) return fail(LCPC_ERR_DIMS, "no chunk-aligned row partition for this shape");
    if (np2 < W) return fail(LCPC_ERR_DIMS, "fewer leaves than devices");
    const size_t cb = np2 / W;
    const uint64_t nc = hash_leaf_chunks(fid, n_rows), total = hash_leaf_bytes(fid, n_rows);
    c->cb = cb;
    c->n_chunks = nc;
    c->shards.assign(W, lcpc_shard{});
    CvScatter sc{};
    sc.log_cb = log2_exact(cb);
    // ---- allocations on every device (stream-ordered); the stores must exist before any device scatters into them
    for (size_t g = 0; g < W; g++) {
        lcpc_ctx *s = ctx->subs[g];
        lcpc_shard &sh = c->shards[g];
        sh.row0 = part.row0[g];
        sh.rows = part.rows[g];
        sh.c0 = part.c0[g];
        sh.c1 = part.c1[g];
        sh.cols_local = n_cols > g * cb ? std::min(cb, n_cols - g * cb) : 0;
        CU(cudaSetDevice(s->device));
        CU(cudaEventCreateWithFlags(&sh.ev, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&sh.ev_tree, cudaEventDisableTiming));
        CU(cudaMallocAsync((void **)&sh.d_coeffs, std::max<size_t>(1, sh.rows * npr) * w, s->stream));
        CU(cudaMallocAsync((void **)&sh.d_comm, std::max<size_t>(1, sh.rows * n_cols) * w, s->stream));
        CU(cudaMallocAsync((void **)&sh.d_cvs, (size_t)nc * cb * 32, s->stream));
        CU(cudaMallocAsync((void **)&sh.d_subtree, (2 * cb - 1) * 32, s->stream));
        CU(cudaEventRecord(sh.ev, s->stream));
        sc.base[g] = reinterpret_cast<uint32_t *>(sh.d_cvs);
    }
    // ---- per device: rows in, encode, encoded rows out, chunk hashing with the exchange fused in
    int32_t rc = for_each_device(W, [&](size_t g) -> int32_t {
        lcpc_ctx *s = ctx->subs[g];
        lcpc_shard &sh = c->shards[g];
        CU(cudaSetDevice(s->device));
        const size_t e0 = sh.row0 * npr, e1 = (sh.row0 + sh.rows) * npr;  // my elements of the padded coefficient matrix
        if (coeffs) {
            const size_t have = n_coeffs > e0 ? std::min(n_coeffs, e1) - e0 : 0;
            if (have) CU(cudaMemcpyAsync(sh.d_coeffs, coeffs + e0 * L, have * w, cudaMemcpyHostToDevice, s->stream));
            if (have < e1 - e0)  // lib.rs:665-674: zero fill of the ragged tail
                CU(cudaMemsetAsync(sh.d_coeffs + have * L, 0, (e1 - e0 - have) * w, s->stream));
        } else {
            // DataField::from_byte_vec on my slice of the file: element k = bytes [7k, 7k + 7)
            const size_t b0 = std::min(n_bytes, e0 * 7), b1 = std::min(n_bytes, e1 * 7);
            const size_t n_el = (b1 - b0 + 6) / 7;
            if (b1 > b0) {
                DevBuf raw;
                CU(raw.alloc(b1 - b0, s->stream));
                CU(cudaMemcpyAsync(raw.p, file_bytes + b0, b1 - b0, cudaMemcpyHostToDevice, s->stream));
                CU(pack_bytes7(raw.as<uint8_t>(), b1 - b0, sh.d_coeffs, s->lc()));
            }
            if (n_el < e1 - e0) CU(cudaMemsetAsync(sh.d_coeffs + n_el, 0, (e1 - e0 - n_el) * w, s->stream));
        }
        int32_t r = encode_dev(plan->subs[g], sh.d_coeffs, sh.rows, sh.d_comm);
        if (r != LCPC_OK) return r;
        if (coeffs_out) CU(cudaMemcpyAsync(coeffs_out + e0 * L, sh.d_coeffs, (e1 - e0) * w, cudaMemcpyDeviceToHost, s->stream));
        // every store is allocated (stream order of its device) before my kernel writes into it
        for (size_t h = 0; h < W; h++)
            if (h != g) CU(cudaStreamWaitEvent(s->stream, c->shards[h].ev, 0));
        CU(hash_chunk_range_scatter(fid, sh.d_comm, (int64_t)sh.row0, n_rows, n_cols, n_cols, sh.c0, sh.c1, total, nc, sc, s->lc()));
        return LCPC_OK;
    });
    if (rc != LCPC_OK) return rc;
    // ---- exchange complete on device h when every device's scatter kernel has finished
    for (size_t g = 0; g < W; g++) {
        CU(cudaSetDevice(ctx->subs[g]->device));
        CU(cudaEventRecord(c->shards[g].ev, ctx->subs[g]->stream));
    }
    for (size_t h = 0; h < W; h++) {
        lcpc_ctx *s = ctx->subs[h];
        lcpc_shard &sh = c->shards[h];
        CU(cudaSetDevice(s->device));
        for (size_t g = 0; g < W; g++)
            if (g != h) CU(cudaStreamWaitEvent(s->stream, c->shards[g].ev, 0));
        unsigned *tk = nullptr;
        CU(s->tickets(1, &tk));
        CU(merge_tree(sh.d_cvs, sh.cols_local, nc, sh.d_subtree, cb, tk, s->lc(), cb));
        CU(cudaEventRecord(sh.ev_tree, s->stream));  // the join below waits for this, not for the host copy behind it
        // the encoded rows leave while the top of the tree is being built (comm_out is large: PCIe-bound)
        if (comm_out)
            CU(cudaMemcpyAsync(comm_out + sh.row0 * n_cols * L, sh.d_comm, sh.rows * n_cols * w, cudaMemcpyDeviceToHost, s->stream));
    }
    // ---- join: subtree roots to the first device, top log2(W) levels there
    {
        lcpc_ctx *s0 = ctx->subs[0];
        CU(cudaSetDevice(s0->device));
        CU(cudaMallocAsync((void **)&c->d_top, (2 * W - 1) * 32, s0->stream));
        for (size_t h = 0; h < W; h++) {
            if (h != 0) CU(cudaStreamWaitEvent(s0->stream, c->shards[h].ev_tree, 0));
            CU(cudaMemcpyPeerAsync(c->d_top + h * 32, s0->device, c->shards[h].d_subtree + (2 * cb - 2) * 32, ctx->subs[h]->device, 32,
                                   s0->stream));
        }
        CU(merkle_tree(c->d_top, W, s0->lc()));
    }
    if (hashes_out) {
        // device 0's stream holds the top levels; the subtrees are final on their own streams
        int32_t r = assemble_hashes(c, hashes_out);
        if (r != LCPC_OK) return r;
    }
    return sync_all(ctx);
}

int32_t root(lcpc_commit *c, uint8_t *root_out) {
    lcpc_ctx *s0 = c->plan->ctx->subs[0];
    const size_t W = c->shards.size();
    CU(cudaSetDevice(s0->device));
    CU(cudaMemcpyAsync(root_out, c->d_top + (2 * W - 2) * 32, 32, cudaMemcpyDeviceToHost, s0->stream));
    CU(cudaStreamSynchronize(s0->stream));
    return LCPC_OK;
}

int32_t download(lcpc_commit *c, uint64_t *coeffs_out, uint64_t *comm_out, uint8_t *hashes_out) {
    lcpc_ctx *ctx = c->plan->ctx;
    const int L = limbs_of(c->plan->fid);
    const size_t w = (size_t)L * 8;
    for (size_t g = 0; g < c->shards.size(); g++) {
        lcpc_ctx *s = ctx->subs[g];
        lcpc_shard &sh = c->shards[g];
        CU(cudaSetDevice(s->device));
        if (coeffs_out)
            CU(cudaMemcpyAsync(coeffs_out + sh.row0 * c->n_per_row * L, sh.d_coeffs, sh.rows * c->n_per_row * w, cudaMemcpyDeviceToHost, s->stream));
        if (comm_out)
            CU(cudaMemcpyAsync(comm_out + sh.row0 * c->n_cols * L, sh.d_comm, sh.rows * c->n_cols * w, cudaMemcpyDeviceToHost, s->stream));
    }
    if (hashes_out) {
        int32_t r = assemble_hashes(c, hashes_out);
        if (r != LCPC_OK) return r;
    }
    return sync_all(ctx);
}

// out[t][j] = sum_r tensors[t][r] * M[r][j]: partial sums over each device's rows, summed mod p on the first device
int32_t fold_host(lcpc_commit *c, int32_t which, const uint64_t *tensors, size_t n_tensors, uint64_t *out) {
    lcpc_ctx *ctx = c->plan->ctx;
    const size_t W = c->shards.size();
    const int fid = c->plan->fid, L = limbs_of(fid);
    const size_t w = (size_t)L * 8, n_rows = c->n_rows;
    const size_t width = which == 0 ? c->n_per_row : c->n_cols;
    const size_t part_elems = n_tensors * width;
    lcpc_ctx *s0 = ctx->subs[0];
    CU(cudaSetDevice(s0->device));
    DevBuf d_parts, d_sum;
    CU(d_parts.alloc(W * part_elems * w, s0->stream));
    CU(d_sum.alloc(part_elems * w, s0->stream));
    cudaEvent_t ready;
    CU(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    CU(cudaEventRecord(ready, s0->stream));
    std::vector<cudaEvent_t> done(W, nullptr);
    int32_t rc = for_each_device(W, [&](size_t g) -> int32_t {
        lcpc_ctx *s = ctx->subs[g];
        lcpc_shard &sh = c->shards[g];
        CU(cudaSetDevice(s->device));
        CU(cudaEventCreateWithFlags(&done[g], cudaEventDisableTiming));
        DevBuf d_t, d_o, d_s;
        CU(d_t.alloc(n_tensors * sh.rows * w, s->stream));
        CU(d_o.alloc(part_elems * w, s->stream));
        CU(d_s.alloc(fold_scratch_bytes(fid, sh.rows, width, n_tensors), s->stream));
        // my rows of every tensor: [n_tensors][rows] from [n_tensors][n_rows]
        CU(cudaMemcpy2DAsync(d_t.p, sh.rows * w, tensors + sh.row0 * L, n_rows * w, sh.rows * w, n_tensors, cudaMemcpyHostToDevice,
                             s->stream));
        const uint64_t *mat = which == 0 ? sh.d_coeffs : sh.d_comm;
        CU(cudaMemsetAsync(d_o.p, 0, part_elems * w, s->stream));
        CU(fold(fid, mat, sh.rows, width, width, d_t.as<uint64_t>(), n_tensors, d_o.as<uint64_t>(), d_s.as<uint64_t>(), s->lc()));
        CU(cudaStreamWaitEvent(s->stream, ready, 0));  // the gather buffer exists on the first device
        CU(cudaMemcpyPeerAsync(d_parts.as<uint8_t>() + g * part_elems * w, s0->device, d_o.p, s->device, part_elems * w, s->stream));
        CU(cudaEventRecord(done[g], s->stream));
        return LCPC_OK;
    });
    if (rc == LCPC_OK) {
        CU(cudaSetDevice(s0->device));
        for (size_t g = 0; g < W; g++) CU(cudaStreamWaitEvent(s0->stream, done[g], 0));
        CU(add_partials(fid, d_parts.as<uint64_t>(), W, part_elems, d_sum.as<uint64_t>(), s0->lc()));
        CU(cudaMemcpyAsync(out, d_sum.p, part_elems * w, cudaMemcpyDeviceToHost, s0->stream));
        rc = sync_all(ctx);
    }
    for (size_t g = 0; g < W; g++)
        if (done[g]) cudaEventDestroy(done[g]);
    cudaEventDestroy(ready);
    return rc;
}

int32_t open_columns_host(lcpc_commit *c, const uint64_t *cols, size_t n, uint64_t *cols_out, uint8_t *paths_out) {
    lcpc_ctx *ctx = c->plan->ctx;
    const size_t W = c->shards.size(), cb = c->cb;
    const int fid = c->plan->fid, L = limbs_of(fid);
    const size_t w = (size_t)L * 8, n_rows = c->n_rows, n_cols = c->n_cols;
    const int depth = log2_exact(c->np2), depth_sub = log2_exact(cb);
    std::vector<uint64_t> local(n);
    for (size_t i = 0; i < n; i++) local[i] = cols[i] & (cb - 1);
    std::vector<std::vector<uint8_t>> sub_paths(W);
    int32_t rc = for_each_device(W, [&](size_t g) -> int32_t {
        lcpc_ctx *s = ctx->subs[g];
        lcpc_shard &sh = c->shards[g];
        CU(cudaSetDevice(s->device));
        DevBuf d_cols, d_local, d_out, d_paths;
        CU(d_cols.alloc(n * 8, s->stream));
        CU(cudaMemcpyAsync(d_cols.p, cols, n * 8, cudaMemcpyHostToDevice, s->stream));
        if (cols_out && sh.rows) {
            CU(d_out.alloc(n * sh.rows * w, s->stream));
            CU(gather_columns(fid, sh.d_comm, sh.rows, n_cols, d_cols.as<uint64_t>(), n, d_out.as<uint64_t>(), s->lc()));
            // cols_out[i][row0 .. row0 + rows) <- d_out[i][0 .. rows)
            CU(cudaMemcpy2DAsync(cols_out + sh.row0 * L, n_rows * w, d_out.p, sh.rows * w, sh.rows * w, n, cudaMemcpyDeviceToHost,
                                 s->stream));
        }
        if (paths_out && depth_sub > 0) {
            // only the columns this device owns are kept below; the others cost a few hundred bytes each
            sub_paths[g].resize(n * (size_t)depth_sub * 32);
            CU(d_local.alloc(n * 8, s->stream));
            CU(d_paths.alloc(n * (size_t)depth_sub * 32, s->stream));
            CU(cudaMemcpyAsync(d_local.p, local.data(), n * 8, cudaMemcpyHostToDevice, s->stream));
            CU(gather_paths(sh.d_subtree, cb, d_local.as<uint64_t>(), n, d_paths.as<uint8_t>(), s->lc()));
            CU(cudaMemcpyAsync(sub_paths[g].data(), d_paths.p, sub_paths[g].size(), cudaMemcpyDeviceToHost, s->stream));
        }
        CU(cudaStreamSynchronize(s->stream));
        return LCPC_OK;
    });
    if (rc != LCPC_OK) return rc;
    if (paths_out && depth > 0) {
        std::vector<uint8_t> top((2 * W - 1) * 32);
        lcpc_ctx *s0 = ctx->subs[0];
        CU(cudaSetDevice(s0->device));
        CU(cudaMemcpyAsync(top.data(), c->d_top, top.size(), cudaMemcpyDeviceToHost, s0->stream));
        CU(cudaStreamSynchronize(s0->stream));
        for (size_t i = 0; i < n; i++) {
            const size_t owner = (size_t)(cols[i] / cb);
            uint8_t *dst = paths_out + i * (size_t)depth * 32;
            if (depth_sub > 0) std::memcpy(dst, sub_paths[owner].data() + i * (size_t)depth_sub * 32, (size_t)depth_sub * 32);
            size_t node = owner;
            for (int l = 0; l < depth - depth_sub; l++) {  // siblings among / above the subtree roots (lib.rs:841-851)
                std::memcpy(dst + (size_t)(depth_sub + l) * 32, top.data() + (level_off(W, l) + (node ^ 1)) * 32, 32);
                node >>= 1;
            }
        }
    }
    return LCPC_OK;
}

int32_t leaves_host(lcpc_commit *c, const uint64_t *cols, size_t n, uint8_t *leaves_out) {
    lcpc_ctx *ctx = c->plan->ctx;
    const size_t cb = c->cb;
    for (size_t i = 0; i < n; i++) {
        const size_t owner = (size_t)(cols[i] / cb);
        lcpc_ctx *s = ctx->subs[owner];
        CU(cudaSetDevice(s->device));
        CU(cudaMemcpyAsync(leaves_out + i * 32, c->shards[owner].d_subtree + (cols[i] & (cb - 1)) * 32, 32, cudaMemcpyDeviceToHost,
                           s->stream));
    }
    return sync_all(ctx);
}

}  // namespace multi
}  // namespace abi
}  // namespace lcpc
