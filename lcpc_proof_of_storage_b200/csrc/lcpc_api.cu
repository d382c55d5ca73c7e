// C ABI of lcpc_b200 (see include/lcpc_b200.h for the contract and the reference
// items each entry point replaces).  Handles own device memory; host buffers belong to
// the caller.  No CPU fallback: without a CUDA device every call fails loudly.
#include <algorithm>
#include <cstdlib>

#include "lcpc_handles.h"

using namespace lcpc;
using namespace lcpc::abi;

namespace lcpc {
void timer_begin(KernelTimer *t, const char *name, cudaStream_t s) {
    KernelTimer::Rec r{name, t->get(), t->get()};
    cudaEventRecord(r.a, s);
    t->recs.push_back(r);
}
void timer_end(KernelTimer *t, cudaStream_t s) {
    if (!t->recs.empty()) cudaEventRecord(t->recs.back().b, s);
}
namespace abi {
std::string &last_error() {
    thread_local std::string err;
    return err;
}
}  // namespace abi
}  // namespace lcpc

// encode rows already resident: both codes read d_coeffs (stride n_per_row) and write every entry of d_comm
int32_t lcpc::abi::encode_dev(lcpc_plan *plan, const uint64_t *d_coeffs, size_t n_rows, uint64_t *d_comm) {
    lcpc_ctx *ctx = plan->ctx;
    if (plan->kind == 0) {
        CU(ntt_encode(plan->ntt, d_coeffs, plan->n_per_row, plan->n_per_row, d_comm, n_rows, ctx->lc()));
    } else {
        DevBuf tmp;
        CU(tmp.alloc(sdig_tmp_elems(plan->sdig, n_rows) * limbs_of(plan->fid) * sizeof(uint64_t), ctx->stream));
        CU(sdig_encode(plan->sdig, d_coeffs, plan->n_per_row, d_comm, n_rows, tmp.as<uint64_t>(), ctx->lc()));
    }
    return LCPC_OK;
}

namespace {

void free_csr(DevCsr &m) {
    if (m.d_rowptr) cudaFree(m.d_rowptr);
    if (m.d_colidx) cudaFree(m.d_colidx);
    if (m.d_data) cudaFree(m.d_data);
    if (m.d_order) cudaFree(m.d_order);
    m = DevCsr{};
}

// CSC (host, caller's) -> CSR (device)
int32_t upload_csr(const lcpc_csc &a, int fid, int L, DevCsr &out) {
    if (!a.indptr || (a.indptr[a.cols] && (!a.indices || !a.data))) return fail(LCPC_ERR_INVALID_ARG, "null CSC arrays");
    const size_t nnz = (size_t)a.indptr[a.cols];
    if (a.rows > 0xffffffffull || a.cols > 0xffffffffull || nnz > 0xffffffffull)
        return fail(LCPC_ERR_TOO_BIG, "code matrix too large for 32-bit indices");
    std::vector<uint32_t> rowptr(a.rows + 1, 0), colidx(nnz);
    std::vector<uint64_t> data(nnz * (size_t)L);
    for (size_t k = 0; k < nnz; k++) {
        if (a.indices[k] >= a.rows) return fail(LCPC_ERR_INVALID_ARG, "CSC row index out of range");
        rowptr[a.indices[k] + 1]++;
    }
    for (size_t i = 0; i < a.rows; i++) rowptr[i + 1] += rowptr[i];
    std::vector<uint32_t> fill(rowptr.begin(), rowptr.end() - 1);
    for (size_t j = 0; j < a.cols; j++)
        for (uint64_t k = a.indptr[j]; k < a.indptr[j + 1]; k++) {
            const uint32_t dst = fill[a.indices[k]]++;
            colidx[dst] = (uint32_t)j;
            std::memcpy(&data[(size_t)dst * L], &a.data[k * L], L * sizeof(uint64_t));
        }
    // Rows by decreasing length: a warp of k_spmv_t serves 32/gs rows at once and runs as long as the longest of them
    // (the generated codes have 7 - 76 non-zeros per row: 15 % of the lane cycles idle in matrix order), and the longest
    // rows start first.
    std::vector<uint32_t> order(a.rows);
    for (size_t i = 0; i < a.rows; i++) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        return rowptr[x + 1] - rowptr[x] > rowptr[y + 1] - rowptr[y];
    });
    out.rows = a.rows;
    out.cols = a.cols;
    out.nnz = nnz;
    CU(cudaMalloc(&out.d_order, (a.rows ? a.rows : 1) * sizeof(uint32_t)));
    if (a.rows) CU(cudaMemcpy(out.d_order, order.data(), a.rows * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&out.d_rowptr, (a.rows + 1) * sizeof(uint32_t)));
    CU(cudaMalloc(&out.d_colidx, (nnz ? nnz : 1) * sizeof(uint32_t)));
    CU(cudaMalloc(&out.d_data, (nnz ? nnz : 1) * L * sizeof(uint64_t)));
    CU(cudaMemcpy(out.d_rowptr, rowptr.data(), (a.rows + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (nnz) {
        CU(cudaMemcpy(out.d_colidx, colidx.data(), nnz * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(out.d_data, data.data(), nnz * L * sizeof(uint64_t), cudaMemcpyHostToDevice));
        CU(scale_csr_data(fid, out.d_data, nnz, nullptr));
        CU(cudaDeviceSynchronize());
    }
    return LCPC_OK;
}

}  // namespace

// One launch (k_hash_tree) where the grid limits allow, else chunk hashing + merge + tree levels
int32_t lcpc::abi::merkleize_dev(lcpc_ctx *ctx, int fid, const uint64_t *d_comm, size_t n_rows, size_t row_stride, size_t n_cols,
                                 size_t np2, uint8_t *d_hashes, uint8_t **d_cvs_keep) {
    DevBuf scratch;
    uint8_t *cvs = nullptr;
    const size_t cv_bytes = hash_scratch_bytes(fid, n_rows, n_cols);
    if (d_cvs_keep && cv_bytes) {
        CU(cudaMallocAsync((void **)d_cvs_keep, cv_bytes, ctx->stream));
        cvs = *d_cvs_keep;
    } else {
        CU(scratch.alloc(cv_bytes, ctx->stream));
        cvs = scratch.as<uint8_t>();
    }
    // mode 2: everything in one launch (k_hash_tree), chosen when the whole grid is resident at once; mode 1: chunk
    // hashing, then leaf merge + tree in one launch (k_merge_tree); mode 0: the four-launch form.  LCPC_HASH_MODE
    // overrides the choice (tools/bench_hash_tail.py).
    int mode = hash_tree_preferred(fid, n_rows, np2) ? 2 : 1;
    if (const char *e = getenv("LCPC_HASH_MODE")) mode = atoi(e);
    if (mode == 2 && hash_tree_supported(fid, n_rows, np2)) {
        unsigned *tk = nullptr;
        CU(ctx->tickets(hash_tree_tickets(np2), &tk));
        CU(hash_tree(fid, d_comm, n_rows, row_stride, n_cols, np2, d_hashes, cvs, tk, ctx->lc()));
        return LCPC_OK;
    }
    if (mode != 0) {
        unsigned *tk = nullptr;
        CU(ctx->tickets(1, &tk));
        const uint64_t nc = hash_leaf_chunks(fid, n_rows);
        if (nc > 1) CU(hash_chunk_range(fid, d_comm, 0, n_rows, row_stride, n_cols, 0, nc, hash_leaf_bytes(fid, n_rows), nc, cvs, ctx->lc()));
        else CU(hash_columns(fid, d_comm, n_rows, row_stride, n_cols, nullptr, d_hashes, cvs, ctx->lc()));
        CU(merge_tree(cvs, n_cols, nc, d_hashes, np2, tk, ctx->lc()));
        return LCPC_OK;
    }
    // padding leaves n_cols..np2 stay all-zero (lib.rs:685-695)
    if (np2 > n_cols) CU(cudaMemsetAsync(d_hashes + n_cols * 32, 0, (np2 - n_cols) * 32, ctx->stream));
    CU(hash_columns(fid, d_comm, n_rows, row_stride, n_cols, nullptr, d_hashes, cvs, ctx->lc()));
    CU(merkle_tree(d_hashes, np2, ctx->lc()));
    return LCPC_OK;
}

namespace {

// Handles are reference counted (ctx <- plan <- commit), so the order in which a caller
// (e.g. a garbage collector) destroys them does not matter.
void ctx_unref(lcpc_ctx *ctx) {
    if (!ctx || ctx->refs.fetch_sub(1) != 1) return;
    for (lcpc_ctx *s : ctx->subs) ctx_unref(s);
    ctx->subs.clear();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    for (auto e : ctx->events) cudaEventDestroy(e);
    if (ctx->d_tickets) cudaFree(ctx->d_tickets);
    delete ctx->timer;
    delete ctx;
}

void plan_unref(lcpc_plan *plan) {
    if (!plan || plan->refs.fetch_sub(1) != 1) return;
    for (lcpc_plan *s : plan->subs) plan_unref(s);
    plan->subs.clear();
    lcpc_ctx *ctx = plan->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ntt_plan_free(plan->ntt);
    for (auto &m : plan->sdig.pre) free_csr(m);
    for (auto &m : plan->sdig.post) free_csr(m);
    delete plan;
    ctx_unref(ctx);
}


void commit_release(lcpc_commit *c) {
    if (!c) return;
    if (c->plan && !c->shards.empty()) multi::release(c);
    if (c->plan) {
        cudaStream_t s = c->plan->ctx->stream;
        if (c->d_coeffs) cudaFreeAsync(c->d_coeffs, s);
        if (c->d_comm) cudaFreeAsync(c->d_comm, s);
        if (c->d_hashes) cudaFreeAsync(c->d_hashes, s);
        if (c->d_cvs) cudaFreeAsync(c->d_cvs, s);
    }
    lcpc_plan *plan = c->plan;
    delete c;
    plan_unref(plan);
}

// shared tail of the commit entry points: d_coeffs (padded, device) is ready in `c`
int32_t commit_finish(lcpc_plan *plan, lcpc_commit *c, uint64_t *coeffs_out, uint64_t *comm_out, uint8_t *hashes_out) {
    lcpc_ctx *ctx = plan->ctx;
    const int L = limbs_of(plan->fid);
    const size_t wbytes = (size_t)L * sizeof(uint64_t);
    CU(cudaMallocAsync((void **)&c->d_comm, c->n_rows * c->n_cols * wbytes, ctx->stream));
    CU(cudaMallocAsync((void **)&c->d_hashes, (2 * c->np2 - 1) * 32, ctx->stream));
    int32_t rc = encode_dev(plan, c->d_coeffs, c->n_rows, c->d_comm);
    if (rc != LCPC_OK) return rc;
    rc = merkleize_dev(ctx, plan->fid, c->d_comm, c->n_rows, c->n_cols, c->n_cols, c->np2, c->d_hashes, &c->d_cvs);
    if (rc != LCPC_OK) return rc;
    if (coeffs_out)
        CU(cudaMemcpyAsync(coeffs_out, c->d_coeffs, c->n_rows * c->n_per_row * wbytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (comm_out)
        CU(cudaMemcpyAsync(comm_out, c->d_comm, c->n_rows * c->n_cols * wbytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (hashes_out)
        CU(cudaMemcpyAsync(hashes_out, c->d_hashes, (2 * c->np2 - 1) * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LCPC_OK;
}

// Host commit with the three stages overlapped over row chunks: H2D of coefficient rows (stream
// s_in), encode on the compute stream, D2H of encoded rows (stream s_out).  Column hashing needs
// every row, so it runs once after the last chunk is encoded, while earlier chunks are still
// draining over PCIe.  Bit-identical to the unchunked path: rows are independent.
int32_t commit_host_pipelined(lcpc_plan *plan, lcpc_commit *c, const uint64_t *h_coeffs, size_t n_coeffs,
                              uint64_t *coeffs_out, uint64_t *comm_out, uint8_t *hashes_out) {
    lcpc_ctx *ctx = plan->ctx;
    const int L = limbs_of(plan->fid);
    const size_t wbytes = (size_t)L * sizeof(uint64_t);
    const size_t n_rows = c->n_rows, npr = c->n_per_row, n_cols = c->n_cols;
    if (!ctx->s_in) CU(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    if (!ctx->s_out) CU(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    CU(cudaMallocAsync((void **)&c->d_coeffs, n_rows * npr * wbytes, ctx->stream));
    CU(cudaMallocAsync((void **)&c->d_comm, n_rows * n_cols * wbytes, ctx->stream));
    CU(cudaMallocAsync((void **)&c->d_hashes, (2 * c->np2 - 1) * 32, ctx->stream));
    cudaEvent_t ready = ctx->event(0);
    CU(cudaEventRecord(ready, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->s_in, ready, 0));
    CU(cudaStreamWaitEvent(ctx->s_out, ready, 0));
    // Row chunks: 8 equal ones measured best at 2^24 in round 1 (32 chunks of 8 MiB: many small copies in both directions
    // leave bubbles between the copy engines' event waits); the first chunk is split in two so that the device-to-host
    // stream, which bounds the call, starts after half as much input.  LCPC_COMMIT_CHUNKS / LCPC_COMMIT_RAMP override
    // for experiments (tools/e2e_sharded_probe.py).
    size_t want_chunks = 8;
    bool ramp = true;
    if (const char *e = getenv("LCPC_COMMIT_CHUNKS")) want_chunks = (size_t)std::max(1, atoi(e));
    if (const char *e = getenv("LCPC_COMMIT_RAMP")) ramp = atoi(e) != 0;
    std::vector<size_t> bounds{0};
    {
        size_t k = n_rows < want_chunks ? n_rows : want_chunks;
        const size_t per = (n_rows + k - 1) / k;
        if (ramp && per >= 2) bounds.push_back(per / 2);
        for (size_t r = per; r < n_rows; r += per) bounds.push_back(r);
        bounds.push_back(n_rows);
    }
    const size_t n_chunks = bounds.size() - 1;
    for (size_t k = 0; k < n_chunks; k++) {
        const size_t r0 = bounds[k], nr = bounds[k + 1] - bounds[k];
        const size_t e0 = r0 * npr, e1 = (r0 + nr) * npr;  // element range of this chunk's coefficient rows
        const size_t have = n_coeffs > e0 ? (n_coeffs < e1 ? n_coeffs - e0 : e1 - e0) : 0;
        if (have)
            CU(cudaMemcpyAsync(c->d_coeffs + e0 * L, h_coeffs + e0 * L, have * wbytes, cudaMemcpyHostToDevice, ctx->s_in));
        if (have < e1 - e0)  // lib.rs:665-674: zero fill of the ragged tail
            CU(cudaMemsetAsync(c->d_coeffs + (e0 + have) * L, 0, (e1 - e0 - have) * wbytes, ctx->s_in));
        cudaEvent_t in_done = ctx->event(1 + 2 * k), enc_done = ctx->event(2 + 2 * k);
        CU(cudaEventRecord(in_done, ctx->s_in));
        CU(cudaStreamWaitEvent(ctx->stream, in_done, 0));
        int32_t rc = encode_dev(plan, c->d_coeffs + e0 * L, nr, c->d_comm + r0 * n_cols * L);
        if (rc != LCPC_OK) return rc;
        CU(cudaEventRecord(enc_done, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->s_out, enc_done, 0));
        if (comm_out)
            CU(cudaMemcpyAsync(comm_out + r0 * n_cols * L, c->d_comm + r0 * n_cols * L, nr * n_cols * wbytes,
                               cudaMemcpyDeviceToHost, ctx->s_out));
        if (coeffs_out)
            CU(cudaMemcpyAsync(coeffs_out + e0 * L, c->d_coeffs + e0 * L, (e1 - e0) * wbytes, cudaMemcpyDeviceToHost, ctx->s_out));
    }
    int32_t rc = merkleize_dev(ctx, plan->fid, c->d_comm, n_rows, n_cols, n_cols, c->np2, c->d_hashes, &c->d_cvs);
    if (rc != LCPC_OK) return rc;
    if (hashes_out)
        CU(cudaMemcpyAsync(hashes_out, c->d_hashes, (2 * c->np2 - 1) * 32, cudaMemcpyDeviceToHost, ctx->stream));
    // everything is joined on the compute stream before returning
    cudaEvent_t out_done = ctx->event(1 + 2 * n_chunks);
    CU(cudaEventRecord(out_done, ctx->s_out));
    CU(cudaStreamWaitEvent(ctx->stream, out_done, 0));
    CU(cudaStreamSynchronize(ctx->stream));
    return LCPC_OK;
}

int32_t commit_shape(lcpc_plan *plan, size_t n_coeffs, lcpc_commit *c) {
    // lib.rs:656-661
    if (n_coeffs == 0) return fail(LCPC_ERR_DIMS, "cannot commit to zero coefficients");
    c->plan = plan;
    plan->refs.fetch_add(1);
    c->n_per_row = plan->n_per_row;
    c->n_cols = plan->n_cols;
    c->n_rows = (n_coeffs + plan->n_per_row - 1) / plan->n_per_row;
    c->np2 = next_pow2(plan->n_cols);
    if (c->np2 == 0) return fail(LCPC_ERR_TOO_BIG, "n_cols is too large for this encoding");
    return LCPC_OK;
}

// commit through a plan made on a multi-device context (lcpc_multi.cu)
int32_t commit_multi(lcpc_plan *plan, const uint64_t *coeffs, size_t n_coeffs, const uint8_t *file_bytes, size_t n_bytes,
                     uint64_t *coeffs_out, uint64_t *comm_out, uint8_t *hashes_out, lcpc_commit **keep) {
    lcpc_ctx *ctx = plan->ctx;
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    lcpc_commit *c = new (std::nothrow) lcpc_commit;
    if (!c) return fail(LCPC_ERR_NOMEM, "host allocation failed");
    int32_t rc = commit_shape(plan, n_coeffs, c);
    if (rc == LCPC_OK) rc = multi::commit_host(plan, c, coeffs, n_coeffs, file_bytes, n_bytes, coeffs_out, comm_out, hashes_out);
    if (rc != LCPC_OK || !keep) {
        commit_release(c);
        if (rc == LCPC_OK) lcpc_ctx_synchronize(ctx);
    } else {
        *keep = c;
    }
    return rc;
}

}  // namespace

extern "C" {

uint32_t lcpc_abi_version(void) { return 1; }

const char *lcpc_last_error(void) { return last_error().c_str(); }

int32_t lcpc_field_limbs(int32_t field) { return valid_field(field) ? limbs_of(field) : 0; }

int32_t lcpc_field_constants(int32_t field, uint64_t *modulus, uint64_t *one_mont, uint64_t *root_of_unity_mont,
                             int32_t *two_adicity, int32_t *num_bits) {
    if (!valid_field(field)) return fail(LCPC_ERR_INVALID_ARG, "unknown field id");
    const FieldConsts fc = field_consts(field);
    for (int i = 0; i < fc.limbs; i++) {
        if (modulus) modulus[i] = fc.p[i];
        if (one_mont) one_mont[i] = fc.r[i];
        if (root_of_unity_mont) root_of_unity_mont[i] = fc.root[i];
    }
    if (two_adicity) *two_adicity = fc.two_adicity;
    if (num_bits) *num_bits = fc.num_bits;
    return LCPC_OK;
}

static int32_t ctx_create(int32_t device, void *stream, bool own, lcpc_ctx **out) {
    if (!out) return fail(LCPC_ERR_INVALID_ARG, "null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(LCPC_ERR_CUDA, std::string("no CUDA device available (lcpc_b200 has no CPU fallback): ") +
                                       (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= n) return fail(LCPC_ERR_INVALID_ARG, "device index out of range");
    CU(cudaSetDevice(device));
    lcpc_ctx *ctx = new (std::nothrow) lcpc_ctx;
    if (!ctx) return fail(LCPC_ERR_NOMEM, "host allocation failed");
    ctx->device = device;
    if (own) {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete ctx;
            return cuda_fail(e, "cudaStreamCreateWithFlags");
        }
        ctx->own_stream = true;
    } else {
        ctx->stream = (cudaStream_t)stream;
    }
    // keep freed blocks in the pool: repeated commits reuse them without driver calls
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thresh = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
    }
    *out = ctx;
    return LCPC_OK;
}

int32_t lcpc_ctx_create(int32_t device, lcpc_ctx **out) { return ctx_create(device, nullptr, true, out); }

int32_t lcpc_ctx_create_on_stream(int32_t device, void *cuda_stream, lcpc_ctx **out) {
    return ctx_create(device, cuda_stream, false, out);
}

int32_t lcpc_ctx_create_multi(const int32_t *devices, int32_t n_devices, lcpc_ctx **out) {
    if (!out || !devices) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (n_devices < 1 || n_devices > 16 || (n_devices & (n_devices - 1)))
        return fail(LCPC_ERR_INVALID_ARG, "the number of devices must be a power of two, at most 16");
    if (n_devices == 1) return ctx_create(devices[0], nullptr, true, out);
    lcpc_ctx *ctx = nullptr;
    int32_t rc = ctx_create(devices[0], nullptr, true, &ctx);
    if (rc != LCPC_OK) return rc;
    for (int32_t g = 0; g < n_devices; g++) {
        lcpc_ctx *sub = nullptr;
        rc = ctx_create(devices[g], nullptr, true, &sub);
        if (rc != LCPC_OK) {
            ctx_unref(ctx);
            return rc;
        }
        ctx->subs.push_back(sub);
    }
    // every device reads and writes every other device's memory (the chaining-value stores of a sharded commitment, the
    // partial sums of a fold): peer access for plain allocations and for the stream-ordered pools
    for (int32_t a = 0; a < n_devices; a++) {
        for (int32_t b = 0; b < n_devices; b++) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            cudaError_t e = cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (e != cudaSuccess || !can) {
                ctx_unref(ctx);
                return fail(LCPC_ERR_CUDA, "no peer access between the listed devices (NVLink / PCIe P2P required)");
            }
            cudaSetDevice(devices[a]);
            e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
            } else if (e != cudaSuccess) {
                ctx_unref(ctx);
                return cuda_fail(e, "cudaDeviceEnablePeerAccess");
            }
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, devices[b]) == cudaSuccess) {
                cudaMemAccessDesc d{};
                d.location.type = cudaMemLocationTypeDevice;
                d.location.id = devices[a];
                d.flags = cudaMemAccessFlagsProtReadWrite;
                e = cudaMemPoolSetAccess(pool, &d, 1);
                if (e != cudaSuccess) {
                    ctx_unref(ctx);
                    return cuda_fail(e, "cudaMemPoolSetAccess");
                }
            }
        }
    }
    cudaSetDevice(devices[0]);
    *out = ctx;
    return LCPC_OK;
}

int32_t lcpc_ctx_device_count(const lcpc_ctx *ctx) { return !ctx ? 0 : (ctx->subs.empty() ? 1 : (int32_t)ctx->subs.size()); }

int32_t lcpc_ctx_synchronize(lcpc_ctx *ctx) {
    if (!ctx) return fail(LCPC_ERR_INVALID_ARG, "null context");
    for (lcpc_ctx *s : ctx->subs) {
        CU(cudaSetDevice(s->device));
        CU(cudaStreamSynchronize(s->stream));
    }
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return LCPC_OK;
}

int32_t lcpc_ctx_stream(const lcpc_ctx *ctx, void **cuda_stream_out) {
    ctx = primary(const_cast<lcpc_ctx *>(ctx));
    if (!ctx || !cuda_stream_out) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    *cuda_stream_out = (void *)ctx->stream;
    return LCPC_OK;
}

void lcpc_ctx_destroy(lcpc_ctx *ctx) { ctx_unref(ctx); }

uint64_t lcpc_ctx_launch_count(const lcpc_ctx *ctx) {
    if (!ctx) return 0;
    uint64_t n = ctx->launches;
    for (const lcpc_ctx *s : ctx->subs) n += s->launches;
    return n;
}

int32_t lcpc_ctx_kernel_timing(lcpc_ctx *ctx, int32_t enable) {
    ctx = primary(ctx);
    if (!ctx) return fail(LCPC_ERR_INVALID_ARG, "null context");
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (enable && !ctx->timer) ctx->timer = new (std::nothrow) lcpc::KernelTimer;
    if (!enable && ctx->timer) {
        cudaStreamSynchronize(ctx->stream);
        delete ctx->timer;
        ctx->timer = nullptr;
    }
    return LCPC_OK;
}

const char *lcpc_ctx_kernel_timing_report(lcpc_ctx *ctx) {
    ctx = primary(ctx);
    if (!ctx) return "";
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->timing_report.clear();
    if (!ctx->timer) return ctx->timing_report.c_str();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    struct Agg { const char *name; uint64_t n; double ms; };
    std::vector<Agg> agg;
    for (auto &r : ctx->timer->recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        bool found = false;
        for (auto &a : agg)
            if (std::strcmp(a.name, r.name) == 0) { a.n++; a.ms += ms; found = true; break; }
        if (!found) agg.push_back({r.name, 1, (double)ms});
        ctx->timer->pool.push_back(r.a);
        ctx->timer->pool.push_back(r.b);
    }
    ctx->timer->recs.clear();
    char line[160];
    for (auto &a : agg) {
        std::snprintf(line, sizeof line, "%s %llu %.6f\n", a.name, (unsigned long long)a.n, a.ms);
        ctx->timing_report += line;
    }
    return ctx->timing_report.c_str();
}

int32_t lcpc_plan_ligero(lcpc_ctx *ctx, int32_t field, size_t n_per_row, size_t n_cols,
                         const uint64_t *root_of_unity_mont, lcpc_plan **out) {
    if (!ctx || !out) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (!valid_field(field)) return fail(LCPC_ERR_INVALID_ARG, "unknown field id");
    // _dims_ok (lcpc-ligero-pc/src/lib.rs:114-118) + the 2-adicity bound of precomp_fft
    if (!(n_per_row < n_cols) || n_per_row == 0 || (n_cols & (n_cols - 1)) != 0)
        return fail(LCPC_ERR_DIMS, "need 0 < n_per_row < n_cols and n_cols a power of two");
    int log_n = 0;
    while (((size_t)1 << log_n) < n_cols) log_n++;
    if (log_n > field_consts(field).two_adicity) return fail(LCPC_ERR_TOO_BIG, "n_cols exceeds the field's 2-adicity");
    if (!ctx->subs.empty()) {  // multi-device context: the same plan on every device
        lcpc_plan *p = new (std::nothrow) lcpc_plan;
        if (!p) return fail(LCPC_ERR_NOMEM, "host allocation failed");
        p->ctx = ctx;
        ctx->refs.fetch_add(1);
        p->kind = 0;
        p->fid = field;
        p->n_per_row = n_per_row;
        p->n_cols = n_cols;
        for (lcpc_ctx *sub : ctx->subs) {
            lcpc_plan *sp = nullptr;
            int32_t rc = lcpc_plan_ligero(sub, field, n_per_row, n_cols, root_of_unity_mont, &sp);
            if (rc != LCPC_OK) {
                plan_unref(p);
                return rc;
            }
            p->subs.push_back(sp);
        }
        *out = p;
        return LCPC_OK;
    }
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    lcpc_plan *p = new (std::nothrow) lcpc_plan;
    if (!p) return fail(LCPC_ERR_NOMEM, "host allocation failed");
    p->ctx = ctx;
    ctx->refs.fetch_add(1);
    p->kind = 0;
    p->fid = field;
    p->n_per_row = n_per_row;
    p->n_cols = n_cols;
    cudaError_t e = ntt_plan_build(p->ntt, field, log_n, root_of_unity_mont, ctx->lc());
    if (e != cudaSuccess) {
        ntt_plan_free(p->ntt);
        delete p;
        ctx->refs.fetch_sub(1);
        return cuda_fail(e, "ntt_plan_build");
    }
    *out = p;
    return LCPC_OK;
}

int32_t lcpc_plan_brakedown(lcpc_ctx *ctx, int32_t field, size_t n_per_row, size_t n_cols, size_t n_levels,
                            const lcpc_csc *precodes, const lcpc_csc *postcodes, lcpc_plan **out) {
    if (!ctx || !out || !precodes || !postcodes) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (!valid_field(field)) return fail(LCPC_ERR_INVALID_ARG, "unknown field id");
    if (n_levels == 0) return fail(LCPC_ERR_DIMS, "need at least one code level");
    // codeword_length (encode.rs:18-33) and the dims_ok checks of brakedown lib.rs:160-167
    size_t len = precodes[0].cols + postcodes[n_levels - 1].cols;
    for (size_t l = 0; l + 1 < n_levels; l++) len += precodes[l].rows;
    for (size_t l = 0; l < n_levels; l++) len += postcodes[l].rows;
    if (precodes[0].cols != n_per_row || len != n_cols || !(n_per_row < n_cols))
        return fail(LCPC_ERR_DIMS, "n_per_row / n_cols do not match the code matrices");
    // chain consistency: what each product reads must be what earlier levels wrote
    for (size_t l = 0; l + 1 < n_levels; l++)
        if (precodes[l + 1].cols != precodes[l].rows) return fail(LCPC_ERR_DIMS, "precode chain mismatch");
    {
        size_t in_len = postcodes[n_levels - 1].cols;
        for (size_t l = n_levels; l-- > 0;) {
            if (postcodes[l].cols != in_len) return fail(LCPC_ERR_DIMS, "postcode chain mismatch");
            if (l > 0) in_len = precodes[l - 1].rows + in_len + postcodes[l].rows;
        }
    }
    if (!ctx->subs.empty()) {  // multi-device context: the matrices are uploaded to every device
        lcpc_plan *p = new (std::nothrow) lcpc_plan;
        if (!p) return fail(LCPC_ERR_NOMEM, "host allocation failed");
        p->ctx = ctx;
        ctx->refs.fetch_add(1);
        p->kind = 1;
        p->fid = field;
        p->n_per_row = n_per_row;
        p->n_cols = n_cols;
        for (lcpc_ctx *sub : ctx->subs) {
            lcpc_plan *sp = nullptr;
            int32_t rc = lcpc_plan_brakedown(sub, field, n_per_row, n_cols, n_levels, precodes, postcodes, &sp);
            if (rc != LCPC_OK) {
                plan_unref(p);
                return rc;
            }
            p->subs.push_back(sp);
        }
        *out = p;
        return LCPC_OK;
    }
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    lcpc_plan *p = new (std::nothrow) lcpc_plan;
    if (!p) return fail(LCPC_ERR_NOMEM, "host allocation failed");
    p->ctx = ctx;
    ctx->refs.fetch_add(1);
    p->kind = 1;
    p->fid = field;
    p->n_per_row = n_per_row;
    p->n_cols = n_cols;
    p->sdig.fid = field;
    p->sdig.n_per_row = n_per_row;
    p->sdig.n_cols = n_cols;
    p->sdig.pre.resize(n_levels);
    p->sdig.post.resize(n_levels);
    const int L = limbs_of(field);
    for (size_t l = 0; l < n_levels; l++) {
        int32_t rc = upload_csr(precodes[l], field, L, p->sdig.pre[l]);
        if (rc == LCPC_OK) rc = upload_csr(postcodes[l], field, L, p->sdig.post[l]);
        if (rc != LCPC_OK) {
            lcpc_plan_destroy(p);
            return rc;
        }
    }
    *out = p;
    return LCPC_OK;
}

int32_t lcpc_plan_get_dims(const lcpc_plan *plan, size_t len, size_t *n_rows, size_t *n_per_row, size_t *n_cols) {
    if (!plan) return fail(LCPC_ERR_INVALID_ARG, "null plan");
    if (n_rows) *n_rows = (len + plan->n_per_row - 1) / plan->n_per_row;
    if (n_per_row) *n_per_row = plan->n_per_row;
    if (n_cols) *n_cols = plan->n_cols;
    return LCPC_OK;
}

void lcpc_plan_destroy(lcpc_plan *plan) { plan_unref(plan); }


int32_t lcpc_encode_rows(lcpc_plan *plan, uint64_t *rows, size_t n_rows) {
    plan = primary(plan);  // a plan made on a multi-device context: this call runs on its first device
    if (!plan || (!rows && n_rows)) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (n_rows == 0) return LCPC_OK;
    lcpc_ctx *ctx = plan->ctx;
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const int L = limbs_of(plan->fid);
    const size_t bytes = n_rows * plan->n_cols * L * sizeof(uint64_t);
    DevBuf buf;
    CU(buf.alloc(bytes, ctx->stream));
    CU(cudaMemcpyAsync(buf.p, rows, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (plan->kind == 0) {
        // fft_io_pc transforms whatever the row holds (all n_cols entries)
        CU(ntt_encode(plan->ntt, buf.as<uint64_t>(), plan->n_cols, plan->n_cols, buf.as<uint64_t>(), n_rows, ctx->lc()));
    } else {
        DevBuf tmp;
        CU(tmp.alloc(sdig_tmp_elems(plan->sdig, n_rows) * L * sizeof(uint64_t), ctx->stream));
        CU(sdig_encode(plan->sdig, buf.as<uint64_t>(), plan->n_cols, buf.as<uint64_t>(), n_rows, tmp.as<uint64_t>(), ctx->lc()));
    }
    CU(cudaMemcpyAsync(rows, buf.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LCPC_OK;
}

int32_t lcpc_dev_decode(lcpc_plan *plan, uint64_t *d_rows, size_t n_rows) {
    plan = primary(plan);  // a plan made on a multi-device context: this call runs on its first device
    if (!plan || (!d_rows && n_rows)) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (plan->kind != 0) return fail(LCPC_ERR_ENCODE, "only Ligero (Reed-Solomon) rows have an inverse transform");
    if (n_rows == 0) return LCPC_OK;
    lcpc_ctx *ctx = plan->ctx;
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(ntt_decode(plan->ntt, d_rows, n_rows, ctx->lc()));
    return LCPC_OK;
}

int32_t lcpc_decode_rows(lcpc_plan *plan, uint64_t *rows, size_t n_rows) {
    plan = primary(plan);  // a plan made on a multi-device context: this call runs on its first device
    if (!plan || (!rows && n_rows)) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (plan->kind != 0) return fail(LCPC_ERR_ENCODE, "only Ligero (Reed-Solomon) rows have an inverse transform");
    if (n_rows == 0) return LCPC_OK;
    lcpc_ctx *ctx = plan->ctx;
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = n_rows * plan->n_cols * limbs_of(plan->fid) * sizeof(uint64_t);
    DevBuf buf;
    CU(buf.alloc(bytes, ctx->stream));
    CU(cudaMemcpyAsync(buf.p, rows, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(ntt_decode(plan->ntt, buf.as<uint64_t>(), n_rows, ctx->lc()));
    CU(cudaMemcpyAsync(rows, buf.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LCPC_OK;
}

int32_t lcpc_commit_host(lcpc_plan *plan, const uint64_t *coeffs, size_t n_coeffs, uint64_t *coeffs_out,
                         uint64_t *comm_out, uint8_t *hashes_out, lcpc_commit **keep) {
    if (keep) *keep = nullptr;
    if (!plan || !coeffs) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (!plan->subs.empty()) {
        // multi-device context: shard over its devices when every device gets whole BLAKE3 chunks of rows, else (small
        // commitments, 24-byte elements) the first device takes it
        const size_t rows = n_coeffs ? (n_coeffs + plan->n_per_row - 1) / plan->n_per_row : 0;
        if (multi::usable(plan, rows)) return commit_multi(plan, coeffs, n_coeffs, nullptr, 0, coeffs_out, comm_out, hashes_out, keep);
        plan = primary(plan);
    }
    lcpc_ctx *ctx = plan->ctx;
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    lcpc_commit *c = new (std::nothrow) lcpc_commit;
    if (!c) return fail(LCPC_ERR_NOMEM, "host allocation failed");
    int32_t rc = commit_shape(plan, n_coeffs, c);
    const size_t wbytes = (size_t)limbs_of(plan->fid) * sizeof(uint64_t);
    auto body = [&]() -> int32_t {
        const size_t padded = c->n_rows * c->n_per_row;
        // large commits: overlap PCIe in, kernels and PCIe out over row chunks
        if (c->n_rows >= 8 && c->n_rows * c->n_cols * wbytes >= ((size_t)8 << 20))
            return commit_host_pipelined(plan, c, coeffs, n_coeffs, coeffs_out, comm_out, hashes_out);
        CU(cudaMallocAsync((void **)&c->d_coeffs, padded * wbytes, ctx->stream));
        CU(cudaMemcpyAsync(c->d_coeffs, coeffs, n_coeffs * wbytes, cudaMemcpyHostToDevice, ctx->stream));
        if (padded > n_coeffs)  // lib.rs:665-674: the last row is zero-filled
            CU(cudaMemsetAsync(c->d_coeffs + n_coeffs * limbs_of(plan->fid), 0, (padded - n_coeffs) * wbytes, ctx->stream));
        return commit_finish(plan, c, coeffs_out, comm_out, hashes_out);
    };
    if (rc == LCPC_OK) rc = body();
    if (rc != LCPC_OK || !keep) {
        commit_release(c);
        if (rc == LCPC_OK) cudaStreamSynchronize(ctx->stream);
    } else {
        *keep = c;
    }
    return rc;
}

int32_t lcpc_commit_bytes_host(lcpc_plan *plan, const uint8_t *file_bytes, size_t n_bytes, uint64_t *coeffs_out,
                               uint64_t *comm_out, uint8_t *hashes_out, lcpc_commit **keep) {
    if (keep) *keep = nullptr;
    if (!plan || !file_bytes) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (plan->fid != FT63 && plan->fid != FT253_192)
        return fail(LCPC_ERR_INVALID_ARG, "byte packing is defined for the DataField types only (Ft63, Ft253_192)");
    if (!plan->subs.empty()) {
        const size_t n_el = (n_bytes + 6) / 7;
        const size_t rows = n_el ? (n_el + plan->n_per_row - 1) / plan->n_per_row : 0;
        if (plan->fid == FT63 && multi::usable(plan, rows))
            return commit_multi(plan, nullptr, n_el, file_bytes, n_bytes, coeffs_out, comm_out, hashes_out, keep);
        plan = primary(plan);
    }
    lcpc_ctx *ctx = plan->ctx;
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const bool wide = plan->fid == FT253_192;  // 31 data bytes per element instead of 7
    const size_t L = (size_t)limbs_of(plan->fid);
    const size_t n_coeffs = wide ? (n_bytes + 30) / 31 : (n_bytes + 6) / 7;
    lcpc_commit *c = new (std::nothrow) lcpc_commit;
    if (!c) return fail(LCPC_ERR_NOMEM, "host allocation failed");
    int32_t rc = commit_shape(plan, n_coeffs, c);
    auto body = [&]() -> int32_t {
        const size_t padded = c->n_rows * c->n_per_row;
        DevBuf raw;
        CU(raw.alloc(n_bytes, ctx->stream));
        CU(cudaMemcpyAsync(raw.p, file_bytes, n_bytes, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMallocAsync((void **)&c->d_coeffs, padded * L * sizeof(uint64_t), ctx->stream));
        if (wide) {
            DevBuf bad;
            CU(bad.alloc(sizeof(uint32_t), ctx->stream));
            CU(cudaMemsetAsync(bad.p, 0, sizeof(uint32_t), ctx->stream));
            CU(pack_bytes31(raw.as<uint8_t>(), n_bytes, c->d_coeffs, bad.as<uint32_t>(), ctx->lc()));
            uint32_t h_bad = 0;
            CU(cudaMemcpyAsync(&h_bad, bad.p, sizeof h_bad, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            if (h_bad)
                return fail(LCPC_ERR_INVALID_ARG,
                            "Ft253_192::from_data_bytes: a 31-byte group is not below the modulus (byte 24 of the group > 0x1f)");
        } else {
            CU(pack_bytes7(raw.as<uint8_t>(), n_bytes, c->d_coeffs, ctx->lc()));
        }
        if (padded > n_coeffs)
            CU(cudaMemsetAsync(c->d_coeffs + n_coeffs * L, 0, (padded - n_coeffs) * L * sizeof(uint64_t), ctx->stream));
        return commit_finish(plan, c, coeffs_out, comm_out, hashes_out);
    };
    if (rc == LCPC_OK) rc = body();
    if (rc != LCPC_OK || !keep) {
        commit_release(c);
        if (rc == LCPC_OK) cudaStreamSynchronize(ctx->stream);
    } else {
        *keep = c;
    }
    return rc;
}

int32_t lcpc_commit_dev(lcpc_plan *plan, const uint64_t *d_coeffs, size_t n_coeffs, lcpc_commit **keep) {
    plan = primary(plan);  // a plan made on a multi-device context: this call runs on its first device
    if (!plan || !d_coeffs || !keep) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    *keep = nullptr;
    lcpc_ctx *ctx = plan->ctx;
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    lcpc_commit *c = new (std::nothrow) lcpc_commit;
    if (!c) return fail(LCPC_ERR_NOMEM, "host allocation failed");
    int32_t rc = commit_shape(plan, n_coeffs, c);
    const int L = limbs_of(plan->fid);
    const size_t wbytes = (size_t)L * sizeof(uint64_t);
    auto body = [&]() -> int32_t {
        const size_t padded = c->n_rows * c->n_per_row;
        CU(cudaMallocAsync((void **)&c->d_coeffs, padded * wbytes, ctx->stream));
        CU(cudaMemcpyAsync(c->d_coeffs, d_coeffs, n_coeffs * wbytes, cudaMemcpyDeviceToDevice, ctx->stream));
        if (padded > n_coeffs)
            CU(cudaMemsetAsync(c->d_coeffs + n_coeffs * L, 0, (padded - n_coeffs) * wbytes, ctx->stream));
        return commit_finish(plan, c, nullptr, nullptr, nullptr);
    };
    if (rc == LCPC_OK) rc = body();
    if (rc != LCPC_OK) {
        commit_release(c);
        return rc;
    }
    *keep = c;
    return LCPC_OK;
}

int32_t lcpc_commit_get_dims(const lcpc_commit *c, size_t *n_rows, size_t *n_per_row, size_t *n_cols) {
    if (!c) return fail(LCPC_ERR_INVALID_ARG, "null commit");
    if (n_rows) *n_rows = c->n_rows;
    if (n_per_row) *n_per_row = c->n_per_row;
    if (n_cols) *n_cols = c->n_cols;
    return LCPC_OK;
}

int32_t lcpc_commit_root(lcpc_commit *c, uint8_t root_out[LCPC_DIGEST_BYTES]) {
    if (!c || !root_out) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (!c->shards.empty()) {
        std::lock_guard<std::mutex> gm(c->mu);
        std::lock_guard<std::mutex> gm2(c->plan->ctx->mu);
        return multi::root(c, root_out);
    }
    lcpc_ctx *ctx = c->plan->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(root_out, c->d_hashes + (2 * c->np2 - 2) * 32, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LCPC_OK;
}

int32_t lcpc_commit_download(lcpc_commit *c, uint64_t *coeffs_out, uint64_t *comm_out, uint8_t *hashes_out) {
    if (!c) return fail(LCPC_ERR_INVALID_ARG, "null commit");
    if (!c->shards.empty()) {
        std::lock_guard<std::mutex> gm(c->mu);
        std::lock_guard<std::mutex> gm2(c->plan->ctx->mu);
        return multi::download(c, coeffs_out, comm_out, hashes_out);
    }
    lcpc_ctx *ctx = c->plan->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const size_t wbytes = (size_t)limbs_of(c->plan->fid) * sizeof(uint64_t);
    if (coeffs_out)
        CU(cudaMemcpyAsync(coeffs_out, c->d_coeffs, c->n_rows * c->n_per_row * wbytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (comm_out)
        CU(cudaMemcpyAsync(comm_out, c->d_comm, c->n_rows * c->n_cols * wbytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (hashes_out)
        CU(cudaMemcpyAsync(hashes_out, c->d_hashes, (2 * c->np2 - 1) * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LCPC_OK;
}

int32_t lcpc_commit_device_ptrs(lcpc_commit *c, uint64_t **d_coeffs, uint64_t **d_comm, uint8_t **d_hashes) {
    if (!c) return fail(LCPC_ERR_INVALID_ARG, "null commit");
    if (!c->shards.empty()) return fail(LCPC_ERR_INVALID_ARG, "a multi-device commitment has no single set of device buffers");
    if (d_coeffs) *d_coeffs = c->d_coeffs;
    if (d_comm) *d_comm = c->d_comm;
    if (d_hashes) *d_hashes = c->d_hashes;
    return LCPC_OK;
}

void lcpc_commit_free(lcpc_commit *c) {
    if (!c) return;
    if (c->plan && !c->shards.empty()) lcpc_ctx_synchronize(c->plan->ctx);
    if (c->plan) cudaSetDevice(c->plan->ctx->device);
    commit_release(c);  // may drop the last reference to the plan / context: hold no lock here
}

int32_t lcpc_fold_host(lcpc_commit *c, int32_t which, const uint64_t *tensors, size_t n_tensors, uint64_t *out) {
    if (!c || !tensors || !out) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (which != 0 && which != 1) return fail(LCPC_ERR_INVALID_ARG, "which must be 0 (coeffs) or 1 (comm)");
    if (n_tensors == 0) return LCPC_OK;
    if (!c->shards.empty()) {
        std::lock_guard<std::mutex> gm(c->mu);
        std::lock_guard<std::mutex> gm2(c->plan->ctx->mu);
        return multi::fold_host(c, which, tensors, n_tensors, out);
    }
    lcpc_ctx *ctx = c->plan->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const int fid = c->plan->fid;
    const size_t wbytes = (size_t)limbs_of(fid) * sizeof(uint64_t);
    const size_t width = which == 0 ? c->n_per_row : c->n_cols;
    const uint64_t *mat = which == 0 ? c->d_coeffs : c->d_comm;
    DevBuf d_t, d_o, d_s;
    CU(d_t.alloc(n_tensors * c->n_rows * wbytes, ctx->stream));
    CU(d_o.alloc(n_tensors * width * wbytes, ctx->stream));
    CU(d_s.alloc(fold_scratch_bytes(fid, c->n_rows, width, n_tensors), ctx->stream));
    CU(cudaMemcpyAsync(d_t.p, tensors, n_tensors * c->n_rows * wbytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(fold(fid, mat, c->n_rows, width, width, d_t.as<uint64_t>(), n_tensors, d_o.as<uint64_t>(), d_s.as<uint64_t>(),
            ctx->lc()));
    CU(cudaMemcpyAsync(out, d_o.p, n_tensors * width * wbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LCPC_OK;
}

int32_t lcpc_open_columns_host(lcpc_commit *c, const uint64_t *cols, size_t n, uint64_t *cols_out, uint8_t *paths_out) {
    if (!c || (!cols && n)) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    for (size_t i = 0; i < n; i++)
        if (cols[i] >= c->n_cols) return fail(LCPC_ERR_COLUMN_NUMBER, "bad column number");
    if (n == 0) return LCPC_OK;
    if (!c->shards.empty()) {
        std::lock_guard<std::mutex> gm(c->mu);
        std::lock_guard<std::mutex> gm2(c->plan->ctx->mu);
        return multi::open_columns_host(c, cols, n, cols_out, paths_out);
    }
    lcpc_ctx *ctx = c->plan->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const int fid = c->plan->fid;
    const size_t wbytes = (size_t)limbs_of(fid) * sizeof(uint64_t);
    int depth = 0;
    while (((size_t)1 << depth) < c->np2) depth++;
    DevBuf d_cols, d_out, d_paths;
    CU(d_cols.alloc(n * sizeof(uint64_t), ctx->stream));
    CU(cudaMemcpyAsync(d_cols.p, cols, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    if (cols_out) {
        CU(d_out.alloc(n * c->n_rows * wbytes, ctx->stream));
        CU(gather_columns(fid, c->d_comm, c->n_rows, c->n_cols, d_cols.as<uint64_t>(), n, d_out.as<uint64_t>(), ctx->lc()));
        CU(cudaMemcpyAsync(cols_out, d_out.p, n * c->n_rows * wbytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (paths_out && depth > 0) {
        CU(d_paths.alloc(n * (size_t)depth * 32, ctx->stream));
        CU(gather_paths(c->d_hashes, c->np2, d_cols.as<uint64_t>(), n, d_paths.as<uint8_t>(), ctx->lc()));
        CU(cudaMemcpyAsync(paths_out, d_paths.p, n * (size_t)depth * 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return LCPC_OK;
}

int32_t lcpc_leaves_host(lcpc_commit *c, const uint64_t *cols, size_t n, uint8_t *leaves_out) {
    if (!c || (!cols && n) || (!leaves_out && n)) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    for (size_t i = 0; i < n; i++)
        if (cols[i] >= c->n_cols) return fail(LCPC_ERR_COLUMN_NUMBER, "bad column number");
    if (n == 0) return LCPC_OK;
    if (!c->shards.empty()) {
        std::lock_guard<std::mutex> gm(c->mu);
        std::lock_guard<std::mutex> gm2(c->plan->ctx->mu);
        return multi::leaves_host(c, cols, n, leaves_out);
    }
    lcpc_ctx *ctx = c->plan->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const int fid = c->plan->fid;
    DevBuf d_cols, d_leaves, d_s;
    CU(d_cols.alloc(n * sizeof(uint64_t), ctx->stream));
    CU(d_leaves.alloc(n * 32, ctx->stream));
    CU(d_s.alloc(hash_scratch_bytes(fid, c->n_rows, n), ctx->stream));
    CU(cudaMemcpyAsync(d_cols.p, cols, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(hash_columns(fid, c->d_comm, c->n_rows, c->n_cols, n, d_cols.as<uint64_t>(), d_leaves.as<uint8_t>(), d_s.as<uint8_t>(),
                    ctx->lc()));
    CU(cudaMemcpyAsync(leaves_out, d_leaves.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LCPC_OK;
}

// ---- device-pointer building blocks -------------------------------------------------

int32_t lcpc_dev_encode(lcpc_plan *plan, const uint64_t *d_coeffs, size_t n_rows, uint64_t *d_comm) {
    plan = primary(plan);  // a plan made on a multi-device context: this call runs on its first device
    if (!plan || !d_coeffs || !d_comm) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(plan->ctx->mu);
    CU(cudaSetDevice(plan->ctx->device));
    return encode_dev(plan, d_coeffs, n_rows, d_comm);
}

int32_t lcpc_dev_encode_scatter(lcpc_plan *plan, const uint64_t *d_coeffs, size_t n_rows, uint64_t row0,
                                uint64_t *d_scratch, uint64_t *const *peer_blocks, size_t n_peers) {
    plan = primary(plan);  // a plan made on a multi-device context: this call runs on its first device
    if (!plan || !d_coeffs || !peer_blocks) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    // the PADDED column range (next power of two: Brakedown's n_cols is not one) is what the ranks split
    size_t np2 = 1;
    while (np2 < plan->n_cols) np2 <<= 1;
    if (n_peers == 0 || n_peers > 16 || (n_peers & (n_peers - 1)) || np2 % n_peers)
        return fail(LCPC_ERR_DIMS, "n_peers must be a power of two <= 16 dividing the padded column count");
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(plan->ctx->mu);
    lcpc_ctx *ctx = plan->ctx;
    CU(cudaSetDevice(ctx->device));
    ScatterDst sc{};
    size_t cb = np2 / n_peers;
    sc.log_cb = 0;
    while (((size_t)1 << sc.log_cb) < cb) sc.log_cb++;
    sc.row0 = row0;
    for (size_t i = 0; i < n_peers; i++) {
        if (!peer_blocks[i]) return fail(LCPC_ERR_INVALID_ARG, "null peer pointer");
        sc.base[i] = peer_blocks[i];
    }
    if (plan->kind != 0) {
        // Brakedown: the transposing passes that would write comm store into the owners' matrices instead
        if (n_rows == 0) return LCPC_OK;
        DevBuf tmp;
        CU(tmp.alloc(sdig_tmp_elems(plan->sdig, n_rows) * limbs_of(plan->fid) * sizeof(uint64_t), ctx->stream));
        CU(sdig_encode(plan->sdig, d_coeffs, plan->n_per_row, nullptr, n_rows, tmp.as<uint64_t>(), ctx->lc(), &sc));
        return LCPC_OK;
    }
    if (plan->ntt.passes.size() > 1 && !d_scratch) return fail(LCPC_ERR_INVALID_ARG, "scratch needed for multi-pass transforms");
    cudaError_t e = ntt_encode(plan->ntt, d_coeffs, plan->n_per_row, plan->n_per_row, d_scratch, n_rows, ctx->lc(), &sc);
    if (e == cudaErrorInvalidValue) return fail(LCPC_ERR_DIMS, "column blocks narrower than the transform's shared-memory block");
    CU(e);
    return LCPC_OK;
}

int32_t lcpc_dev_hash_columns(lcpc_ctx *ctx, int32_t field, const uint64_t *d_mat, size_t n_rows, size_t row_stride,
                              size_t n_cols, uint8_t *d_leaves) {
    ctx = primary(ctx);
    if (!ctx || !d_mat || !d_leaves) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (!valid_field(field)) return fail(LCPC_ERR_INVALID_ARG, "unknown field id");
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    DevBuf scratch;
    CU(scratch.alloc(hash_scratch_bytes(field, n_rows, n_cols), ctx->stream));
    CU(hash_columns(field, d_mat, n_rows, row_stride, n_cols, nullptr, d_leaves, scratch.as<uint8_t>(), ctx->lc()));
    return LCPC_OK;
}

int32_t lcpc_dev_hash_chunk_range(lcpc_ctx *ctx, int32_t field, const uint64_t *d_mat, uint64_t row_base,
                                  size_t n_rows_total, size_t row_stride, size_t n_cols, uint64_t chunk0,
                                  uint64_t chunk_end, uint8_t *d_cvs) {
    ctx = primary(ctx);
    if (!ctx || !d_mat || !d_cvs) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (!valid_field(field)) return fail(LCPC_ERR_INVALID_ARG, "unknown field id");
    const uint64_t w = 8ull * (uint64_t)limbs_of(field);
    const uint64_t total = 32 + (uint64_t)n_rows_total * w, n_chunks = (total + 1023) / 1024;
    if (1024 % w != 0) return fail(LCPC_ERR_DIMS, "elements straddle BLAKE3 chunk boundaries for this field");
    if (n_chunks < 2) return fail(LCPC_ERR_DIMS, "single-chunk leaf: use lcpc_dev_hash_columns");
    if (chunk0 > chunk_end || chunk_end > n_chunks) return fail(LCPC_ERR_INVALID_ARG, "chunk range outside the leaf");
    // first row of chunk0 (chunk 0 starts with the 32-byte zero prefix) must not lie in front of the window
    const uint64_t first_row = chunk0 == 0 ? 0 : (chunk0 * 1024 - 32) / w;
    if (chunk0 < chunk_end && first_row < row_base) return fail(LCPC_ERR_INVALID_ARG, "chunk range starts before row_base");
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    // the kernel indexes chaining values by absolute chunk number; the caller's buffer starts at chunk0
    uint8_t *cvs_abs = reinterpret_cast<uint8_t *>(reinterpret_cast<uintptr_t>(d_cvs) - (uintptr_t)(chunk0 * n_cols * 32));
    CU(hash_chunk_range(field, d_mat, (int64_t)row_base, n_rows_total, row_stride, n_cols, chunk0, chunk_end, total, n_chunks,
                        cvs_abs, ctx->lc()));
    return LCPC_OK;
}

int32_t lcpc_dev_hash_chunk_range_scatter(lcpc_ctx *ctx, int32_t field, const uint64_t *d_mat, uint64_t row_base,
                                          size_t n_rows_total, size_t row_stride, size_t n_cols, uint64_t chunk0,
                                          uint64_t chunk_end, uint8_t *const *peer_cvs, size_t n_peers) {
    ctx = primary(ctx);
    if (!ctx || !d_mat || !peer_cvs) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (!valid_field(field)) return fail(LCPC_ERR_INVALID_ARG, "unknown field id");
    if (n_peers == 0 || n_peers > 16 || (n_peers & (n_peers - 1)) || n_cols == 0 || next_pow2(n_cols) < n_peers)
        return fail(LCPC_ERR_DIMS, "n_peers must be a power of two <= 16 and <= the padded column count");
    const uint64_t w = 8ull * (uint64_t)limbs_of(field);
    const uint64_t total = 32 + (uint64_t)n_rows_total * w, n_chunks = (total + 1023) / 1024;
    if (1024 % w != 0) return fail(LCPC_ERR_DIMS, "elements straddle BLAKE3 chunk boundaries for this field");
    if (n_chunks < 2) return fail(LCPC_ERR_DIMS, "single-chunk leaf: use lcpc_dev_hash_columns");
    if (chunk0 > chunk_end || chunk_end > n_chunks) return fail(LCPC_ERR_INVALID_ARG, "chunk range outside the leaf");
    const uint64_t first_row = chunk0 == 0 ? 0 : (chunk0 * 1024 - 32) / w;
    if (chunk0 < chunk_end && first_row < row_base) return fail(LCPC_ERR_INVALID_ARG, "chunk range starts before row_base");
    CvScatter sc{};
    sc.log_cb = 0;
    while (((size_t)1 << sc.log_cb) < next_pow2(n_cols) / n_peers) sc.log_cb++;  // blocks of the PADDED leaf range
    for (size_t i = 0; i < n_peers; i++) {
        if (!peer_cvs[i]) return fail(LCPC_ERR_INVALID_ARG, "null peer pointer");
        sc.base[i] = reinterpret_cast<uint32_t *>(peer_cvs[i]);
    }
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(hash_chunk_range_scatter(field, d_mat, (int64_t)row_base, n_rows_total, row_stride, n_cols, chunk0, chunk_end, total,
                                n_chunks, sc, ctx->lc()));
    return LCPC_OK;
}

int32_t lcpc_dev_hash_merge(lcpc_ctx *ctx, const uint8_t *d_cvs, size_t n_cols, uint64_t n_chunks, uint8_t *d_leaves) {
    ctx = primary(ctx);
    if (!ctx || !d_cvs || !d_leaves) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (n_chunks < 2) return fail(LCPC_ERR_DIMS, "a single chaining value is the leaf itself");
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(hash_merge(d_cvs, n_cols, n_chunks, d_leaves, ctx->lc()));
    return LCPC_OK;
}

int32_t lcpc_dev_hash_merge_tree(lcpc_ctx *ctx, const uint8_t *d_cvs, size_t n_cols, uint64_t n_chunks, uint8_t *d_hashes,
                                 size_t n_leaves, size_t cv_stride) {
    ctx = primary(ctx);
    if (!ctx || !d_hashes || (!d_cvs && n_chunks > 1)) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (n_leaves == 0 || (n_leaves & (n_leaves - 1)) || n_cols > n_leaves) return fail(LCPC_ERR_DIMS, "n_leaves must be a power of two >= n_cols");
    if (n_chunks == 0) return fail(LCPC_ERR_DIMS, "a leaf has at least one chunk");
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    unsigned *tk = nullptr;
    CU(ctx->tickets(1, &tk));
    if (cv_stride != 0 && cv_stride < n_cols) return fail(LCPC_ERR_DIMS, "cv_stride below n_cols");
    CU(merge_tree(d_cvs, n_cols, n_chunks, d_hashes, n_leaves, tk, ctx->lc(), cv_stride));
    return LCPC_OK;
}

int32_t lcpc_dev_merkleize(lcpc_ctx *ctx, int32_t field, const uint64_t *d_mat, size_t n_rows, size_t row_stride, size_t n_cols,
                           uint8_t *d_hashes) {
    ctx = primary(ctx);
    if (!ctx || !d_mat || !d_hashes) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (!valid_field(field)) return fail(LCPC_ERR_INVALID_ARG, "unknown field id");
    if (n_cols == 0 || n_cols > row_stride) return fail(LCPC_ERR_DIMS, "bad column window");
    const size_t np2 = next_pow2(n_cols);
    if (np2 == 0) return fail(LCPC_ERR_TOO_BIG, "n_cols is too large");
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    return merkleize_dev(ctx, field, d_mat, n_rows, row_stride, n_cols, np2, d_hashes, nullptr);
}

int32_t lcpc_dev_merkle_tree(lcpc_ctx *ctx, uint8_t *d_hashes, size_t n_leaves) {
    ctx = primary(ctx);
    if (!ctx || !d_hashes) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (n_leaves == 0 || (n_leaves & (n_leaves - 1))) return fail(LCPC_ERR_DIMS, "n_leaves must be a power of two");
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(merkle_tree(d_hashes, n_leaves, ctx->lc()));
    return LCPC_OK;
}

int32_t lcpc_dev_fold(lcpc_ctx *ctx, int32_t field, const uint64_t *d_mat, size_t n_rows, size_t width,
                      size_t row_stride, const uint64_t *d_tensors, size_t n_tensors, uint64_t *d_out) {
    ctx = primary(ctx);
    if (!ctx || !d_mat || !d_tensors || !d_out) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (!valid_field(field)) return fail(LCPC_ERR_INVALID_ARG, "unknown field id");
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    DevBuf d_s;
    CU(d_s.alloc(fold_scratch_bytes(field, n_rows, width, n_tensors), ctx->stream));
    CU(fold(field, d_mat, n_rows, width, row_stride, d_tensors, n_tensors, d_out, d_s.as<uint64_t>(), ctx->lc()));
    return LCPC_OK;
}

int32_t lcpc_dev_add_partials(lcpc_ctx *ctx, int32_t field, const uint64_t *d_parts, size_t n_parts, size_t n,
                              uint64_t *d_out) {
    ctx = primary(ctx);
    if (!ctx || !d_parts || !d_out) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (!valid_field(field)) return fail(LCPC_ERR_INVALID_ARG, "unknown field id");
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(add_partials(field, d_parts, n_parts, n, d_out, ctx->lc()));
    return LCPC_OK;
}

int32_t lcpc_dev_gather_columns(lcpc_ctx *ctx, int32_t field, const uint64_t *d_mat, size_t n_rows, size_t row_stride,
                                const uint64_t *d_cols, size_t n, uint64_t *d_out) {
    ctx = primary(ctx);
    if (!ctx || !d_mat || !d_cols || !d_out) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (!valid_field(field)) return fail(LCPC_ERR_INVALID_ARG, "unknown field id");
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(gather_columns(field, d_mat, n_rows, row_stride, d_cols, n, d_out, ctx->lc()));
    return LCPC_OK;
}

int32_t lcpc_dev_gather_paths(lcpc_ctx *ctx, const uint8_t *d_hashes, size_t n_leaves, const uint64_t *d_cols, size_t n,
                              uint8_t *d_paths) {
    ctx = primary(ctx);
    if (!ctx || !d_hashes || (n && (!d_cols || !d_paths))) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (n_leaves == 0 || (n_leaves & (n_leaves - 1))) return fail(LCPC_ERR_DIMS, "n_leaves must be a power of two");
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(gather_paths(d_hashes, n_leaves, d_cols, n, d_paths, ctx->lc()));
    return LCPC_OK;
}

int32_t lcpc_dev_pack_bytes7(lcpc_ctx *ctx, const uint8_t *d_bytes, size_t n_bytes, uint64_t *d_elems) {
    ctx = primary(ctx);
    if (!ctx || !d_bytes || !d_elems) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(pack_bytes7(d_bytes, n_bytes, d_elems, ctx->lc()));
    return LCPC_OK;
}

}  // extern "C"
