// Streaming commit and in-place row edits: the proof-of-storage file path beyond one resident matrix.
//
// lcpc_stream_* stands in for proof-of-storage's EncodedFileWriter + ColumnDigestAccumulator
// (src/lcpc_online/encoded_file_writer.rs:33-508, column_digest_accumulator.rs:17-118): rows arrive in order,
// each is encoded, folded into every column's running digest and appended to a column-major encoded file.
// The reference keeps one BLAKE3 hasher per column and feeds it 8 bytes per row; here a block of rows is
// encoded at once and the column digests advance one BLAKE3 *chunk* (1 KiB = 128 rows of Ft63) at a time:
// chunk chaining values are independent, so a chunk is hashed as soon as one byte beyond it exists and its
// rows can be dropped.  Device memory is O(block_rows * n_cols), not O(file).
//
// lcpc_commit_update_rows_host stands in for FileHandler::edit_bytes -> reencode_row ->
// recalculate_merkle_tree (src/lcpc_online/file_handler.rs:279-402, 474-481): the edited rows are re-encoded
// and only the chunks of the column leaves that contain them are re-hashed (the reference re-reads and
// re-hashes every column from disk).
#include <algorithm>
#include <thread>

#include "lcpc_handles.h"

using namespace lcpc;
using namespace lcpc::abi;

struct lcpc_stream {
    lcpc_plan *plan = nullptr;
    int fid = 0, L = 1;
    size_t n_per_row = 0, n_cols = 0, np2 = 0, wbytes = 8;
    size_t max_rows = 0, block_rows = 0, pend_cap = 0;
    // staging is double buffered: the H2D copy of block k+1 (copy stream) overlaps the encode + hash of block k
    uint64_t *d_in[2] = {nullptr, nullptr};    // block_rows x n_per_row coefficients
    uint8_t *d_bytes[2] = {nullptr, nullptr};  // block_rows x n_per_row x 7 raw file bytes (63-bit field only)
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_emit[2] = {nullptr, nullptr},
                ev_out[2] = {nullptr, nullptr};
    size_t blocks = 0;              // blocks processed so far (selects the staging buffer)
    struct Pending { bool live = false; size_t row0 = 0, nr = 0; int buf = 0; } pending;  // D2H'd block awaiting its scatter
    uint64_t *d_pend[2] = {nullptr, nullptr};  // encoded rows not yet consumed by the column digests
    int cur = 0;
    size_t pend_rows = 0, pend_base = 0;
    uint8_t *d_cvs = nullptr;       // [chunk][column][32]
    uint64_t chunks_cap = 0, chunks_done = 0;
    uint8_t *d_hashes = nullptr;
    uint64_t *d_col[2] = {nullptr, nullptr};  // [n_cols][block_rows] canonical reprs (column-major emit)
    uint8_t *h_col[2] = {nullptr, nullptr};   // pinned mirrors of d_col
    uint8_t *sink = nullptr;        // host image of the encoded file (may be an mmap), column c at c*sink_cap*w
    size_t sink_cap = 0;
    size_t rows_total = 0, elems_total = 0;
    bool ragged = false, finished = false;
    std::mutex mu;
};

namespace {

void stream_release(lcpc_stream *s) {
    if (!s) return;
    lcpc_plan *plan = s->plan;
    if (plan) {
        cudaStream_t st = plan->ctx->stream;
        cudaSetDevice(plan->ctx->device);
        cudaStreamSynchronize(st);
        if (plan->ctx->s_in) cudaStreamSynchronize(plan->ctx->s_in);
        if (plan->ctx->s_out) cudaStreamSynchronize(plan->ctx->s_out);
        for (void *p : {(void *)s->d_in[0], (void *)s->d_in[1], (void *)s->d_bytes[0], (void *)s->d_bytes[1], (void *)s->d_pend[0],
                        (void *)s->d_pend[1], (void *)s->d_cvs, (void *)s->d_hashes, (void *)s->d_col[0], (void *)s->d_col[1]})
            if (p) cudaFree(p);
        for (int i = 0; i < 2; i++) {
            if (s->h_col[i]) cudaFreeHost(s->h_col[i]);
            for (cudaEvent_t e : {s->ev_in[i], s->ev_comp[i], s->ev_emit[i], s->ev_out[i]})
                if (e) cudaEventDestroy(e);
        }
    }
    delete s;
    if (plan) lcpc_plan_destroy(plan);  // drops the stream's reference (handles are reference counted)
}

// rows [row0, row0 + nr) of the column-major block in h_col[buf] -> the sink image
void scatter_to_sink(lcpc_stream *s, size_t row0, size_t nr, int buf) {
    const size_t w = s->wbytes, n_cols = s->n_cols;
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const unsigned nt = n_cols * nr * w < (1u << 20) ? 1u : hw;
    auto work = [&](unsigned t) {
        for (size_t c = t; c < n_cols; c += nt)
            memcpy(s->sink + (c * s->sink_cap + row0) * w, s->h_col[buf] + c * nr * w, nr * w);
    };
    if (nt == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto &x : th) x.join();
}

// the block whose device-to-host copy was enqueued earlier: wait for it and scatter it into the sink
int32_t flush_pending(lcpc_stream *s) {
    if (!s->pending.live) return LCPC_OK;
    CU(cudaEventSynchronize(s->ev_out[s->pending.buf]));
    scatter_to_sink(s, s->pending.row0, s->pending.nr, s->pending.buf);
    s->pending.live = false;
    return LCPC_OK;
}

// one block of rows is in d_in[b] (padded with zeros; ready on the compute stream): encode, emit, advance the digests
int32_t process_block(lcpc_stream *s, size_t nr, int b) {
    lcpc_plan *plan = s->plan;
    lcpc_ctx *ctx = plan->ctx;
    const int L = s->L;
    const size_t n_cols = s->n_cols, w = s->wbytes;
    uint64_t *pend = s->d_pend[s->cur];
    uint64_t *dst = pend + s->pend_rows * n_cols * L;
    int32_t rc = encode_dev(plan, s->d_in[b], nr, dst);
    if (rc != LCPC_OK) return rc;
    CU(cudaEventRecord(s->ev_comp[b], ctx->stream));  // staging buffer b may be refilled
    if (s->sink) {
        if (s->rows_total + nr > s->sink_cap) return fail(LCPC_ERR_TOO_BIG, "stream: sink row capacity exceeded");
        // h_col[b] / d_col[b] were last used two blocks ago; that block's scatter is flushed before this point
        CU(emit_colmajor(s->fid, dst, nr, n_cols, n_cols, s->d_col[b], nr, ctx->lc()));
        CU(cudaEventRecord(s->ev_emit[b], ctx->stream));
        CU(cudaStreamWaitEvent(ctx->s_out, s->ev_emit[b], 0));
        CU(cudaMemcpyAsync(s->h_col[b], s->d_col[b], n_cols * nr * w, cudaMemcpyDeviceToHost, ctx->s_out));
        CU(cudaEventRecord(s->ev_out[b], ctx->s_out));
    }
    s->pend_rows += nr;
    s->rows_total += nr;
    // chunks that are complete and certainly not the last one: 1024*(c+1) < bytes so far
    const uint64_t bytes = 32 + (uint64_t)s->rows_total * w;
    const uint64_t chunk_end = (bytes - 1) / 1024;
    if (chunk_end > s->chunks_done) {
        if (chunk_end > s->chunks_cap) return fail(LCPC_ERR_TOO_BIG, "stream: chunk store exceeded");
        CU(hash_chunk_range(s->fid, pend, (int64_t)s->pend_base, s->rows_total, n_cols, n_cols, s->chunks_done, chunk_end,
                            UINT64_MAX, UINT64_MAX, s->d_cvs, ctx->lc()));
        s->chunks_done = chunk_end;
        // rows below the first byte of the next chunk are finished with
        const uint64_t first_byte = 1024 * chunk_end;
        const size_t new_base = first_byte > 32 ? (size_t)((first_byte - 32) / w) : 0;
        const size_t keep = s->rows_total - new_base;
        if (keep)
            CU(cudaMemcpyAsync(s->d_pend[s->cur ^ 1], pend + (new_base - s->pend_base) * n_cols * L, keep * n_cols * w,
                               cudaMemcpyDeviceToDevice, ctx->stream));
        s->cur ^= 1;
        s->pend_base = new_base;
        s->pend_rows = keep;
    }
    if (s->sink) {
        // the previous block's copy has had this block's GPU work to hide behind; scatter it now, then queue ours
        int32_t rc2 = flush_pending(s);
        if (rc2 != LCPC_OK) return rc2;
        s->pending.live = true;
        s->pending.row0 = s->rows_total - nr;
        s->pending.nr = nr;
        s->pending.buf = b;
    }
    return LCPC_OK;
}

int32_t push_common(lcpc_stream *s, const void *data, size_t n_units, bool bytes) {
    if (!s || (!data && n_units)) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    lcpc_plan *plan = s->plan;
    lcpc_ctx *ctx = plan->ctx;
    std::lock_guard<std::mutex> g0(s->mu);
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    if (s->finished) return fail(LCPC_ERR_INVALID_ARG, "stream: already finished");
    if (n_units == 0) return LCPC_OK;
    if (s->ragged) return fail(LCPC_ERR_DIMS, "stream: only the final push may end inside a row");
    if (bytes && s->fid != FT63) return fail(LCPC_ERR_INVALID_ARG, "byte packing is defined for the 63-bit field only");
    CU(cudaSetDevice(ctx->device));
    const size_t unit_per_row = bytes ? s->n_per_row * 7 : s->n_per_row;  // bytes or elements in a full row
    const size_t usize = bytes ? 1 : s->wbytes;
    // all or nothing: a push that would overflow the stream is rejected before any of it is consumed
    if (s->rows_total + (n_units + unit_per_row - 1) / unit_per_row > s->max_rows)
        return fail(LCPC_ERR_TOO_BIG, "stream: more rows than max_rows");
    if (!ctx->s_in) CU(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    if (!ctx->s_out) CU(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    size_t off = 0;
    while (off < n_units) {
        const size_t take = std::min(n_units - off, s->block_rows * unit_per_row);
        const size_t nr = (take + unit_per_row - 1) / unit_per_row;
        const size_t n_elems = bytes ? (take + 6) / 7 : take;
        if (take % unit_per_row) s->ragged = true;
        const uint8_t *src = static_cast<const uint8_t *>(data) + off * usize;
        const int b = (int)(s->blocks & 1);
        // copy stream: wait until the compute stream has consumed what staging buffer b held (block k-2)
        if (s->blocks >= 2) CU(cudaStreamWaitEvent(ctx->s_in, s->ev_comp[b], 0));
        if (bytes) CU(cudaMemcpyAsync(s->d_bytes[b], src, take, cudaMemcpyHostToDevice, ctx->s_in));
        else CU(cudaMemcpyAsync(s->d_in[b], src, take * usize, cudaMemcpyHostToDevice, ctx->s_in));
        CU(cudaEventRecord(s->ev_in[b], ctx->s_in));
        CU(cudaStreamWaitEvent(ctx->stream, s->ev_in[b], 0));
        if (bytes) CU(pack_bytes7(s->d_bytes[b], take, s->d_in[b], ctx->lc()));
        if (n_elems < nr * s->n_per_row)  // zero fill of the last row (lib.rs:665-674; data_field.rs:38-46)
            CU(cudaMemsetAsync(s->d_in[b] + n_elems * s->L, 0, (nr * s->n_per_row - n_elems) * s->wbytes, ctx->stream));
        int32_t rc = process_block(s, nr, b);
        if (rc != LCPC_OK) return rc;
        s->blocks++;
        s->elems_total += n_elems;
        off += take;
    }
    // the caller's buffer may be reused after return: every copy out of it has to be complete
    CU(cudaStreamSynchronize(ctx->s_in));
    return LCPC_OK;
}

}  // namespace

extern "C" {

int32_t lcpc_stream_begin(lcpc_plan *plan, size_t max_rows, size_t block_rows, uint8_t *sink, size_t sink_row_capacity,
                          lcpc_stream **out) {
    plan = primary(plan);  // streaming runs on the first device of a multi-device context
    if (!plan || !out) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (max_rows == 0) return fail(LCPC_ERR_DIMS, "stream: max_rows must be positive");
    if (sink && sink_row_capacity < 1) return fail(LCPC_ERR_DIMS, "stream: sink row capacity must be positive");
    lcpc_ctx *ctx = plan->ctx;
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    lcpc_stream *s = new (std::nothrow) lcpc_stream;
    if (!s) return fail(LCPC_ERR_NOMEM, "host allocation failed");
    s->plan = plan;
    plan->refs.fetch_add(1);
    s->fid = plan->fid;
    s->L = limbs_of(plan->fid);
    s->wbytes = (size_t)s->L * 8;
    s->n_per_row = plan->n_per_row;
    s->n_cols = plan->n_cols;
    s->np2 = next_pow2(plan->n_cols);
    s->max_rows = max_rows;
    if (block_rows == 0) {  // about 256 MiB of encoded rows per block
        block_rows = std::max<size_t>(1, ((size_t)256 << 20) / (s->n_cols * s->wbytes));
    }
    s->block_rows = std::min(block_rows, max_rows);
    s->pend_cap = s->block_rows + 1024 / s->wbytes + 4;
    s->chunks_cap = hash_leaf_chunks(s->fid, max_rows);
    s->sink = sink;
    s->sink_cap = sink_row_capacity;
    auto body = [&]() -> int32_t {
        for (int i = 0; i < 2; i++) {
            CU(cudaMalloc((void **)&s->d_in[i], s->block_rows * s->n_per_row * s->wbytes));
            if (s->fid == FT63) CU(cudaMalloc((void **)&s->d_bytes[i], s->block_rows * s->n_per_row * 7 + 8));
            CU(cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s->ev_comp[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s->ev_emit[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s->ev_out[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < 2; i++) CU(cudaMalloc((void **)&s->d_pend[i], s->pend_cap * s->n_cols * s->wbytes));
        CU(cudaMalloc((void **)&s->d_cvs, (size_t)s->chunks_cap * s->n_cols * 32));
        CU(cudaMalloc((void **)&s->d_hashes, (2 * s->np2 - 1) * 32));
        if (sink) {
            for (int i = 0; i < 2; i++) {
                CU(cudaMalloc((void **)&s->d_col[i], s->block_rows * s->n_cols * s->wbytes));
                CU(cudaMallocHost((void **)&s->h_col[i], s->block_rows * s->n_cols * s->wbytes));
            }
        }
        return LCPC_OK;
    };
    int32_t rc = s->np2 ? body() : fail(LCPC_ERR_TOO_BIG, "n_cols is too large for this encoding");
    if (rc != LCPC_OK) {
        stream_release(s);
        return rc;
    }
    *out = s;
    return LCPC_OK;
}

int32_t lcpc_stream_push_elems_host(lcpc_stream *s, const uint64_t *elems, size_t n_elems) {
    return push_common(s, elems, n_elems, false);
}

int32_t lcpc_stream_push_bytes_host(lcpc_stream *s, const uint8_t *bytes, size_t n_bytes) {
    return push_common(s, bytes, n_bytes, true);
}

int32_t lcpc_stream_finish(lcpc_stream *s, uint8_t *hashes_out, size_t *n_rows_out) {
    if (!s) return fail(LCPC_ERR_INVALID_ARG, "null stream");
    lcpc_plan *plan = s->plan;
    lcpc_ctx *ctx = plan->ctx;
    std::lock_guard<std::mutex> g0(s->mu);
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (s->rows_total == 0) return fail(LCPC_ERR_DIMS, "cannot commit to zero coefficients");
    {
        int32_t rc = flush_pending(s);
        if (rc != LCPC_OK) return rc;
    }
    if (!s->finished) {
        const uint64_t total = hash_leaf_bytes(s->fid, s->rows_total), nc = hash_leaf_chunks(s->fid, s->rows_total);
        CU(hash_chunk_range(s->fid, s->d_pend[s->cur], (int64_t)s->pend_base, s->rows_total, s->n_cols, s->n_cols,
                            s->chunks_done, nc, total, nc, s->d_cvs, ctx->lc()));
        s->chunks_done = nc;
        // leaf merge + the whole tree in one launch (a single-chunk leaf is its own chaining value: copied into place)
        if (nc <= 1) CU(hash_merge(s->d_cvs, s->n_cols, nc, s->d_hashes, ctx->lc()));
        unsigned *tk = nullptr;
        CU(ctx->tickets(1, &tk));
        CU(merge_tree(s->d_cvs, s->n_cols, nc, s->d_hashes, s->np2, tk, ctx->lc()));
        s->finished = true;
    }
    if (hashes_out) CU(cudaMemcpyAsync(hashes_out, s->d_hashes, (2 * s->np2 - 1) * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (n_rows_out) *n_rows_out = s->rows_total;
    return LCPC_OK;
}

void lcpc_stream_free(lcpc_stream *s) { stream_release(s); }

// rows [row0, row0 + n_rows) <- coeff_rows; the commitment may grow (append) but never shrink
static int32_t commit_write_rows(lcpc_commit *c, size_t row0, size_t n_rows, const uint64_t *coeff_rows, bool allow_grow,
                                 uint64_t *comm_rows_out, uint8_t *hashes_out) {
    if (!c || !coeff_rows) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (!c->shards.empty()) return fail(LCPC_ERR_INVALID_ARG, "row edits are not supported on a multi-device commitment");
    lcpc_plan *plan = c->plan;
    lcpc_ctx *ctx = plan->ctx;
    std::lock_guard<std::mutex> g0(c->mu);
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    // the range is checked under the commit's lock: a concurrent append may have changed n_rows
    if (n_rows == 0 || row0 + n_rows < row0 || row0 > c->n_rows) return fail(LCPC_ERR_DIMS, "row range outside the commitment");
    if (!allow_grow && row0 + n_rows > c->n_rows) return fail(LCPC_ERR_DIMS, "row range outside the commitment");
    if (allow_grow && row0 + n_rows < c->n_rows) return fail(LCPC_ERR_DIMS, "an append must reach the end of the commitment");
    CU(cudaSetDevice(ctx->device));
    const int L = limbs_of(plan->fid);
    const size_t w = (size_t)L * 8, npr = c->n_per_row, n_cols = c->n_cols;
    const size_t old_rows = c->n_rows, new_rows = std::max(old_rows, row0 + n_rows);
    const uint64_t old_nc = hash_leaf_chunks(plan->fid, old_rows);
    const uint64_t total = hash_leaf_bytes(plan->fid, new_rows), nc = hash_leaf_chunks(plan->fid, new_rows);
    // Growth is staged: the grown matrices and chaining-value store are built beside the resident ones and published to
    // the handle only after encode, hashing and the tree have succeeded, so a failure leaves the commitment as it was.
    const bool grow = new_rows > old_rows;
    uint64_t *d_coeffs = c->d_coeffs, *d_comm = c->d_comm;
    uint8_t *d_cvs = c->d_cvs, *d_hashes = c->d_hashes;
    const size_t tree_bytes = (2 * c->np2 - 1) * 32;
    auto drop_staged = [&]() {
        if (!grow) return;
        if (d_coeffs && d_coeffs != c->d_coeffs) cudaFreeAsync(d_coeffs, ctx->stream);
        if (d_comm && d_comm != c->d_comm) cudaFreeAsync(d_comm, ctx->stream);
        if (d_cvs && d_cvs != c->d_cvs) cudaFreeAsync(d_cvs, ctx->stream);
        if (d_hashes && d_hashes != c->d_hashes) cudaFreeAsync(d_hashes, ctx->stream);
    };
#define CUS(call)                                                  \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) {                                  \
            drop_staged();                                         \
            return ::lcpc::abi::cuda_fail(e__, #call);             \
        }                                                          \
    } while (0)
    if (grow) {
        // (the reference doubles the file's row capacity, encoded_file_writer.rs:429-462)
        d_coeffs = d_comm = nullptr;
        d_cvs = d_hashes = nullptr;
        CUS(cudaMallocAsync((void **)&d_coeffs, new_rows * npr * w, ctx->stream));
        CUS(cudaMallocAsync((void **)&d_comm, new_rows * n_cols * w, ctx->stream));
        CUS(cudaMallocAsync((void **)&d_hashes, tree_bytes, ctx->stream));
        CUS(cudaMemcpyAsync(d_coeffs, c->d_coeffs, old_rows * npr * w, cudaMemcpyDeviceToDevice, ctx->stream));
        CUS(cudaMemcpyAsync(d_comm, c->d_comm, old_rows * n_cols * w, cudaMemcpyDeviceToDevice, ctx->stream));
        CUS(cudaMemcpyAsync(d_hashes, c->d_hashes, tree_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        if (nc > 1) {
            CUS(cudaMallocAsync((void **)&d_cvs, (size_t)nc * n_cols * 32, ctx->stream));
            if (c->d_cvs && old_nc > 1)
                CUS(cudaMemcpyAsync(d_cvs, c->d_cvs, (size_t)old_nc * n_cols * 32, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    uint64_t *d_rows = d_coeffs + row0 * npr * L;
    uint64_t *d_enc = d_comm + row0 * n_cols * L;
    CUS(cudaMemcpyAsync(d_rows, coeff_rows, n_rows * npr * w, cudaMemcpyHostToDevice, ctx->stream));
    int32_t rc = encode_dev(plan, d_rows, n_rows, d_enc);
    if (rc != LCPC_OK) {
        drop_staged();
        return rc;
    }
    // an in-place edit that fails from here on leaves rows and tree out of step; re-hashing everything from the resident
    // matrix (a second update of the same rows) repairs it -- only growth can be made atomic without a full copy
    unsigned *tk = nullptr;
    CUS(ctx->tickets(1, &tk));
    if (nc <= 1 || !d_cvs) {
        DevBuf scratch;
        CUS(scratch.alloc(hash_scratch_bytes(plan->fid, new_rows, n_cols), ctx->stream));
        CUS(hash_columns(plan->fid, d_comm, new_rows, n_cols, n_cols, nullptr, d_hashes, scratch.as<uint8_t>(), ctx->lc()));
        CUS(merge_tree(nullptr, n_cols, 1, d_hashes, c->np2, tk, ctx->lc()));
    } else {
        // chunks that contain the written rows; when the leaf grew, also its former last chunk (it was hashed as a final,
        // possibly partial, possibly ROOT-flagged chunk)
        uint64_t c_lo = (32 + (uint64_t)row0 * w) / 1024;
        if (grow) c_lo = std::min<uint64_t>(c_lo, old_nc - 1);
        const uint64_t c_hi = (32 + (uint64_t)(row0 + n_rows) * w - 1) / 1024 + 1;
        CUS(hash_chunk_range(plan->fid, d_comm, 0, new_rows, n_cols, n_cols, c_lo, std::min(c_hi, nc), total, nc, d_cvs, ctx->lc()));
        CUS(merge_tree(d_cvs, n_cols, nc, d_hashes, c->np2, tk, ctx->lc()));
    }
    if (comm_rows_out) CUS(cudaMemcpyAsync(comm_rows_out, d_enc, n_rows * n_cols * w, cudaMemcpyDeviceToHost, ctx->stream));
    if (hashes_out) CUS(cudaMemcpyAsync(hashes_out, d_hashes, tree_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUS(cudaStreamSynchronize(ctx->stream));
#undef CUS
    if (grow) {  // publish
        cudaFreeAsync(c->d_coeffs, ctx->stream);
        cudaFreeAsync(c->d_comm, ctx->stream);
        cudaFreeAsync(c->d_hashes, ctx->stream);
        if (c->d_cvs) cudaFreeAsync(c->d_cvs, ctx->stream);
        c->d_coeffs = d_coeffs;
        c->d_comm = d_comm;
        c->d_hashes = d_hashes;
        c->d_cvs = d_cvs;
        c->n_rows = new_rows;
    }
    return LCPC_OK;
}

int32_t lcpc_commit_update_rows_host(lcpc_commit *c, size_t row0, size_t n_rows, const uint64_t *coeff_rows,
                                     uint64_t *comm_rows_out, uint8_t *hashes_out) {
    return commit_write_rows(c, row0, n_rows, coeff_rows, false, comm_rows_out, hashes_out);
}

int32_t lcpc_commit_append_rows_host(lcpc_commit *c, size_t row0, size_t n_rows, const uint64_t *coeff_rows,
                                     uint64_t *comm_rows_out, uint8_t *hashes_out) {
    return commit_write_rows(c, row0, n_rows, coeff_rows, true, comm_rows_out, hashes_out);
}

}  // extern "C"
