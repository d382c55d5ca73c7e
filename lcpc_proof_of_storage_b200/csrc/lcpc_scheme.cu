// prove / verify sequencing of lcpc-2d over the device-resident commitment, plus the
// host-side pieces they need (transcript, challenge expansion, code generation).
//
//   lcpc_prove   = lcpc-2d/src/lib.rs:1034-1123 (LcCommit::prove, :319)
//   lcpc_verify  = lcpc-2d/src/lib.rs:862-982   (LcEvalProof::verify, :547)
//
// All field arithmetic (folds, encodes, dot products, de-Montgomery for the transcript,
// leaf hashes and Merkle climbs) runs in CUDA kernels; the host only drives the
// transcript, expands challenges and compares results.
#include "lcpc_handles.h"
#include "lcpc_hostrand.h"

using namespace lcpc;
using namespace lcpc::abi;

struct lcpc_transcript {
    host::Transcript t;
    std::mutex mu;
    explicit lcpc_transcript(const uint8_t *label, size_t len) : t(label, len) {}
    lcpc_transcript(const lcpc_transcript &o) : t(o.t) {}
};

namespace {

// def_labels! (lcpc-2d/src/macros.rs:28-36): `$l` is not substituted inside the byte-string
// literal, so every encoding uses these literal six bytes.
const uint8_t LABEL_DT[] = {'$', 'l', '/', '/', 'D', 'T'};
const uint8_t LABEL_PR[] = {'$', 'l', '/', '/', 'P', 'R'};
const uint8_t LABEL_PE[] = {'$', 'l', '/', '/', 'P', 'E'};
const uint8_t LABEL_CO[] = {'$', 'l', '/', '/', 'C', 'O'};

// FieldHash::transcript_update for a vector: one message per element, canonical LE bytes
void transcript_update(host::Transcript &t, const uint8_t *label, const uint64_t *canon, size_t n, int L) {
    for (size_t i = 0; i < n; i++)
        t.append_message(label, 6, reinterpret_cast<const uint8_t *>(canon + i * L), (size_t)L * 8);
}

void expand_tensor(host::Transcript &t, int fid, size_t n_rows, uint64_t *out) {
    uint8_t key[32];
    t.challenge_bytes(LABEL_DT, 6, key, 32);
    host::ChaCha20Rng rng = host::ChaCha20Rng::from_seed(key);
    const int L = limbs_of(fid);
    for (size_t i = 0; i < n_rows; i++) host::field_random(fid, rng, out + i * L);
}

// Every element below the modulus?  The kernels assume reduced operands (ff_derive's from_repr rejects anything else,
// so the reference's verifier never sees such a value): to_canon(x + p) == to_canon(x) would give a non-canonical
// alias the same leaf hash and transcript bytes as the honest element while the dot products go wrong.
bool all_reduced(int fid, const uint64_t *v, size_t n) {
    const FieldConsts fc = field_consts(fid);
    const int L = fc.limbs;
    for (size_t i = 0; i < n; i++) {
        const uint64_t *a = v + i * L;
        bool lt = false;
        for (int l = L - 1; l >= 0; l--) {
            if (a[l] != fc.p[l]) {
                lt = a[l] < fc.p[l];
                break;
            }
        }
        if (!lt) return false;
    }
    return true;
}

void expand_columns(host::Transcript &t, size_t n_cols, size_t n, uint64_t *out) {
    uint8_t key[32];
    t.challenge_bytes(LABEL_CO, 6, key, 32);
    host::ChaCha20Rng rng = host::ChaCha20Rng::from_seed(key);
    for (size_t i = 0; i < n; i++) out[i] = rng.uniform(n_cols);
}

}  // namespace

extern "C" {

// ---- transcript ---------------------------------------------------------------------

int32_t lcpc_transcript_new(const uint8_t *label, size_t label_len, lcpc_transcript **out) {
    if (!out || (!label && label_len)) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    *out = new (std::nothrow) lcpc_transcript(label, label_len);
    return *out ? LCPC_OK : fail(LCPC_ERR_NOMEM, "host allocation failed");
}

int32_t lcpc_transcript_clone(const lcpc_transcript *t, lcpc_transcript **out) {
    if (!t || !out) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    *out = new (std::nothrow) lcpc_transcript(*t);
    return *out ? LCPC_OK : fail(LCPC_ERR_NOMEM, "host allocation failed");
}

int32_t lcpc_transcript_append_message(lcpc_transcript *t, const uint8_t *label, size_t label_len, const uint8_t *msg,
                                       size_t msg_len) {
    if (!t || (!label && label_len) || (!msg && msg_len)) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    t->t.append_message(label, label_len, msg, msg_len);
    return LCPC_OK;
}

int32_t lcpc_transcript_challenge_bytes(lcpc_transcript *t, const uint8_t *label, size_t label_len, uint8_t *dest,
                                        size_t dest_len) {
    if (!t || (!label && label_len) || (!dest && dest_len)) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    t->t.challenge_bytes(label, label_len, dest, dest_len);
    return LCPC_OK;
}

void lcpc_transcript_free(lcpc_transcript *t) { delete t; }

// ---- challenge expansion (exposed for tests and for callers that keep the transcript in Rust) ----

int32_t lcpc_random_field_vec(int32_t field, const uint8_t key[32], uint64_t *out, size_t n) {
    if (!valid_field(field) || !key || (!out && n)) return fail(LCPC_ERR_INVALID_ARG, "bad argument");
    host::ChaCha20Rng rng = host::ChaCha20Rng::from_seed(key);
    const int L = limbs_of(field);
    for (size_t i = 0; i < n; i++) host::field_random(field, rng, out + i * L);
    return LCPC_OK;
}

int32_t lcpc_random_columns(const uint8_t key[32], uint64_t n_cols, uint64_t *out, size_t n) {
    if (!key || (!out && n) || n_cols == 0) return fail(LCPC_ERR_INVALID_ARG, "bad argument");
    host::ChaCha20Rng rng = host::ChaCha20Rng::from_seed(key);
    for (size_t i = 0; i < n; i++) out[i] = rng.uniform(n_cols);
    return LCPC_OK;
}

// ---- proof-of-storage helpers -------------------------------------------------------------

int32_t lcpc_pos_choose_columns(uint64_t seed, size_t amount, size_t max_index, uint64_t *out, size_t *n_out) {
    if ((!out && amount) || !n_out) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    host::ChaCha20Rng rng = host::ChaCha20Rng::seed_from_u64(seed, 8);  // ChaCha8Rng (networking/client.rs:448)
    std::vector<uint64_t> cols;
    host::choose_multiple_indices(rng, amount, max_index, cols);
    for (size_t i = 0; i < cols.size(); i++) out[i] = cols[i];
    *n_out = cols.size();
    return LCPC_OK;
}

int32_t lcpc_verify_columns_host(lcpc_ctx *ctx, int32_t field, const uint64_t *columns, size_t n_rows, const uint8_t *paths,
                                 size_t path_len, const uint64_t *col_idx, size_t n, const uint8_t root[LCPC_DIGEST_BYTES],
                                 uint8_t *leaves_out, uint32_t *ok_out) {
    ctx = primary(ctx);
    if (!ctx || (!columns && n && n_rows) || (!col_idx && n && paths)) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (!valid_field(field)) return fail(LCPC_ERR_INVALID_ARG, "unknown field id");
    if (paths && (!root || !ok_out)) return fail(LCPC_ERR_INVALID_ARG, "paths given without root / ok_out");
    if (n == 0) return LCPC_OK;
    if (!all_reduced(field, columns, n * n_rows)) return fail(LCPC_ERR_INVALID_ARG, "column element not below the modulus");
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    const size_t wbytes = (size_t)limbs_of(field) * 8;
    DevBuf d_cols, d_leafidx, d_leaves, d_hs, d_paths, d_idx, d_root, d_ok;
    CU(d_cols.alloc(n * n_rows * wbytes, ctx->stream));
    CU(d_leafidx.alloc(n * 8, ctx->stream));
    CU(d_leaves.alloc(n * 32, ctx->stream));
    CU(d_hs.alloc(hash_scratch_bytes(field, n_rows, n), ctx->stream));
    CU(cudaMemcpyAsync(d_cols.p, columns, n * n_rows * wbytes, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<uint64_t> leafidx(n);
    for (size_t i = 0; i < n; i++) leafidx[i] = i * n_rows;  // column i is contiguous: a stride-1 "matrix"
    CU(cudaMemcpyAsync(d_leafidx.p, leafidx.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(hash_columns(field, d_cols.as<uint64_t>(), n_rows, 1, n, d_leafidx.as<uint64_t>(), d_leaves.as<uint8_t>(),
                    d_hs.as<uint8_t>(), ctx->lc()));
    if (leaves_out) CU(cudaMemcpyAsync(leaves_out, d_leaves.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    if (paths) {
        CU(d_paths.alloc(n * path_len * 32, ctx->stream));
        CU(d_idx.alloc(n * 8, ctx->stream));
        CU(d_root.alloc(32, ctx->stream));
        CU(d_ok.alloc(n * 4, ctx->stream));
        if (path_len) CU(cudaMemcpyAsync(d_paths.p, paths, n * path_len * 32, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_idx.p, col_idx, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_root.p, root, 32, cudaMemcpyHostToDevice, ctx->stream));
        CU(verify_paths(d_leaves.as<uint8_t>(), d_paths.as<uint8_t>(), (int)path_len, d_idx.as<uint64_t>(), n,
                        d_root.as<uint8_t>(), d_ok.as<uint32_t>(), ctx->lc()));
        CU(cudaMemcpyAsync(ok_out, d_ok.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return LCPC_OK;
}

// ---- Brakedown code generation ---------------------------------------------------------

int32_t lcpc_sdig_get_dims(int32_t code, uint64_t n_per_row, int32_t field, uint64_t *pre_dims, uint64_t *post_dims,
                           int32_t max_levels, int32_t *n_levels) {
    if (!valid_field(field) || !pre_dims || !post_dims || !n_levels) return fail(LCPC_ERR_INVALID_ARG, "bad argument");
    std::vector<host::SdigDims> pre, post;
    // matgen.rs:34: log2p = F::FLOG2 = NUM_BITS - 1
    if (!host::sdig_get_dims(code, n_per_row, (double)(field_consts(field).num_bits - 1), pre, post))
        return fail(LCPC_ERR_DIMS, "n_per_row must exceed the base-case length of a known code (1..6)");
    if ((int32_t)pre.size() > max_levels) return fail(LCPC_ERR_TOO_BIG, "more code levels than max_levels");
    for (size_t i = 0; i < pre.size(); i++) {
        pre_dims[3 * i] = pre[i].n; pre_dims[3 * i + 1] = pre[i].m; pre_dims[3 * i + 2] = pre[i].d;
        post_dims[3 * i] = post[i].n; post_dims[3 * i + 1] = post[i].m; post_dims[3 * i + 2] = post[i].d;
    }
    *n_levels = (int32_t)pre.size();
    return LCPC_OK;
}

int32_t lcpc_sdig_gen_level(int32_t field, uint64_t seed, uint64_t level, const uint64_t pre_dim[3],
                            const uint64_t post_dim[3], uint64_t *pre_indptr, uint64_t *pre_indices, uint64_t *pre_data,
                            uint64_t *post_indptr, uint64_t *post_indices, uint64_t *post_data) {
    if (!valid_field(field) || !pre_dim || !post_dim || !pre_indptr || !post_indptr)
        return fail(LCPC_ERR_INVALID_ARG, "bad argument");
    // matgen.rs:43-46: one ChaCha20 stream per level, precode first, then the postcode
    host::ChaCha20Rng rng = host::ChaCha20Rng::seed_from_u64(seed);
    rng.set_stream(level);
    host::sdig_gen_code(field, rng, {pre_dim[0], pre_dim[1], pre_dim[2]}, pre_indptr, pre_indices, pre_data);
    host::sdig_gen_code(field, rng, {post_dim[0], post_dim[1], post_dim[2]}, post_indptr, post_indices, post_data);
    return LCPC_OK;
}

double lcpc_sdig_dist(int32_t code) { return (code >= 1 && code <= 6) ? host::sdig_dist(code) : 0.0; }

// ---- prove ---------------------------------------------------------------------------------

int32_t lcpc_prove(lcpc_commit *c, const uint64_t *outer_tensor, size_t outer_len, size_t n_degree_tests,
                   size_t n_col_opens, lcpc_transcript *tr, uint64_t *p_eval_out, uint64_t *p_random_out,
                   uint64_t *col_idx_out, uint64_t *columns_out, uint8_t *paths_out) {
    if (!c || !outer_tensor || !tr || !p_eval_out || (!p_random_out && n_degree_tests) ||
        ((!columns_out || !paths_out) && n_col_opens))
        return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (outer_len != c->n_rows) return fail(LCPC_ERR_OUTER_TENSOR, "bad outer tensor size");  // lib.rs:1046-1048
    if (!c->shards.empty()) {
        // multi-device commitment: the same sequence with the sharded fold / open (lcpc_multi.cu); the canonical bytes for
        // the transcript are produced on the first device
        lcpc_ctx *mctx = c->plan->ctx, *p0 = primary(mctx);
        std::lock_guard<std::mutex> g(c->mu);
        std::lock_guard<std::mutex> g2(mctx->mu);
        std::lock_guard<std::mutex> g3(tr->mu);
        const int fid = c->plan->fid, L = limbs_of(fid);
        const size_t wbytes = (size_t)L * 8, n_rows = c->n_rows, npr = c->n_per_row;
        std::vector<uint64_t> tensor(n_rows * L), canon(npr * L);
        auto fold_one = [&](const uint64_t *t_host, const uint8_t *label, uint64_t *dst) -> int32_t {
            int32_t rc = multi::fold_host(c, 0, t_host, 1, dst);
            if (rc != LCPC_OK) return rc;
            CU(cudaSetDevice(p0->device));
            DevBuf d_p, d_pc;
            CU(d_p.alloc(npr * wbytes, p0->stream));
            CU(d_pc.alloc(npr * wbytes, p0->stream));
            CU(cudaMemcpyAsync(d_p.p, dst, npr * wbytes, cudaMemcpyHostToDevice, p0->stream));
            CU(to_canon(fid, d_p.as<uint64_t>(), npr, d_pc.as<uint64_t>(), p0->lc()));
            CU(cudaMemcpyAsync(canon.data(), d_pc.p, npr * wbytes, cudaMemcpyDeviceToHost, p0->stream));
            CU(cudaStreamSynchronize(p0->stream));
            transcript_update(tr->t, label, canon.data(), npr, L);
            return LCPC_OK;
        };
        for (size_t i = 0; i < n_degree_tests; i++) {
            expand_tensor(tr->t, fid, n_rows, tensor.data());
            int32_t rc = fold_one(tensor.data(), LABEL_PR, p_random_out + i * npr * L);
            if (rc != LCPC_OK) return rc;
        }
        int32_t rc = fold_one(outer_tensor, LABEL_PE, p_eval_out);
        if (rc != LCPC_OK) return rc;
        std::vector<uint64_t> cols(n_col_opens);
        expand_columns(tr->t, c->n_cols, n_col_opens, cols.data());
        if (col_idx_out) std::memcpy(col_idx_out, cols.data(), n_col_opens * sizeof(uint64_t));
        if (n_col_opens) return multi::open_columns_host(c, cols.data(), n_col_opens, columns_out, paths_out);
        return LCPC_OK;
    }
    lcpc_ctx *ctx = c->plan->ctx;
    std::lock_guard<std::mutex> g(c->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    std::lock_guard<std::mutex> g3(tr->mu);
    CU(cudaSetDevice(ctx->device));
    const int fid = c->plan->fid, L = limbs_of(fid);
    const size_t wbytes = (size_t)L * 8, n_rows = c->n_rows, npr = c->n_per_row;
    DevBuf d_t, d_p, d_pc, d_s;
    CU(d_t.alloc(n_rows * wbytes, ctx->stream));
    CU(d_p.alloc(npr * wbytes, ctx->stream));
    CU(d_pc.alloc(npr * wbytes, ctx->stream));
    CU(d_s.alloc(fold_scratch_bytes(fid, n_rows, npr, 1), ctx->stream));
    std::vector<uint64_t> tensor(n_rows * L), canon(npr * L);
    // one fold + transcript update; `dst` receives the Montgomery-form result
    auto fold_one = [&](const uint64_t *t_host, const uint8_t *label, uint64_t *dst) -> int32_t {
        CU(cudaMemcpyAsync(d_t.p, t_host, n_rows * wbytes, cudaMemcpyHostToDevice, ctx->stream));
        CU(fold(fid, c->d_coeffs, n_rows, npr, npr, d_t.as<uint64_t>(), 1, d_p.as<uint64_t>(), d_s.as<uint64_t>(), ctx->lc()));
        CU(to_canon(fid, d_p.as<uint64_t>(), npr, d_pc.as<uint64_t>(), ctx->lc()));
        CU(cudaMemcpyAsync(dst, d_p.p, npr * wbytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(canon.data(), d_pc.p, npr * wbytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        transcript_update(tr->t, label, canon.data(), npr, L);
        return LCPC_OK;
    };
    // lib.rs:1054-1080: degree tests are sequential (tensor i+1 depends on p_random_i through the transcript)
    for (size_t i = 0; i < n_degree_tests; i++) {
        expand_tensor(tr->t, fid, n_rows, tensor.data());
        int32_t rc = fold_one(tensor.data(), LABEL_PR, p_random_out + i * npr * L);
        if (rc != LCPC_OK) return rc;
    }
    // lib.rs:1083-1098
    int32_t rc = fold_one(outer_tensor, LABEL_PE, p_eval_out);
    if (rc != LCPC_OK) return rc;
    // lib.rs:1101-1115
    std::vector<uint64_t> cols(n_col_opens);
    expand_columns(tr->t, c->n_cols, n_col_opens, cols.data());
    if (col_idx_out) std::memcpy(col_idx_out, cols.data(), n_col_opens * sizeof(uint64_t));
    if (n_col_opens) {
        int depth = 0;
        while (((size_t)1 << depth) < c->np2) depth++;
        DevBuf d_cols, d_out, d_paths;
        CU(d_cols.alloc(n_col_opens * 8, ctx->stream));
        CU(d_out.alloc(n_col_opens * n_rows * wbytes, ctx->stream));
        CU(d_paths.alloc(n_col_opens * (size_t)depth * 32, ctx->stream));
        CU(cudaMemcpyAsync(d_cols.p, cols.data(), n_col_opens * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(gather_columns(fid, c->d_comm, n_rows, c->n_cols, d_cols.as<uint64_t>(), n_col_opens, d_out.as<uint64_t>(), ctx->lc()));
        CU(gather_paths(c->d_hashes, c->np2, d_cols.as<uint64_t>(), n_col_opens, d_paths.as<uint8_t>(), ctx->lc()));
        CU(cudaMemcpyAsync(columns_out, d_out.p, n_col_opens * n_rows * wbytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (depth) CU(cudaMemcpyAsync(paths_out, d_paths.p, n_col_opens * (size_t)depth * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return LCPC_OK;
}

// ---- verify --------------------------------------------------------------------------------

int32_t lcpc_verify(lcpc_plan *plan, const uint8_t root[LCPC_DIGEST_BYTES], const uint64_t *outer_tensor, size_t outer_len,
                    const uint64_t *inner_tensor, size_t inner_len, size_t proof_n_cols, const uint64_t *p_eval,
                    size_t n_per_row, const uint64_t *p_random, size_t n_p_random, const uint64_t *columns, size_t n_rows,
                    const uint8_t *paths, size_t path_len, size_t n_columns, size_t n_col_opens, size_t n_degree_tests,
                    lcpc_transcript *tr, uint64_t *result_out) {
    plan = primary(plan);  // verification is replicas-only work: the first device of a multi-device context
    if (!plan || !root || !outer_tensor || !inner_tensor || !p_eval || !tr || !result_out)
        return fail(LCPC_ERR_INVALID_ARG, "null argument");
    // lib.rs:874-891 argument checks, in the reference's order
    if (n_col_opens != n_columns || n_col_opens == 0) return fail(LCPC_VERR_NUM_COL_OPENS, "wrong number of column openings in proof");
    if (!columns || (!paths && path_len)) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    if (inner_len != n_per_row) return fail(LCPC_VERR_INNER_TENSOR, "bad inner tensor size for this proof");
    if (outer_len != n_rows) return fail(LCPC_VERR_OUTER_TENSOR, "bad outer tensor size for this proof");
    if (!(n_per_row < proof_n_cols) || n_per_row != plan->n_per_row || proof_n_cols != plan->n_cols)
        return fail(LCPC_VERR_ENCODING_DIMS, "incorrect encoding dimensions");
    if (n_p_random < n_degree_tests) return fail(LCPC_ERR_INVALID_ARG, "proof holds fewer p_random vectors than degree tests");
    if (n_degree_tests && !p_random) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    // untrusted limbs: the reference's deserialiser (PrimeField::from_repr) refuses elements that are not below the
    // modulus, so its verifier never computes on them; here they are refused before any kernel sees them
    if (!all_reduced(plan->fid, outer_tensor, outer_len) || !all_reduced(plan->fid, inner_tensor, inner_len) ||
        !all_reduced(plan->fid, p_eval, n_per_row) || !all_reduced(plan->fid, p_random, n_degree_tests * n_per_row))
        return fail(LCPC_ERR_INVALID_ARG, "field element not below the modulus in the proof or the tensors");
    if (!all_reduced(plan->fid, columns, n_columns * n_rows))
        return fail(LCPC_VERR_COLUMN_EVAL, "column eval invalid (element not below the modulus)");
    lcpc_ctx *ctx = plan->ctx;
    std::lock_guard<std::mutex> g(plan->mu);
    std::lock_guard<std::mutex> g2(ctx->mu);
    std::lock_guard<std::mutex> g3(tr->mu);
    CU(cudaSetDevice(ctx->device));
    const int fid = plan->fid, L = limbs_of(fid);
    const size_t wbytes = (size_t)L * 8, n_cols = plan->n_cols, nt = n_degree_tests + 1;
    {
        size_t np2 = next_pow2(n_cols);
        size_t depth = 0;
        while (((size_t)1 << depth) < np2) depth++;
        // a path of the wrong length can never hash to the root; the reference would simply
        // compute a different digest (lib.rs:999-1011), so report it the same way
        if (path_len != depth) return fail(LCPC_VERR_COLUMN_PATH, "column path invalid");
    }
    // device copies of the polynomials to encode: rows 0..n_dt-1 = p_random_i, row n_dt = p_eval
    DevBuf d_polys, d_canon, d_rows;
    CU(d_polys.alloc(nt * n_per_row * wbytes, ctx->stream));
    CU(d_canon.alloc(nt * n_per_row * wbytes, ctx->stream));
    CU(d_rows.alloc(nt * n_cols * wbytes, ctx->stream));
    if (n_degree_tests)
        CU(cudaMemcpyAsync(d_polys.p, p_random, n_degree_tests * n_per_row * wbytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_polys.as<uint64_t>() + n_degree_tests * n_per_row * L, p_eval, n_per_row * wbytes,
                       cudaMemcpyHostToDevice, ctx->stream));
    CU(to_canon(fid, d_polys.as<uint64_t>(), nt * n_per_row, d_canon.as<uint64_t>(), ctx->lc()));
    std::vector<uint64_t> canon(nt * n_per_row * L);
    CU(cudaMemcpyAsync(canon.data(), d_canon.p, nt * n_per_row * wbytes, cudaMemcpyDeviceToHost, ctx->stream));
    // lib.rs:912-918, 944-950: encode p_random_i and p_eval as single rows
    int32_t rc = encode_dev(plan, d_polys.as<uint64_t>(), nt, d_rows.as<uint64_t>());
    if (rc != LCPC_OK) return fail(LCPC_VERR_ENCODE, std::string("encoding error: ") + lcpc_last_error());
    CU(cudaStreamSynchronize(ctx->stream));
    // lib.rs:898-941: replay the transcript
    std::vector<uint64_t> tensors(nt * n_rows * L);
    for (size_t i = 0; i < n_degree_tests; i++) {
        expand_tensor(tr->t, fid, n_rows, tensors.data() + i * n_rows * L);
        transcript_update(tr->t, LABEL_PR, canon.data() + i * n_per_row * L, n_per_row, L);
    }
    transcript_update(tr->t, LABEL_PE, canon.data() + n_degree_tests * n_per_row * L, n_per_row, L);
    std::memcpy(tensors.data() + n_degree_tests * n_rows * L, outer_tensor, n_rows * wbytes);
    std::vector<uint64_t> cols(n_col_opens);
    expand_columns(tr->t, n_cols, n_col_opens, cols.data());
    // lib.rs:953-974: per opened column, degree-test values, evaluation value, Merkle path
    DevBuf d_tens, d_colsv, d_idx, d_dots, d_expect, d_paths, d_leaves, d_leafidx, d_root, d_ok, d_hs;
    CU(d_tens.alloc(nt * n_rows * wbytes, ctx->stream));
    CU(d_colsv.alloc(n_col_opens * n_rows * wbytes, ctx->stream));
    CU(d_idx.alloc(n_col_opens * 8, ctx->stream));
    CU(d_dots.alloc(n_col_opens * nt * wbytes, ctx->stream));
    CU(d_expect.alloc(n_col_opens * nt * wbytes, ctx->stream));
    CU(d_paths.alloc(n_col_opens * path_len * 32, ctx->stream));
    CU(d_leaves.alloc(n_col_opens * 32, ctx->stream));
    CU(d_leafidx.alloc(n_col_opens * 8, ctx->stream));
    CU(d_root.alloc(32, ctx->stream));
    CU(d_ok.alloc(n_col_opens * 4, ctx->stream));
    CU(d_hs.alloc(hash_scratch_bytes(fid, n_rows, n_col_opens), ctx->stream));
    CU(cudaMemcpyAsync(d_tens.p, tensors.data(), nt * n_rows * wbytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_colsv.p, columns, n_col_opens * n_rows * wbytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_idx.p, cols.data(), n_col_opens * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (path_len) CU(cudaMemcpyAsync(d_paths.p, paths, n_col_opens * path_len * 32, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_root.p, root, 32, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<uint64_t> leafidx(n_col_opens);
    for (size_t i = 0; i < n_col_opens; i++) leafidx[i] = i * n_rows;  // opened column i is contiguous: stride-1 "rows"
    CU(cudaMemcpyAsync(d_leafidx.p, leafidx.data(), n_col_opens * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(column_dots(fid, d_colsv.as<uint64_t>(), n_rows, n_col_opens, d_tens.as<uint64_t>(), nt, d_dots.as<uint64_t>(), ctx->lc()));
    // expected[i][t] = encoded_row_t[cols[i]]
    CU(gather_columns(fid, d_rows.as<uint64_t>(), nt, n_cols, d_idx.as<uint64_t>(), n_col_opens, d_expect.as<uint64_t>(), ctx->lc()));
    CU(hash_columns(fid, d_colsv.as<uint64_t>(), n_rows, 1, n_col_opens, d_leafidx.as<uint64_t>(), d_leaves.as<uint8_t>(),
                    d_hs.as<uint8_t>(), ctx->lc()));
    CU(verify_paths(d_leaves.as<uint8_t>(), d_paths.as<uint8_t>(), (int)path_len, d_idx.as<uint64_t>(), n_col_opens,
                    d_root.as<uint8_t>(), d_ok.as<uint32_t>(), ctx->lc()));
    // lib.rs:977-981: sum_j inner[j] * p_eval[j] -- a fold of the n_per_row x 1 "matrix" p_eval by inner
    DevBuf d_inner, d_res, d_fs;
    CU(d_inner.alloc(n_per_row * wbytes, ctx->stream));
    CU(d_res.alloc(wbytes, ctx->stream));
    CU(d_fs.alloc(fold_scratch_bytes(fid, n_per_row, 1, 1), ctx->stream));
    CU(cudaMemcpyAsync(d_inner.p, inner_tensor, n_per_row * wbytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(fold(fid, d_polys.as<uint64_t>() + n_degree_tests * n_per_row * L, n_per_row, 1, 1, d_inner.as<uint64_t>(), 1,
            d_res.as<uint64_t>(), d_fs.as<uint64_t>(), ctx->lc()));
    std::vector<uint64_t> dots(n_col_opens * nt * L), expect(n_col_opens * nt * L);
    std::vector<uint32_t> ok(n_col_opens);
    CU(cudaMemcpyAsync(dots.data(), d_dots.p, dots.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(expect.data(), d_expect.p, expect.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ok.data(), d_ok.p, ok.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(result_out, d_res.p, wbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < n_col_opens; i++) {
        bool rand_ok = true;
        for (size_t t = 0; t < n_degree_tests; t++)
            rand_ok &= std::memcmp(&dots[(i * nt + t) * L], &expect[(i * nt + t) * L], wbytes) == 0;
        const bool eval_ok = std::memcmp(&dots[(i * nt + n_degree_tests) * L], &expect[(i * nt + n_degree_tests) * L], wbytes) == 0;
        // match (rand, eval, path) at lib.rs:968-973
        if (!rand_ok) return fail(LCPC_VERR_COLUMN_DEGREE, "column degree check failed");
        if (!eval_ok) return fail(LCPC_VERR_COLUMN_EVAL, "column eval invalid");
        if (!ok[i]) return fail(LCPC_VERR_COLUMN_PATH, "column path invalid");
    }
    return LCPC_OK;
}

}  // extern "C"
