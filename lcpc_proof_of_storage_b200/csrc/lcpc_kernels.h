// Internal launcher interface between the CUDA kernels (lcpc_*.cu) and the C ABI
// (lcpc_api.cu).  Everything takes raw device pointers and a stream; `launches`
// counts kernels enqueued.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

namespace lcpc {

constexpr int MAX_LIMBS = 4;

// Per-kernel device timing (CUDA events around every launch), off unless enabled on the
// context; bench.py reads it to get the dominant kernel's duration inside the timed step.
struct KernelTimer;
void timer_begin(KernelTimer *t, const char *name, cudaStream_t s);
void timer_end(KernelTimer *t, cudaStream_t s);

// Where and how a launcher enqueues: stream, launch counter, optional timer.
struct Launch {
    cudaStream_t s;
    uint64_t *count;
    KernelTimer *timer;
    void begin(const char *name) const {
        if (timer) timer_begin(timer, name, s);
    }
    void end() const {
        ++*count;
        if (timer) timer_end(timer, s);
    }
};

// w16^e, e = 0..7 (Montgomery), passed to NTT kernels by value (constant bank).
struct SmallTwRaw {
    uint64_t w[8][MAX_LIMBS];
};

struct NttPass {
    int kind;        // 0 = strided register-radix pass over global memory, 1 = shared-memory block pass
    int bits;        // radix bits R (kind 0) or block bits LB (kind 1)
    int log_sub;     // log2 of the sub-transform size this pass starts from
    size_t tw_off;   // element offset of this pass's twiddle table in d_tw
};

struct NttPlan {
    int fid = 0;
    int log_n = 0;
    size_t n = 0;
    std::vector<NttPass> passes;
    uint64_t *d_tw = nullptr;  // all twiddle tables, device
    size_t tw_elems = 0;
    SmallTwRaw stw{};
    // inverse transform: the same tables built from w^-1 (second half of the d_tw allocation), w16^-e, n^-1
    uint64_t *d_tw_inv = nullptr;
    SmallTwRaw stw_inv{};
    uint64_t ninv[MAX_LIMBS] = {0, 0, 0, 0};
};

// Builds twiddle tables for a 2^log_n-point transform.  root_mont: LIMBS words (host), or
// nullptr for ROOT_OF_UNITY^(2^(S-log_n)).
cudaError_t ntt_plan_build(NttPlan &plan, int fid, int log_n, const uint64_t *root_mont, const Launch &lc);
void ntt_plan_free(NttPlan &plan);

// Destination of a fused encode + re-shard (one process per GPU): the final pass of the transform
// stores each row block straight into the column-block matrix of the rank that will hash those
// columns -- local HBM for this rank's own block, peer HBM over NVLink for the others.  Rank g owns
// columns [g << log_cb, (g+1) << log_cb); its matrix is [n_rows_total][1 << log_cb] row-major.
struct ScatterDst {
    uint64_t *base[16];
    int log_cb;
    uint64_t row0;  // global index of this rank's first row
};

// n_rows independent transforms.  src rows have stride src_stride and src_valid leading
// valid elements (the rest of each row reads as zero); dst rows have stride n.
// src == dst (with src_stride == n) is the in-place case.
cudaError_t ntt_encode(const NttPlan &plan, const uint64_t *src, size_t src_stride, size_t src_valid,
                       uint64_t *dst, size_t n_rows, const Launch &lc, const ScatterDst *scatter = nullptr);

// n_rows inverse transforms in place (row stride n): fffft's ifft_oi -- bit-reversed input, in-order output, 1/n scale
cudaError_t ntt_decode(const NttPlan &plan, uint64_t *data, size_t n_rows, const Launch &lc);

// Column hashing.  d_col_idx == nullptr hashes columns [0, n_cols); otherwise column
// d_col_idx[j] for j < n_cols.  d_cv_scratch needs hash_scratch_bytes().
size_t hash_scratch_bytes(int fid, size_t n_rows, size_t n_cols);
cudaError_t hash_columns(int fid, const uint64_t *d_mat, size_t n_rows, size_t row_stride, size_t n_cols,
                         const uint64_t *d_col_idx, uint8_t *d_leaves, uint8_t *d_cv_scratch, const Launch &lc);

// Pieces of hash_columns for streaming commits and row edits.  The leaf of a column is the BLAKE3 tree over
// its 1 KiB chunks; chunk chaining values live in d_cvs as [chunk][column][32 bytes].  hash_chunk_range
// computes chunks [chunk0, chunk_end) from a row window of the matrix (d_mat's first row is global row
// row_base; global rows >= n_rows_valid read as zero); total_bytes / n_chunks_total describe the whole
// leaf and matter only for the final chunk (length, ROOT flag when it is the only one) -- pass
// UINT64_MAX for both while more rows are still to come.  hash_merge folds the chaining values.
uint64_t hash_leaf_bytes(int fid, size_t n_rows);
uint64_t hash_leaf_chunks(int fid, size_t n_rows);
cudaError_t hash_chunk_range(int fid, const uint64_t *d_mat, int64_t row_base, size_t n_rows_valid, size_t row_stride,
                             size_t n_cols, uint64_t chunk0, uint64_t chunk_end, uint64_t total_bytes,
                             uint64_t n_chunks_total, uint8_t *d_cvs, const Launch &lc);
cudaError_t hash_merge(const uint8_t *d_cvs, size_t n_cols, uint64_t n_chunks, uint8_t *d_leaves, const Launch &lc);
// hash_chunk_range with the exchange of a row-sharded commit fused in: every chaining value is stored into the store of
// the rank that owns its column (base[g] = rank g's [n_chunks_total][1 << log_cb][32 B], local or peer-mapped memory)
struct CvScatter {
    uint32_t *base[16];
    int log_cb;
};
cudaError_t hash_chunk_range_scatter(int fid, const uint64_t *d_mat, int64_t row_base, size_t n_rows_valid, size_t row_stride,
                                     size_t n_cols, uint64_t chunk0, uint64_t chunk_end, uint64_t total_bytes,
                                     uint64_t n_chunks_total, const CvScatter &sc, const Launch &lc);

// d_out[c * out_col_stride + r] = canonical repr of d_mat[r][c] (proof-of-storage's column-major encoded file)
cudaError_t emit_colmajor(int fid, const uint64_t *d_mat, size_t n_rows, size_t row_stride, size_t n_cols, uint64_t *d_out,
                          size_t out_col_stride, const Launch &lc);

cudaError_t merkle_tree(uint8_t *d_hashes, size_t n_leaves, const Launch &lc);

// One-launch tails.  merge_tree: chunk chaining values (n_chunks >= 2; with n_chunks == 1 the leaves are already in
// d_hashes) -> leaves and the whole tree over the padded leaf range (padding leaves written as zero).  d_ticket: one
// zeroed counter, left zeroed.  hash_tree: the whole of merkleize (leaf hashing + merge + tree) for columns [0, n_cols)
// of a matrix; d_cvs needs hash_scratch_bytes(), d_tickets hash_tree_tickets(np2) zeroed counters (left zeroed).
// cv_stride: columns per chunk row of d_cvs (0 = n_cols; the column-block width for a row-sharded store)
cudaError_t merge_tree(const uint8_t *d_cvs, size_t n_cols, uint64_t n_chunks, uint8_t *d_hashes, size_t np2, unsigned *d_ticket,
                       const Launch &lc, size_t cv_stride = 0);
size_t hash_tree_tickets(size_t np2);
bool hash_tree_supported(int fid, size_t n_rows, size_t np2);
bool hash_tree_preferred(int fid, size_t n_rows, size_t np2);
cudaError_t hash_tree(int fid, const uint64_t *d_mat, size_t n_rows, size_t row_stride, size_t n_cols, size_t np2,
                      uint8_t *d_hashes, uint8_t *d_cvs, unsigned *d_tickets, const Launch &lc);

// out[t][j] = sum_r tensors[t][r] * mat[r][j].  d_scratch needs fold_scratch_bytes().
size_t fold_scratch_bytes(int fid, size_t n_rows, size_t width, size_t n_tensors);
cudaError_t fold(int fid, const uint64_t *d_mat, size_t n_rows, size_t width, size_t row_stride,
                 const uint64_t *d_tensors, size_t n_tensors, uint64_t *d_out, uint64_t *d_scratch,
                 const Launch &lc);
cudaError_t add_partials(int fid, const uint64_t *d_parts, size_t n_parts, size_t n, uint64_t *d_out,
                         const Launch &lc);

cudaError_t gather_columns(int fid, const uint64_t *d_mat, size_t n_rows, size_t row_stride,
                           const uint64_t *d_cols, size_t n, uint64_t *d_out, const Launch &lc);
// paths[i][l] = level_l[(cols[i] >> l) ^ 1], depth entries per column
cudaError_t gather_paths(const uint8_t *d_hashes, size_t np2, const uint64_t *d_cols, size_t n, uint8_t *d_paths,
                         const Launch &lc);

// verifier / transcript helpers
cudaError_t to_canon(int fid, const uint64_t *d_in, size_t n, uint64_t *d_out, const Launch &lc);
cudaError_t column_dots(int fid, const uint64_t *d_cols, size_t n_rows, size_t n_open, const uint64_t *d_tensors,
                        size_t n_tensors, uint64_t *d_out, const Launch &lc);
cudaError_t verify_paths(const uint8_t *d_leaves, const uint8_t *d_paths, int depth, const uint64_t *d_cols, size_t n,
                         const uint8_t *d_root, uint32_t *d_ok, const Launch &lc);

cudaError_t pack_bytes7(const uint8_t *d_bytes, size_t n_bytes, uint64_t *d_elems, const Launch &lc);
// Ft253_192::from_data_bytes; *d_bad (zeroed by the caller) becomes non-zero when a 31-byte group is not below the modulus
cudaError_t pack_bytes31(const uint8_t *d_bytes, size_t n_bytes, uint64_t *d_elems, uint32_t *d_bad, const Launch &lc);

// Brakedown: matrices in CSR on the device (converted from the caller's CSC once).
struct DevCsr {
    size_t rows = 0, cols = 0, nnz = 0;
    uint32_t *d_rowptr = nullptr;  // rows + 1
    uint32_t *d_colidx = nullptr;  // nnz
    uint64_t *d_data = nullptr;    // nnz * LIMBS
    uint32_t *d_order = nullptr;   // rows: row indices by decreasing length (the lane groups of a warp get equally long rows)
};
// one-time preparation of a matrix's non-zeros for the spmv kernel (multiplies by 2^32: the lazy dot products reduce by one extra word)
cudaError_t scale_csr_data(int fid, uint64_t *d_data, size_t nnz, cudaStream_t s);
struct SdigPlan {
    int fid = 0;
    size_t n_per_row = 0, n_cols = 0;
    std::vector<DevCsr> pre, post;
};
// Message rows at d_msg (stride msg_ld) -> codeword rows at d_comm (stride n_cols), every entry written.  In place
// (d_msg == d_comm, msg_ld == n_cols: the rows already hold the message in their first n_per_row entries) or from the
// coefficient matrix (msg_ld = n_per_row), in which case the pass that builds the transposed working copy also copies
// the message into d_comm.
// sc != nullptr (multi-GPU column blocks): d_comm is not used, both passes that would write it store every element into
// the column-block matrix of the rank that owns its column (ScatterDst: the PADDED column range is what is split).
cudaError_t sdig_encode(const SdigPlan &plan, const uint64_t *d_msg, size_t msg_ld, uint64_t *d_comm, size_t n_rows,
                        uint64_t *d_tmp, const Launch &lc, const ScatterDst *sc = nullptr);
size_t sdig_tmp_elems(const SdigPlan &plan, size_t n_rows);

}  // namespace lcpc
