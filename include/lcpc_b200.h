/*
 * lcpc_b200 -- C ABI of the B200-native lcpc commitment hot path.
 *
 * This is the drop-in boundary: the reference (TrevorGKann/lcpc_proof_of_storage) is
 * pure Rust and has no FFI of its own, so each entry point below names the Rust item
 * it stands in for (paths relative to the reference checkout).  INTEGRATION.md shows
 * the `extern "C"` block and the `LcEncoding` wrapper a maintainer adds on the Rust
 * side.
 *
 * Conventions
 *  - A field element is LIMBS x uint64_t, least-significant limb first, holding the
 *    Montgomery residue a*2^(64*LIMBS) mod p, fully reduced -- bit-identical to the
 *    ff_derive newtypes (lcpc-test-fields/src/lib.rs:18-70), so a Rust `&[F]` is
 *    passed as `(const uint64_t*)slice.as_ptr()` with `len * LIMBS` words.
 *  - Matrices are row-major: coeffs[r*n_per_row + j], comm[r*n_cols + j]
 *    (lcpc-2d/src/lib.rs:174-191).
 *  - Digests are 32-byte BLAKE3 outputs; `hashes` is the flat tree
 *    [np2 leaves | np2/2 | ... | root], np2 = next_power_of_two(n_cols), 2*np2-1
 *    entries, padding leaves all-zero (lib.rs:685-695, 720-734).
 *  - Every function returns an lcpc_status (0 = ok).  No C++ exception crosses the
 *    boundary.  Handles are internally locked: calls on one handle may come from any
 *    thread (`LcEncoding: Sync`, lib.rs:75).  Calls are synchronous on return.
 *  - "host" entry points take host pointers (pageable or pinned) and move the data
 *    themselves; "dev" entry points take device pointers on the context's device and
 *    enqueue on the context's stream (used for device-resident pipelines and for the
 *    one-process-per-GPU sharded path).
 *  - There is no CPU fallback: every entry point fails with LCPC_ERR_CUDA when no
 *    usable device is present.
 */
#ifndef LCPC_B200_H
#define LCPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCPC_DIGEST_BYTES 32

/* prime fields of the reference (lcpc-test-fields/src/lib.rs; WriteableFt63 =
 * proof-of-storage/src/fields/writable_ft63.rs has Ft63's modulus) */
typedef enum {
    LCPC_FT63 = 0,  /* p = 5102708120182849537, 1 limb  */
    LCPC_FT127 = 1, /* 2 limbs */
    LCPC_FT191 = 2, /* 3 limbs */
    LCPC_FT255 = 3, /* 4 limbs */
    /* proof-of-storage/src/fields/ft253_192.rs:6-10: p = (2^61 - 1) * 2^192 + 1, generator 3, 4 limbs, and the one
     * field whose to_repr() is BIG-endian: leaves, transcript messages and the encoded-file image use those bytes.
     * Elements still cross this interface as Montgomery limbs, least-significant limb first, fully reduced. */
    LCPC_FT253_192 = 4
} lcpc_field;

/* status codes; the first block maps onto lcpc_2d::ProverError (lib.rs:113-132) */
typedef enum {
    LCPC_OK = 0,
    LCPC_ERR_TOO_BIG = -1,       /* ProverError::TooBig                           */
    LCPC_ERR_ENCODE = -2,        /* ProverError::Encode(E::Err)                   */
    LCPC_ERR_COMMIT = -3,        /* ProverError::Commit (check_comm, :703-718)    */
    LCPC_ERR_COLUMN_NUMBER = -4, /* ProverError::ColumnNumber (:829-831)          */
    LCPC_ERR_OUTER_TENSOR = -5,  /* ProverError::OuterTensor (:1046-1048)         */
    LCPC_ERR_DIMS = -6,          /* the assert!s at lib.rs:659-661 / dims_ok      */
    LCPC_ERR_INVALID_ARG = -7,   /* null pointer, unknown field, ...              */
    LCPC_ERR_CUDA = -8,          /* CUDA runtime failure or no device             */
    LCPC_ERR_NOMEM = -9,         /* device or host allocation failed              */
    /* lcpc_2d::VerifierError (lib.rs:139-167), returned by lcpc_verify */
    LCPC_VERR_NUM_COL_OPENS = -20, /* VerifierError::NumColOpens  */
    LCPC_VERR_COLUMN_PATH = -21,   /* VerifierError::ColumnPath   */
    LCPC_VERR_COLUMN_EVAL = -22,   /* VerifierError::ColumnEval   */
    LCPC_VERR_COLUMN_DEGREE = -23, /* VerifierError::ColumnDegree */
    LCPC_VERR_OUTER_TENSOR = -24,  /* VerifierError::OuterTensor  */
    LCPC_VERR_INNER_TENSOR = -25,  /* VerifierError::InnerTensor  */
    LCPC_VERR_ENCODING_DIMS = -26, /* VerifierError::EncodingDims */
    LCPC_VERR_ENCODE = -27         /* VerifierError::Encode(E::Err) */
} lcpc_status;

typedef struct lcpc_ctx lcpc_ctx;       /* one device + stream                      */
typedef struct lcpc_plan lcpc_plan;     /* one encoding instance (an `E: LcEncoding`) */
typedef struct lcpc_commit lcpc_commit; /* a device-resident LcCommit               */
typedef struct lcpc_transcript lcpc_transcript; /* a merlin::Transcript (host)         */

/* sprs::CsMat<F> in CSC storage, as lcpc-brakedown-pc keeps its pre/postcodes
 * (lcpc-brakedown-pc/src/matgen.rs:187): indptr[cols+1], indices[nnz] = row numbers,
 * data[nnz*LIMBS] Montgomery limbs.  All host pointers. */
typedef struct {
    uint64_t rows;
    uint64_t cols;
    const uint64_t *indptr;
    const uint64_t *indices;
    const uint64_t *data;
} lcpc_csc;

/* ---- library / context --------------------------------------------------------- */

/* ABI version of this header (bumped on any signature change). */
uint32_t lcpc_abi_version(void);
/* Message for the last failure on this thread (never NULL). */
const char *lcpc_last_error(void);
/* LIMBS of a field, or 0 for an unknown id. */
int32_t lcpc_field_limbs(int32_t field);
/* Field constants (modulus, R mod p = F::ONE, 2-adicity S, ROOT_OF_UNITY in Montgomery
 * form); any out pointer may be NULL. */
int32_t lcpc_field_constants(int32_t field, uint64_t *modulus, uint64_t *one_mont,
                             uint64_t *root_of_unity_mont, int32_t *two_adicity, int32_t *num_bits);

/* Context on `device` with its own non-blocking stream. */
int32_t lcpc_ctx_create(int32_t device, lcpc_ctx **out);
/* Context that enqueues on a caller-owned cudaStream_t (e.g. torch's current stream). */
int32_t lcpc_ctx_create_on_stream(int32_t device, void *cuda_stream, lcpc_ctx **out);
/* Context over several devices of this process (the contract's `lcpc_ctx_create(devices, n)`):
 * n_devices a power of two <= 16, peer access between every pair is enabled inside the library.  Plans made on such a
 * context are built on every device, and lcpc_commit_host / lcpc_commit_bytes_host through them shard the commitment:
 * rows over the devices for encoding (row blocks cut on BLAKE3 chunk boundaries of the column leaves), every device hashes
 * the chunks of all columns over its own rows and stores the 32-byte chaining values straight into the store of the device
 * that owns the column block (peer HBM over NVLink), each device builds its Merkle subtree and the first device joins
 * the subtree roots.  lcpc_commit_root / _download / lcpc_fold_host / lcpc_open_columns_host / lcpc_leaves_host /
 * lcpc_prove work on the resulting handle unchanged, with bit-identical results.  Commitments too small to give every
 * device a BLAKE3 chunk of rows (and 24-byte elements, which straddle chunks) run on the first device, as do the
 * single-row and verifier entry points (lcpc_encode_rows, lcpc_decode_rows, lcpc_verify, lcpc_stream_*).  Row edits
 * (lcpc_commit_update_rows_host / _append_rows_host) and lcpc_commit_device_ptrs are refused on a sharded handle.
 * A device may be listed more than once (its shards then share it). */
int32_t lcpc_ctx_create_multi(const int32_t *devices, int32_t n_devices, lcpc_ctx **out);
/* Devices of a context: 1 for lcpc_ctx_create / _on_stream, n_devices for lcpc_ctx_create_multi. */
int32_t lcpc_ctx_device_count(const lcpc_ctx *ctx);
int32_t lcpc_ctx_synchronize(lcpc_ctx *ctx);
/* The cudaStream_t every call on this context enqueues on (borrowed).  A caller that mixes the device-pointer entry
 * points with its own work on the same buffers must either enqueue that work on this stream or order the two streams
 * with events: nothing else orders them. */
int32_t lcpc_ctx_stream(const lcpc_ctx *ctx, void **cuda_stream_out);
void lcpc_ctx_destroy(lcpc_ctx *ctx);

/* ---- plans = encodings ------------------------------------------------------------ */

/* LigeroEncodingRho::new_from_dims(n_per_row, n_cols) (lcpc-ligero-pc/src/lib.rs:138-148):
 * requires n_per_row < n_cols, n_cols a power of two <= 2^S (_dims_ok, :114-118).
 * `root_of_unity_mont` is the n_cols-th root the NTT uses, LIMBS words; NULL selects
 * ROOT_OF_UNITY^(2^(S-k)), what fffft::precomp_fft(n_cols) derives (:140).  The encode
 * is fffft's fft_io: in-order input, bit-reversed output (:162-164). */
int32_t lcpc_plan_ligero(lcpc_ctx *ctx, int32_t field, size_t n_per_row, size_t n_cols,
                         const uint64_t *root_of_unity_mont, lcpc_plan **out);

/* SdigEncodingS from already generated matrices (lcpc-brakedown-pc/src/lib.rs:126-137,
 * matgen.rs:28-53): n_levels precodes and postcodes; n_cols must equal
 * codeword_length(pre, post) (encode.rs:18-33). */
int32_t lcpc_plan_brakedown(lcpc_ctx *ctx, int32_t field, size_t n_per_row, size_t n_cols,
                            size_t n_levels, const lcpc_csc *precodes, const lcpc_csc *postcodes,
                            lcpc_plan **out);

/* LcEncoding::get_dims (lib.rs:166-169 / brakedown lib.rs:155-158) */
int32_t lcpc_plan_get_dims(const lcpc_plan *plan, size_t len, size_t *n_rows, size_t *n_per_row,
                           size_t *n_cols);
void lcpc_plan_destroy(lcpc_plan *plan);

/* LcEncoding::encode on `n_rows` rows at once, in place, host memory: each row has
 * n_cols elements, the first n_per_row are the message and the rest must be zero on
 * entry (lcpc-2d/src/lib.rs:677-682; verifier use at :912-918, :944-950). */
int32_t lcpc_encode_rows(lcpc_plan *plan, uint64_t *rows, size_t n_rows);

/* decode_row (proof-of-storage/src/lcpc_online.rs:568-573) = fffft::FieldFFT::ifft_oi on `n_rows` rows at once, in
 * place, host memory: every row holds n_cols elements of a Ligero codeword in the encoder's (bit-reversed) output order
 * and is replaced by the n_cols coefficients in natural order, scaled by 1/n_cols, so that
 * decode(encode(x)) == x (lcpc_online.rs:588-601, lcpc-2d/src/tests.rs:223-233).  Used by the server's append
 * (lcpc_online/file_handler.rs:370) and the encoded-file reader (encoded_file_reader.rs:66,83), which hold only encoded
 * rows.  Ligero plans only: LCPC_ERR_ENCODE for a Brakedown plan (the reference has no decoder for it either). */
int32_t lcpc_decode_rows(lcpc_plan *plan, uint64_t *rows, size_t n_rows);

/* ---- commit ----------------------------------------------------------------------- */

/* LcCommit::commit (lcpc-2d/src/lib.rs:314 -> commit :651-700): pad, encode every row,
 * hash every column, build the Merkle tree.
 *   coeffs      n_coeffs elements (host)
 *   coeffs_out  nullable; n_rows*n_per_row elements  = LcCommit.coeffs
 *   comm_out    nullable; n_rows*n_cols elements     = LcCommit.comm
 *   hashes_out  nullable; (2*np2-1)*32 bytes         = LcCommit.hashes
 *   keep        nullable; receives a device-resident handle for fold/open calls
 * Errors: LCPC_ERR_DIMS when n_coeffs == 0; LCPC_ERR_TOO_BIG when next_power_of_two
 * overflows (:685-687). */
int32_t lcpc_commit_host(lcpc_plan *plan, const uint64_t *coeffs, size_t n_coeffs,
                         uint64_t *coeffs_out, uint64_t *comm_out, uint8_t *hashes_out,
                         lcpc_commit **keep);

/* proof-of-storage path: DataField::from_byte_vec for WriteableFt63 (7 file bytes ->
 * one element whose limb IS the little-endian integer, proof-of-storage/src/fields/
 * writable_ft63.rs:35-40, data_field.rs:38-46) fused in front of the commit
 * (lcpc_online.rs:81-143 convert_file_data_to_commit, CommitRequestType::Commit).
 * Plan field LCPC_FT63, or LCPC_FT253_192 with Ft253_192::from_data_bytes (ft253_192.rs:18-30): 31 file bytes per
 * element, limb i = the big-endian integer of bytes [8i, 8i+8) of the zero-padded group.  A group whose limbs are not
 * below the modulus (byte 24 of the group > 0x1f) is refused with LCPC_ERR_INVALID_ARG: the reference computes on
 * such unreduced limbs with a carry-dropping add, and the outcome is not a field computation that could be matched. */
int32_t lcpc_commit_bytes_host(lcpc_plan *plan, const uint8_t *file_bytes, size_t n_bytes,
                               uint64_t *coeffs_out, uint64_t *comm_out, uint8_t *hashes_out,
                               lcpc_commit **keep);

/* Same as lcpc_commit_host with the coefficients already on the device; nothing is
 * copied to the host.  The handle owns coeffs (padded), comm and hashes in HBM. */
int32_t lcpc_commit_dev(lcpc_plan *plan, const uint64_t *d_coeffs, size_t n_coeffs, lcpc_commit **keep);

int32_t lcpc_commit_get_dims(const lcpc_commit *c, size_t *n_rows, size_t *n_per_row, size_t *n_cols);
/* LcCommit::get_root (lib.rs:291-296) */
int32_t lcpc_commit_root(lcpc_commit *c, uint8_t root_out[LCPC_DIGEST_BYTES]);
/* Copies of the handle's buffers to host memory; any pointer may be NULL. */
int32_t lcpc_commit_download(lcpc_commit *c, uint64_t *coeffs_out, uint64_t *comm_out, uint8_t *hashes_out);
/* Raw device pointers of the handle's buffers (borrowed; valid until lcpc_commit_free). */
int32_t lcpc_commit_device_ptrs(lcpc_commit *c, uint64_t **d_coeffs, uint64_t **d_comm, uint8_t **d_hashes);
void lcpc_commit_free(lcpc_commit *c);

/* ---- fold / open ------------------------------------------------------------------ */

/* collapse_columns (lib.rs:1126-1154): out[t][j] = sum_r tensors[t][r] * M[r][j].
 *   which = 0: M = coeffs (n_per_row wide)  -- prove's p_random / p_eval (:1064-1092)
 *   which = 1: M = comm   (n_cols wide)     -- proof-of-storage
 *              verifiable_polynomial_evaluation (lcpc_online.rs:454-484), tests' eval_outer_fft
 * tensors: n_tensors * n_rows elements (host); out: n_tensors * width elements (host). */
int32_t lcpc_fold_host(lcpc_commit *c, int32_t which, const uint64_t *tensors, size_t n_tensors,
                       uint64_t *out);

/* open_column for `n` columns at once (lib.rs:818-855): cols_out[i] = the n_rows
 * elements of column cols[i]; paths_out[i] = log2(np2) sibling digests, leaf level
 * first.  LCPC_ERR_COLUMN_NUMBER if any index >= n_cols.  Either output may be NULL
 * (ColumnsWithoutPath, lcpc_online.rs:191-225). */
int32_t lcpc_open_columns_host(lcpc_commit *c, const uint64_t *cols, size_t n, uint64_t *cols_out,
                               uint8_t *paths_out);

/* Leaf digests of selected columns only (CommitRequestType::Leaves, lcpc_online.rs:144-190). */
int32_t lcpc_leaves_host(lcpc_commit *c, const uint64_t *cols, size_t n, uint8_t *leaves_out);

/* ---- prove / verify ---------------------------------------------------------------- */

/* merlin::Transcript (merlin 2.0: STROBE-128 over Keccak-f[1600]); host-side, no device
 * needed.  new = Transcript::new(label); the other two are append_message and
 * challenge_bytes (call sites lcpc-2d/src/lib.rs:49, 901, 934, 1057, 1104). */
int32_t lcpc_transcript_new(const uint8_t *label, size_t label_len, lcpc_transcript **out);
int32_t lcpc_transcript_clone(const lcpc_transcript *t, lcpc_transcript **out);
int32_t lcpc_transcript_append_message(lcpc_transcript *t, const uint8_t *label, size_t label_len,
                                       const uint8_t *msg, size_t msg_len);
int32_t lcpc_transcript_challenge_bytes(lcpc_transcript *t, const uint8_t *label, size_t label_len,
                                        uint8_t *dest, size_t dest_len);
void lcpc_transcript_free(lcpc_transcript *t);

/* Challenge expansion as prove/verify do it (host-side): n x F::random from
 * ChaCha20Rng::from_seed(key) (lib.rs:1058-1062), and n x Uniform(0, n_cols) column
 * indices, with replacement (lib.rs:1105-1110). */
int32_t lcpc_random_field_vec(int32_t field, const uint8_t key[32], uint64_t *out, size_t n);
int32_t lcpc_random_columns(const uint8_t key[32], uint64_t n_cols, uint64_t *out, size_t n);

/* LcCommit::prove (lib.rs:319 -> prove :1034-1123).  n_degree_tests / n_col_opens are the
 * encoding's get_n_degree_tests() / get_n_col_opens().  Outputs (host):
 *   p_eval_out    n_per_row elements                       = LcEvalProof.p_eval
 *   p_random_out  n_degree_tests * n_per_row elements      = LcEvalProof.p_random_vec
 *   col_idx_out   nullable; the n_col_opens sampled column numbers
 *   columns_out   n_col_opens * n_rows elements            = LcEvalProof.columns[i].col
 *   paths_out     n_col_opens * log2(np2) * 32 bytes       = LcEvalProof.columns[i].path
 * LCPC_ERR_OUTER_TENSOR when outer_len != n_rows. */
int32_t lcpc_prove(lcpc_commit *c, const uint64_t *outer_tensor, size_t outer_len, size_t n_degree_tests,
                   size_t n_col_opens, lcpc_transcript *tr, uint64_t *p_eval_out, uint64_t *p_random_out,
                   uint64_t *col_idx_out, uint64_t *columns_out, uint8_t *paths_out);

/* LcEvalProof::verify (lib.rs:547 -> verify :862-982).  The proof is passed flat:
 * p_eval (n_per_row), p_random (n_p_random vectors of n_per_row), columns (n_columns x
 * n_rows), paths (n_columns x path_len digests), proof_n_cols = LcEvalProof.n_cols.
 * On success writes sum_j inner[j] * p_eval[j] (LIMBS words) to result_out; otherwise
 * returns the LCPC_VERR_* code of the VerifierError variant the reference would raise. */
int32_t lcpc_verify(lcpc_plan *plan, const uint8_t root[LCPC_DIGEST_BYTES], const uint64_t *outer_tensor,
                    size_t outer_len, const uint64_t *inner_tensor, size_t inner_len, size_t proof_n_cols,
                    const uint64_t *p_eval, size_t n_per_row, const uint64_t *p_random, size_t n_p_random,
                    const uint64_t *columns, size_t n_rows, const uint8_t *paths, size_t path_len,
                    size_t n_columns, size_t n_col_opens, size_t n_degree_tests, lcpc_transcript *tr,
                    uint64_t *result_out);

/* ---- proof-of-storage helpers -------------------------------------------------------- */

/* get_column_indicies_from_random_seed (proof-of-storage/src/networking/client.rs:443-456):
 * ChaCha8Rng::seed_from_u64(seed), then IteratorRandom::choose_multiple of `amount` indices out
 * of 0..max_index (without replacement).  Host-side.  *n_out = min(amount, max_index). */
int32_t lcpc_pos_choose_columns(uint64_t seed, size_t amount, size_t max_index, uint64_t *out, size_t *n_out);

/* The client's retrievability check on received columns (lcpc_online.rs:275-281, 370-398,
 * 439-452): leaves_out[i] = hash_column_to_digest(column i) (nullable), and when `paths` is given,
 * ok_out[i] = verify_column_path(column i, col_idx[i], root) (lcpc-2d/src/lib.rs:985-1012).
 * columns: n x n_rows elements, column-contiguous as in LcColumn.col; paths: n x path_len digests. */
int32_t lcpc_verify_columns_host(lcpc_ctx *ctx, int32_t field, const uint64_t *columns, size_t n_rows,
                                 const uint8_t *paths, size_t path_len, const uint64_t *col_idx, size_t n,
                                 const uint8_t root[LCPC_DIGEST_BYTES], uint8_t *leaves_out, uint32_t *ok_out);

/* ---- Brakedown code generation (host-side) ------------------------------------------ */

/* matgen::get_dims (lcpc-brakedown-pc/src/matgen.rs:56-111) for SdigCode<code> (1..6,
 * codespec.rs:168-232): per level (n, m, d) = (columns, rows, non-zeros per column) of the
 * precode and the postcode, 3 words each. */
int32_t lcpc_sdig_get_dims(int32_t code, uint64_t n_per_row, int32_t field, uint64_t *pre_dims,
                           uint64_t *post_dims, int32_t max_levels, int32_t *n_levels);
/* matgen::generate for one level (matgen.rs:38-49, gen_code :114-188): fills CSC arrays
 * (indptr[n+1], indices[n*d], data[n*d*LIMBS]) for the level's precode and postcode. */
int32_t lcpc_sdig_gen_level(int32_t field, uint64_t seed, uint64_t level, const uint64_t pre_dim[3],
                            const uint64_t post_dim[3], uint64_t *pre_indptr, uint64_t *pre_indices,
                            uint64_t *pre_data, uint64_t *post_indptr, uint64_t *post_indices,
                            uint64_t *post_data);
/* SdigSpecification::dist (codespec.rs:40-44) */
double lcpc_sdig_dist(int32_t code);

/* ---- device-pointer building blocks (sharded / device-resident pipelines) --------- */

/* rows [0, n_rows) of the padded coefficient matrix -> encoded rows.  d_coeffs has
 * row stride n_per_row, d_comm row stride n_cols.  d_coeffs may alias nothing in d_comm. */
int32_t lcpc_dev_encode(lcpc_plan *plan, const uint64_t *d_coeffs, size_t n_rows, uint64_t *d_comm);
/* lcpc_decode_rows on device memory, in place: n_rows rows of n_cols elements, row stride n_cols. */
int32_t lcpc_dev_decode(lcpc_plan *plan, uint64_t *d_rows, size_t n_rows);
/* Fused encode + re-shard for the one-process-per-GPU path: like lcpc_dev_encode, but the encoded rows are stored
 * directly into the column-block matrix of the rank that hashes those columns: peer_blocks[g] is rank g's
 * [n_rows_total][np2/n_peers] row-major matrix (np2 = n_cols rounded up to a power of two: the PADDED column range is
 * what the ranks split; this rank's own buffer or a peer-mapped pointer reached over NVLink), rows row0 .. row0+n_rows
 * of it are written.  Ligero plans: the last pass of the transform stores whole row blocks; d_scratch: n_rows*n_cols
 * elements of local scratch for the leading passes (may be NULL when the transform is a single pass).  Brakedown plans:
 * the two transposing passes that would write the encoded matrix (message columns, computed columns) store element by
 * element into the owners' matrices; d_scratch is not used.  The caller synchronises the ranks before anyone reads its
 * matrix. */
int32_t lcpc_dev_encode_scatter(lcpc_plan *plan, const uint64_t *d_coeffs, size_t n_rows, uint64_t row0,
                                uint64_t *d_scratch, uint64_t *const *peer_blocks, size_t n_peers);
/* hash_columns (lib.rs:736-775) on a column window: leaves[j] for j in [0, n_cols) of a
 * matrix with `n_rows` rows whose row stride is `row_stride` elements, starting at d_mat. */
int32_t lcpc_dev_hash_columns(lcpc_ctx *ctx, int32_t field, const uint64_t *d_mat, size_t n_rows,
                              size_t row_stride, size_t n_cols, uint8_t *d_leaves);
/* Row-sharded form of hash_columns (lib.rs:736-775) for the one-process-per-GPU path: a leaf is BLAKE3 over
 * 32 + n_rows_total * w bytes, and BLAKE3 hashes every 1024-byte chunk of that stream independently before its parent
 * tree joins them.  A rank that owns the rows of whole chunks computes their chaining values for ALL columns from its
 * own rows, and only 32 bytes per (chunk, column) travel.  d_mat = encoded rows starting at global row `row_base`
 * (row stride `row_stride` elements); chunks [chunk0, chunk_end) of the n_rows_total-row leaf are written to
 * d_cvs[((chunk - chunk0) * n_cols + column) * 32].  The rows of those chunks must lie inside d_mat.  LCPC_ERR_DIMS when the leaf is a single chunk (its chaining value IS the leaf: use
 * lcpc_dev_hash_columns) or when an element straddles chunk boundaries (24-byte elements). */
int32_t lcpc_dev_hash_chunk_range(lcpc_ctx *ctx, int32_t field, const uint64_t *d_mat, uint64_t row_base,
                                  size_t n_rows_total, size_t row_stride, size_t n_cols, uint64_t chunk0,
                                  uint64_t chunk_end, uint8_t *d_cvs);
/* lcpc_dev_hash_chunk_range with the exchange fused into the kernel: the chaining value of (chunk, column) is stored
 * directly into the chaining-value store of the rank that owns the column.  peer_cvs[g] is rank g's store,
 * [n_chunks_total][cb][32 B], cb = next_power_of_two(n_cols) / n_peers (this rank's own buffer or a peer-mapped pointer
 * reached over NVLink): the PADDED leaf range is what is split, so with a non power-of-two n_cols (Brakedown) the last
 * blocks hold fewer real columns or none; n_peers a power of two <= 16.  The caller synchronises the ranks before anyone runs
 * lcpc_dev_hash_merge on its store. */
int32_t lcpc_dev_hash_chunk_range_scatter(lcpc_ctx *ctx, int32_t field, const uint64_t *d_mat, uint64_t row_base,
                                          size_t n_rows_total, size_t row_stride, size_t n_cols, uint64_t chunk0,
                                          uint64_t chunk_end, uint8_t *const *peer_cvs, size_t n_peers);
/* leaves[j] = BLAKE3 parent tree over d_cvs[(c * n_cols + j) * 32], c in [0, n_chunks), n_chunks >= 2: the second half
 * of hash_columns once every chunk's chaining value is in place. */
int32_t lcpc_dev_hash_merge(lcpc_ctx *ctx, const uint8_t *d_cvs, size_t n_cols, uint64_t n_chunks, uint8_t *d_leaves);
/* lcpc_dev_hash_merge and lcpc_dev_merkle_tree in ONE launch: d_hashes is the flat tree [n_leaves | n_leaves/2 | ... | 1]
 * (n_leaves a power of two >= n_cols); leaves [0, n_cols) come from the chaining values (n_chunks >= 2) or are already in
 * d_hashes (n_chunks == 1, d_cvs may be NULL), leaves [n_cols, n_leaves) are written as zero (lib.rs:685-695), then
 * every level above.  cv_stride: columns per chunk row of d_cvs (0 = n_cols; the column-block width when d_cvs is a store
 * filled by lcpc_dev_hash_chunk_range_scatter and the block holds fewer real columns than it is wide).  The last CTA to finish its tile builds the top levels, so no second launch exists. */
int32_t lcpc_dev_hash_merge_tree(lcpc_ctx *ctx, const uint8_t *d_cvs, size_t n_cols, uint64_t n_chunks, uint8_t *d_hashes,
                                 size_t n_leaves, size_t cv_stride);
/* merkleize (lib.rs:720-734) of a device matrix in one launch: leaf hashing of columns [0, n_cols) (row stride
 * `row_stride` elements), the BLAKE3 parent tree per leaf and the Merkle tree over next_power_of_two(n_cols) leaves
 * into d_hashes ((2*np2-1)*32 bytes). */
int32_t lcpc_dev_merkleize(lcpc_ctx *ctx, int32_t field, const uint64_t *d_mat, size_t n_rows, size_t row_stride,
                           size_t n_cols, uint8_t *d_hashes);
/* merkle_tree (lib.rs:777-815) in place over [n_leaves | n_leaves/2 | ... | 1]; n_leaves a power of two. */
int32_t lcpc_dev_merkle_tree(lcpc_ctx *ctx, uint8_t *d_hashes, size_t n_leaves);
/* collapse_columns on device buffers; d_out has n_tensors*width elements. */
int32_t lcpc_dev_fold(lcpc_ctx *ctx, int32_t field, const uint64_t *d_mat, size_t n_rows, size_t width,
                      size_t row_stride, const uint64_t *d_tensors, size_t n_tensors, uint64_t *d_out);
/* element-wise modular sum of `n_parts` partial fold results (the reduction after the
 * all-gather in the row-sharded fold): d_out[i] = sum_k d_parts[k*n + i]. */
int32_t lcpc_dev_add_partials(lcpc_ctx *ctx, int32_t field, const uint64_t *d_parts, size_t n_parts,
                              size_t n, uint64_t *d_out);
/* strided gather of columns: d_out[i*n_rows + r] = d_mat[r*row_stride + cols[i]] (d_cols on device). */
int32_t lcpc_dev_gather_columns(lcpc_ctx *ctx, int32_t field, const uint64_t *d_mat, size_t n_rows,
                                size_t row_stride, const uint64_t *d_cols, size_t n, uint64_t *d_out);
/* Merkle paths of open_column (lib.rs:841-851) from a flat tree over n_leaves (a power of two) on the device:
 * d_paths[(i * depth + l) * 32 ..] = level_l[(d_cols[i] >> l) ^ 1], depth = log2(n_leaves) digests per column, leaf
 * level first.  Every index must be below n_leaves. */
int32_t lcpc_dev_gather_paths(lcpc_ctx *ctx, const uint8_t *d_hashes, size_t n_leaves, const uint64_t *d_cols, size_t n,
                              uint8_t *d_paths);
/* 7-byte packing (WriteableFt63::from_data_bytes): n_elems = ceil(n_bytes/7) limbs written. */
int32_t lcpc_dev_pack_bytes7(lcpc_ctx *ctx, const uint8_t *d_bytes, size_t n_bytes, uint64_t *d_elems);

/* ---------------------------------------------------------------------------------------------
 * Streaming commit -- proof-of-storage's EncodedFileWriter + ColumnDigestAccumulator
 * (src/lcpc_online/encoded_file_writer.rs:33-508, column_digest_accumulator.rs:17-118,
 * row_generator_iter.rs): rows arrive in order, are encoded block by block, folded into the column
 * digests one BLAKE3 chunk at a time and (optionally) appended to the column-major encoded file.
 * Device memory is O(block_rows * n_cols); the file never has to fit in HBM.
 *
 *   max_rows           upper bound on the rows that will be pushed (sizes the chaining-value store:
 *                      32 bytes per column per KiB of column)
 *   block_rows         rows encoded per step; 0 picks about 256 MiB of encoded rows
 *   sink               nullable host image of the encoded file (e.g. an mmap of the .porenc file):
 *                      element (r, c) is written as its canonical to_repr bytes at
 *                      sink[(c*sink_row_capacity + r) * 8*LIMBS] (encoded_file_writer.rs:327-349,
 *                      fields/data_field.rs:62-70); sink_row_capacity = the writer's row_capacity
 * push_elems / push_bytes take whole rows (n_per_row elements, or 7*n_per_row file bytes for the 63-bit
 * field, data_field.rs:38-46); only the final push may end inside a row, which is zero padded
 * (LCPC_ERR_DIMS if a later push follows).  finish hashes the last chunk, merges the chunk chaining
 * values into the leaves and builds the tree; hashes_out receives (2*np2-1)*32 bytes = the .portree
 * image (merkle_tree.rs:61-86).  The result equals lcpc_commit_host / lcpc_commit_bytes_host on the
 * concatenated input (row_generator_iter.rs:286-364 tests this for the reference). */
typedef struct lcpc_stream lcpc_stream;
int32_t lcpc_stream_begin(lcpc_plan *plan, size_t max_rows, size_t block_rows, uint8_t *sink,
                          size_t sink_row_capacity, lcpc_stream **out);
int32_t lcpc_stream_push_elems_host(lcpc_stream *s, const uint64_t *elems, size_t n_elems);
int32_t lcpc_stream_push_bytes_host(lcpc_stream *s, const uint8_t *bytes, size_t n_bytes);
int32_t lcpc_stream_finish(lcpc_stream *s, uint8_t *hashes_out, size_t *n_rows_out);
void lcpc_stream_free(lcpc_stream *s);

/* Row edit on a device-resident commitment -- FileHandler::edit_bytes -> reencode_row ->
 * recalculate_merkle_tree (src/lcpc_online/file_handler.rs:279-402, 474-481).  Replaces coefficient rows
 * [row0, row0 + n_rows) (coeff_rows: n_rows * n_per_row elements, zero padded by the caller), re-encodes
 * exactly those rows, re-hashes only the BLAKE3 chunks of the column leaves that contain them, and
 * rebuilds the tree.  comm_rows_out (nullable) receives the n_rows re-encoded rows, hashes_out (nullable)
 * the whole tree.  LCPC_ERR_DIMS when the range leaves the committed rows. */
int32_t lcpc_commit_update_rows_host(lcpc_commit *c, size_t row0, size_t n_rows, const uint64_t *coeff_rows,
                                     uint64_t *comm_rows_out, uint8_t *hashes_out);
/* Append -- FileHandler::append_bytes (file_handler.rs:336-402): rows [row0, row0 + n_rows) are written with
 * row0 <= committed rows <= row0 + n_rows, i.e. the last (partially filled) row may be replaced and new rows follow.
 * The resident matrices grow, the new rows are encoded, and only the column-leaf chunks from the first written row
 * (or the former last chunk, whichever is earlier) onwards are re-hashed. */
int32_t lcpc_commit_append_rows_host(lcpc_commit *c, size_t row0, size_t n_rows, const uint64_t *coeff_rows,
                                     uint64_t *comm_rows_out, uint8_t *hashes_out);

/* number of kernels this library has launched on this context since creation */
uint64_t lcpc_ctx_launch_count(const lcpc_ctx *ctx);
/* Per-kernel device timing: when enabled, every launch on this context is bracketed by
 * CUDA events on the context's stream.  The report drains the records collected so far:
 * one line per kernel name, "<name> <launches> <total_ms>\n" (owned by the context,
 * valid until the next call). */
int32_t lcpc_ctx_kernel_timing(lcpc_ctx *ctx, int32_t enable);
const char *lcpc_ctx_kernel_timing_report(lcpc_ctx *ctx);

/* Measures the integer-pipe issue rates of this device (a few milliseconds of microbenchmark kernels on the context's
 * stream): SM sub-partition cycles per warp instruction, all resident warps issuing, for pure streams of
 * [0] IMAD.WIDE.U32, [1] IMAD, [2] LOP3 (the ALU pipe) and [3] an equal mix of IMAD and LOP3 (two pipes at once);
 * sm_ghz_out (nullable) receives the SM clock observed during the run.  These are the denominators of the integer
 * roofline bench.py reports next to the HBM one: the commit kernels are bound by these pipes. */
int32_t lcpc_ctx_measure_int_pipes(lcpc_ctx *ctx, double cycles_out[4], double *sm_ghz_out);

#ifdef __cplusplus
}
#endif
#endif /* LCPC_B200_H */
