/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the lcpc commitment hot path: encode rows ->
 * hash columns -> Merkle tree -> fold rows -> open / verify columns, for the
 * reference's fields, with the reference's own parallel decomposition (rows in
 * parallel for encoding; 32-column blocks for hashing, Merkle layers and folds).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (lcpc_proof_of_storage_b200)
 * never links or calls it.
 *
 * The reference itself is Rust-nightly with two path dependencies outside the
 * tree (Cargo.toml:16-17) and cannot be compiled in this image (no cargo/rustc),
 * so there is no oracle/_ref build.  Ligero Merkle roots depend on the recalled
 * fffft convention: PARITY UNPINNED for that one piece (see orc_tmpl.h).
 *
 * Paths cited below are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "orc_blake3.h"
#include "orc_field.h"
#include "orc_rand.h"

/* sprs::CsMat in CSC storage: indptr[cols+1], indices[nnz] (row numbers), data[nnz*LIMBS] */
typedef struct {
    uint64_t rows, cols;
    const uint64_t *indptr;
    const uint64_t *indices;
    const uint64_t *data;
} orc_csc;

/* ---- field table: lcpc-test-fields/src/lib.rs:18-70 (moduli + generators) ---- */
static const orc_field FIELDS[ORC_N_FIELDS] = {
    {/* Ft63 = WriteableFt63, p = 5102708120182849537, generator 10 */
     1, 63, 41,
     {0x46d0760000000001ull},
     0x46d075ffffffffffull,
     {0x2b8e9dfffffffffdull},
     {0x13085abb0716119eull},
     {0x23bcb75f84213a43ull},
     0x7fffffffffffffffull, 0},
    {/* Ft127, generator 3 */
     2, 127, 40,
     {0x7f2bd90000000001ull, 0x6e754097ba20e0bfull},
     0x7f2bd8ffffffffffull,
     {0x01a84dfffffffffeull, 0x23157ed08bbe3e81ull},
     {0x816bd5407cf6dce5ull, 0x2c1637057de6fce8ull},
     {0xf491a1dff39975f8ull, 0x178fd41c0f6a04faull},
     0x7fffffffffffffffull, 0},
    {/* Ft191, generator 5 */
     3, 191, 41,
     {0xd246820000000001ull, 0x936888270ceecbcdull, 0x453708aa3fbc8ddaull},
     0xd24681ffffffffffull,
     {0x892c79fffffffffdull, 0x45c6678ad9339c96ull, 0x305ae60140ca5670ull},
     {0x6c25128031d873e2ull, 0xf71a3697a97ffdceull, 0x07ef71ae547daef9ull},
     {0xecd905456df2b092ull, 0x53ce189f0df0a05aull, 0x3f6e6da556ed31d9ull},
     0x7fffffffffffffffull, 0},
    {/* Ft255, generator 5 */
     4, 255, 41,
     {0x02a4f20000000001ull, 0xef73c79086595f30ull, 0xfda9df04b9575969ull, 0x663c799b6e4d2900ull},
     0x02a4f1ffffffffffull,
     {0xfab61bfffffffffeull, 0x211870def34d419full, 0x04ac41f68d514d2cull, 0x33870cc92365adfeull},
     {0xcf06aad260ab9990ull, 0x12f0d8856156a683ull, 0x5da77ded73588e21ull, 0x38725a1646845639ull},
     {0x9c745ae52a496067ull, 0x95ee9a4091329682ull, 0x854a3ee53365b80eull, 0x16edffae79969e76ull},
     0x7fffffffffffffffull, 0},
    {/* Ft253_192 (proof-of-storage/src/fields/ft253_192.rs:6-10): p = (2^61 - 1) * 2^192 + 1, generator 3,
        big-endian repr; REPR_SHAVE_BITS = 3 */
     4, 253, 192,
     {0x0000000000000001ull, 0, 0, 0x1fffffffffffffffull},
     0xffffffffffffffffull,
     {0xfffffffffffffff8ull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0x0000000000000007ull},
     {0xffffffffffff8040ull, 0xffffffffffffefffull, 0xfffffffffffffdffull, 0x0000000000007f7full},
     {0xe94731e93d73da14ull, 0x0e0f79fb69eec7bfull, 0x246cdb0e8f061ce3ull, 0x029543679ced2616ull},
     0x1fffffffffffffffull, 1},
};

const orc_field *orc_get_field(int fid) {
    return (fid >= 0 && fid < ORC_N_FIELDS) ? &FIELDS[fid] : NULL;
}

#define LIMBS 1
#define SUF(x) x##_L1
#include "orc_tmpl.h"
#undef LIMBS
#undef SUF
#define LIMBS 2
#define SUF(x) x##_L2
#include "orc_tmpl.h"
#undef LIMBS
#undef SUF
#define LIMBS 3
#define SUF(x) x##_L3
#include "orc_tmpl.h"
#undef LIMBS
#undef SUF
#define LIMBS 4
#define SUF(x) x##_L4
#include "orc_tmpl.h"
#undef LIMBS
#undef SUF

#define DISPATCH(F, call)              \
    switch ((F)->limbs) {              \
    case 1: call(_L1); break;          \
    case 2: call(_L2); break;          \
    case 3: call(_L3); break;          \
    default: call(_L4); break;         \
    }

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_get_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_field_limbs(int fid) { return FIELDS[fid].limbs; }

void orc_field_constants(int fid, uint64_t *p, uint64_t *inv, uint64_t *r, uint64_t *r2,
                         uint64_t *root, int *s, int *num_bits) {
    const orc_field *F = &FIELDS[fid];
    for (int i = 0; i < F->limbs; i++) { p[i] = F->p[i]; r[i] = F->r[i]; r2[i] = F->r2[i]; root[i] = F->root[i]; }
    *inv = F->inv; *s = F->s; *num_bits = F->num_bits;
}

/* element-wise vector ops, for pinning the arithmetic against Python integers */
void orc_fe_binop(int fid, int op, uint64_t *out, const uint64_t *a, const uint64_t *b, size_t n) {
    const orc_field *F = &FIELDS[fid];
    int L = F->limbs;
    for (size_t i = 0; i < n; i++) {
#define CALL(S)                                                            \
    if (op == 0) fe_add##S(F, out + i * L, a + i * L, b + i * L);          \
    else if (op == 1) fe_sub##S(F, out + i * L, a + i * L, b + i * L);     \
    else fe_mul##S(F, out + i * L, a + i * L, b + i * L)
        DISPATCH(F, CALL);
#undef CALL
    }
}

void orc_fe_to_canon(int fid, uint64_t *out, const uint64_t *a, size_t n) {
    const orc_field *F = &FIELDS[fid];
    int L = F->limbs;
    for (size_t i = 0; i < n; i++) {
#define CALL(S) fe_to_canon##S(F, out + i * L, a + i * L)
        DISPATCH(F, CALL);
#undef CALL
    }
}

/* PrimeField::to_repr() bytes of each element (8*LIMBS bytes per element) */
void orc_fe_to_repr(int fid, uint8_t *out, const uint64_t *a, size_t n) {
    const orc_field *F = &FIELDS[fid];
    int L = F->limbs;
    for (size_t i = 0; i < n; i++) {
#define CALL(S) fe_to_repr##S(F, out + i * L * 8, a + i * L)
        DISPATCH(F, CALL);
#undef CALL
    }
}

void orc_fe_from_canon(int fid, uint64_t *out, const uint64_t *a, size_t n) {
    const orc_field *F = &FIELDS[fid];
    int L = F->limbs;
    for (size_t i = 0; i < n; i++) {
#define CALL(S) fe_from_canon##S(F, out + i * L, a + i * L)
        DISPATCH(F, CALL);
#undef CALL
    }
}

void orc_fe_inv(int fid, uint64_t *out, const uint64_t *a, size_t n) {
    const orc_field *F = &FIELDS[fid];
    int L = F->limbs;
    for (size_t i = 0; i < n; i++) {
#define CALL(S) fe_inv##S(F, out + i * L, a + i * L)
        DISPATCH(F, CALL);
#undef CALL
    }
}

void orc_fe_pow(int fid, uint64_t *out, const uint64_t *a, uint64_t e) {
    const orc_field *F = &FIELDS[fid];
#define CALL(S) fe_pow##S(F, out, a, e)
    DISPATCH(F, CALL);
#undef CALL
}

/* w = ROOT_OF_UNITY^(2^(S - log_n)): the n-th root fffft::precomp_fft(n) uses
 * (lcpc-ligero-pc/src/lib.rs:140) */
void orc_ntt_root(int fid, int log_n, uint64_t *out) {
    const orc_field *F = &FIELDS[fid];
#define CALL(S) ntt_root##S(F, out, log_n)
    DISPATCH(F, CALL);
#undef CALL
}

/* fft_io on n_rows independent rows of 2^log_n elements, in place */
void orc_fft_io(int fid, uint64_t *x, int log_n, size_t n_rows) {
    const orc_field *F = &FIELDS[fid];
    size_t n = (size_t)1 << log_n;
    int L = F->limbs;
    if (log_n == 0) return;
    uint64_t *roots = (uint64_t *)malloc((n / 2 + 1) * L * sizeof(uint64_t));
#define CALL(S) ntt_roots##S(F, roots, log_n, 0)
    DISPATCH(F, CALL);
#undef CALL
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t r = 0; r < n_rows; r++) {
#define CALL(S) fft_io##S(F, x + r * n * L, log_n, roots)
        DISPATCH(F, CALL);
#undef CALL
    }
    free(roots);
}

void orc_ifft_oi(int fid, uint64_t *x, int log_n, size_t n_rows) {
    const orc_field *F = &FIELDS[fid];
    size_t n = (size_t)1 << log_n;
    int L = F->limbs;
    if (log_n == 0) return;
    uint64_t *roots = (uint64_t *)malloc((n / 2 + 1) * L * sizeof(uint64_t));
#define CALL(S) ntt_roots##S(F, roots, log_n, 1)
    DISPATCH(F, CALL);
#undef CALL
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t r = 0; r < n_rows; r++) {
#define CALL(S) ifft_oi##S(F, x + r * n * L, log_n, roots)
        DISPATCH(F, CALL);
#undef CALL
    }
    free(roots);
}

void orc_hash_columns(int fid, const uint64_t *comm, uint8_t *hashes, size_t n_rows, size_t n_cols,
                      size_t row_stride) {
    const orc_field *F = &FIELDS[fid];
#define CALL(S) hash_columns##S(F, comm, hashes, n_rows, n_cols, row_stride)
    DISPATCH(F, CALL);
#undef CALL
}

void orc_hash_column(int fid, const uint64_t *col, size_t n_rows, uint8_t *out) {
    const orc_field *F = &FIELDS[fid];
#define CALL(S) hash_column##S(F, col, n_rows, out)
    DISPATCH(F, CALL);
#undef CALL
}

/* lcpc-2d/src/lib.rs:777-815 merkle_tree / merkle_layer over the flat array
 * [np2 leaves | np2/2 | ... | 1]; parent = D(left || right) (:800-805). */
void orc_merkle_tree(uint8_t *hashes, size_t np2) {
    uint8_t *ins = hashes;
    size_t n_in = np2;
    while (n_in > 1) {
        uint8_t *outs = ins + n_in * 32;
        size_t n_out = n_in / 2;
        size_t n_blocks = (n_out + 15) / 16; /* 32 inputs per base-case block */
#pragma omp parallel for schedule(static)
        for (size_t b = 0; b < n_blocks; b++) {
            size_t lo = b * 16, hi = lo + 16 < n_out ? lo + 16 : n_out;
            for (size_t i = lo; i < hi; i++) orc_blake3(ins + 2 * i * 32, 64, outs + i * 32);
        }
        ins = outs;
        n_in = n_out;
    }
}

static size_t next_pow2(size_t v) {
    size_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

/* lib.rs:720-734 merkleize: leaves into hashes[..n_cols], padding leaves
 * n_cols..np2 stay all-zero (:685-695), then the tree. */
static void merkleize(int fid, const uint64_t *comm, uint8_t *hashes, size_t n_rows, size_t n_cols) {
    size_t np2 = next_pow2(n_cols);
    memset(hashes, 0, (2 * np2 - 1) * 32);
    orc_hash_columns(fid, comm, hashes, n_rows, n_cols, n_cols);
    orc_merkle_tree(hashes, np2);
}

void orc_merkleize(int fid, const uint64_t *comm, uint8_t *hashes, size_t n_rows, size_t n_cols) {
    merkleize(fid, comm, hashes, n_rows, n_cols);
}

/* lib.rs:651-700 commit with E = LigeroEncodingRho.  Returns 0, or -1 when the
 * dimension asserts at :659-661 would fire. */
int orc_commit_ligero(int fid, const uint64_t *coeffs_in, size_t len, size_t n_per_row, size_t n_cols,
                      uint64_t *coeffs, uint64_t *comm, uint8_t *hashes) {
    const orc_field *F = &FIELDS[fid];
    if (len == 0 || n_per_row == 0) return -1;
    size_t n_rows = (len + n_per_row - 1) / n_per_row;
    if (!(n_per_row < n_cols) || (n_cols & (n_cols - 1))) return -1;
#define CALL(S)                                                              \
    pad_coeffs##S(coeffs_in, len, coeffs, n_rows, n_per_row);                \
    encode_rows_ligero##S(F, coeffs, comm, n_rows, n_per_row, n_cols)
    DISPATCH(F, CALL);
#undef CALL
    merkleize(fid, comm, hashes, n_rows, n_cols);
    return 0;
}

/* the same with E = SdigEncodingS (lcpc-brakedown-pc/src/lib.rs:140-176) */
int orc_commit_sdig(int fid, const uint64_t *coeffs_in, size_t len, size_t n_per_row, size_t n_cols,
                    size_t n_levels, const orc_csc *pre, const orc_csc *post, uint64_t *coeffs,
                    uint64_t *comm, uint8_t *hashes) {
    const orc_field *F = &FIELDS[fid];
    if (len == 0 || n_per_row == 0 || !(n_per_row < n_cols)) return -1;
    size_t n_rows = (len + n_per_row - 1) / n_per_row;
#define CALL(S)                                                              \
    pad_coeffs##S(coeffs_in, len, coeffs, n_rows, n_per_row);                \
    encode_rows_sdig##S(F, coeffs, comm, n_rows, n_per_row, n_cols, n_levels, pre, post)
    DISPATCH(F, CALL);
#undef CALL
    merkleize(fid, comm, hashes, n_rows, n_cols);
    return 0;
}

void orc_sdig_encode_rows(int fid, uint64_t *rows, size_t n_rows, size_t n_cols, size_t n_levels,
                          const orc_csc *pre, const orc_csc *post) {
    const orc_field *F = &FIELDS[fid];
    int L = F->limbs;
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t r = 0; r < n_rows; r++) {
#define CALL(S) sdig_encode##S(F, rows + r * n_cols * L, n_levels, pre, post)
        DISPATCH(F, CALL);
#undef CALL
    }
}

void orc_collapse_columns(int fid, const uint64_t *coeffs, const uint64_t *tensor, uint64_t *poly,
                          size_t n_rows, size_t n_per_row) {
    const orc_field *F = &FIELDS[fid];
    memset(poly, 0, n_per_row * F->limbs * sizeof(uint64_t));
#define CALL(S) collapse_columns##S(F, coeffs, tensor, poly, n_rows, n_per_row)
    DISPATCH(F, CALL);
#undef CALL
}

static size_t log2_np2(size_t v) { /* lib.rs:857-859 */
    size_t l = 0;
    while (((size_t)1 << l) < v) l++;
    return l;
}

/* lib.rs:818-855 open_column; returns -1 for ProverError::ColumnNumber */
int orc_open_column(int fid, const uint64_t *comm, const uint8_t *hashes, size_t n_rows, size_t n_cols,
                    size_t column, uint64_t *col_out, uint8_t *path_out) {
    int L = FIELDS[fid].limbs;
    if (column >= n_cols) return -1;
    for (size_t r = 0; r < n_rows; r++)
        memcpy(col_out + r * L, comm + (r * n_cols + column) * L, L * sizeof(uint64_t));
    size_t path_len = log2_np2(n_cols), level_len = next_pow2(n_cols);
    const uint8_t *level = hashes;
    for (size_t l = 0; l < path_len; l++) {
        size_t other = column ^ 1;
        memcpy(path_out + l * 32, level + other * 32, 32);
        level += level_len * 32;
        level_len /= 2;
        column >>= 1;
    }
    return 0;
}

/* lib.rs:985-1012 verify_column_path */
int orc_verify_column_path(int fid, const uint64_t *col, size_t n_rows, const uint8_t *path,
                           size_t path_len, size_t col_num, const uint8_t *root) {
    uint8_t hash[32], buf[64];
    orc_hash_column(fid, col, n_rows, hash);
    for (size_t l = 0; l < path_len; l++) {
        if (col_num % 2 == 0) { memcpy(buf, hash, 32); memcpy(buf + 32, path + l * 32, 32); }
        else { memcpy(buf, path + l * 32, 32); memcpy(buf + 32, hash, 32); }
        orc_blake3(buf, 64, hash);
        col_num >>= 1;
    }
    return memcmp(hash, root, 32) == 0;
}

int orc_verify_column_value(int fid, const uint64_t *col, const uint64_t *tensor, size_t n_rows,
                            const uint64_t *poly_eval) {
    const orc_field *F = &FIELDS[fid];
    int ok = 0;
#define CALL(S) ok = verify_column_value##S(F, col, tensor, n_rows, poly_eval)
    DISPATCH(F, CALL);
#undef CALL
    return ok;
}

/* n x F::random from one ChaCha20Rng::from_seed(key): the degree-test tensors of
 * prove/verify (lib.rs:1056-1062, 898-907) */
void orc_random_field_vec(int fid, const uint8_t key[32], uint64_t *out, size_t n) {
    const orc_field *F = &FIELDS[fid];
    orc_chacha_rng rng;
    orc_chacha_from_seed(&rng, key);
    for (size_t i = 0; i < n; i++) {
#define CALL(S) fe_random##S(F, &rng, out + i * F->limbs)
        DISPATCH(F, CALL);
#undef CALL
    }
}

/* n_col_opens x Uniform(0, n_cols) from ChaCha20Rng::from_seed(key) (lib.rs:1103-1110) */
void orc_random_columns(const uint8_t key[32], uint64_t n_cols, uint64_t *out, size_t n) {
    orc_chacha_rng rng;
    orc_chacha_from_seed(&rng, key);
    for (size_t i = 0; i < n; i++) out[i] = orc_uniform_usize(&rng, n_cols);
}

/* ---- Brakedown code generation: lcpc-brakedown-pc/src/{codespec,matgen}.rs ---- */

typedef struct { uint64_t an, ad, bn, bd, rn, rd, blen; } sdig_spec;
/* codespec.rs:168-232 SdigCode1..6 */
static const sdig_spec SDIG_CODES[6] = {
    {239, 2000, 71, 2500, 71, 50, 20},   {69, 500, 111, 2500, 147, 100, 20},
    {89, 500, 61, 1000, 1521, 1000, 20}, {1, 5, 41, 500, 41, 25, 20},
    {211, 1000, 97, 1000, 202, 125, 20}, {119, 500, 241, 2000, 43, 25, 20}};

static double ent(double z) { /* codespec.rs:17-21 */
    double m = 1.0 - z;
    return -z * log2(z) - m * log2(m);
}

static uint64_t ceil_muldiv(uint64_t n, uint64_t num, uint64_t den) { return (n * num + den - 1) / den; }
static uint64_t umin(uint64_t a, uint64_t b) { return a < b ? a : b; }
static uint64_t umax(uint64_t a, uint64_t b) { return a > b ? a : b; }

/* codespec.rs:40-44: dist = beta/r */
double orc_sdig_dist(int code) {
    const sdig_spec *s = &SDIG_CODES[code - 1];
    return (double)(s->bn * s->rd) / (double)(s->bd * s->rn);
}

/* matgen.rs:56-111 get_dims.  Writes up to max_levels (ni, mi, d) triples per
 * array and returns the number of levels (or -1 if n <= baselen / overflow). */
int orc_sdig_get_dims(int code, uint64_t n, double log2p, uint64_t *pre_dims, uint64_t *post_dims,
                      int max_levels) {
    const sdig_spec *s = &SDIG_CODES[code - 1];
    double alpha = (double)s->an / (double)s->ad, beta = (double)s->bn / (double)s->bd;
    double r = (double)s->rn / (double)s->rd;
    double mu = r - 1.0 - r * alpha, nu = beta + alpha * beta + 0.03;
    double cn1 = ent(beta) + alpha * ent(1.28 * beta / alpha);
    double cn2 = beta * log2(alpha / (1.28 * beta));
    double dn1 = r * alpha * ent(beta / r) + mu * ent(nu / mu);
    double dn2 = alpha * beta * log2(mu / nu);
    if (n <= s->blen) return -1;
    uint64_t sizes[64];
    int k = 0;
    for (uint64_t ni = n; ni > s->blen; ni = ceil_muldiv(ni, s->an, s->ad)) {
        if (k >= 62) return -1;
        sizes[k++] = ni;
    }
    sizes[k] = ceil_muldiv(sizes[k - 1], s->an, s->ad);
    k++;
    int levels = k - 1;
    if (levels > max_levels) return -1;
    for (int i = 0; i < levels; i++) {
        uint64_t ni = sizes[i], mi = sizes[i + 1];
        uint64_t cn = umin(umax(ceil_muldiv(ni, 32 * s->bn, 25 * s->bd), 4 + ceil_muldiv(ni, s->bn, s->bd)),
                           (uint64_t)ceil((110.0 / (double)ni + cn1) / cn2));
        cn = umin(cn, mi);
        pre_dims[3 * i] = ni; pre_dims[3 * i + 1] = mi; pre_dims[3 * i + 2] = cn;
        uint64_t nip = ceil_muldiv(mi, s->rn, s->rd);
        uint64_t mip = ceil_muldiv(ni, s->rn, s->rd) - ni - nip;
        uint64_t t1 = ceil_muldiv(ni, 2 * s->bn, s->bd);
        uint64_t t2 = ceil_muldiv(ni, s->rn, s->rd) - ni + 110;
        uint64_t dn = umin(t1 + (uint64_t)ceil((double)t2 / log2p),
                           (uint64_t)ceil((110.0 / (double)ni + dn1) / dn2));
        dn = umin(dn, mip);
        post_dims[3 * i] = nip; post_dims[3 * i + 1] = mip; post_dims[3 * i + 2] = dn;
    }
    return levels;
}

/* matgen.rs:114-188 gen_code: n columns, each with d distinct sorted row indices
 * drawn Uniform(0, m) by rejection, then one non-zero F::random per index. */
static void gen_code(const orc_field *F, orc_chacha_rng *rng, uint64_t n, uint64_t m, uint64_t d,
                     uint64_t *indptr, uint64_t *indices, uint64_t *data) {
    int L = F->limbs;
    uint64_t nnz = 0;
    uint64_t *tmp = (uint64_t *)malloc((d + 1) * sizeof(uint64_t));
    indptr[0] = 0;
    for (uint64_t c = 0; c < n; c++) {
        uint64_t have = 0;
        while (have < d) {
            uint64_t x = orc_uniform_usize(rng, m);
            int dup = 0;
            for (uint64_t i = 0; i < have; i++) dup |= (tmp[i] == x);
            if (!dup) tmp[have++] = x;
        }
        for (uint64_t i = 1; i < d; i++) { /* sort_unstable: any sort, indices distinct */
            uint64_t v = tmp[i], j = i;
            while (j > 0 && tmp[j - 1] > v) { tmp[j] = tmp[j - 1]; j--; }
            tmp[j] = v;
        }
        for (uint64_t i = 0; i < d; i++) {
            uint64_t *val = data + nnz * L;
            for (;;) {
#define CALL(S) fe_random##S(F, rng, val)
                DISPATCH(F, CALL);
#undef CALL
                int z = 1;
                for (int l = 0; l < L; l++) z &= (val[l] == 0);
                if (!z) break;
            }
            indices[nnz++] = tmp[i];
        }
        indptr[c + 1] = nnz;
    }
    free(tmp);
}

/* matgen.rs:38-49: level i uses ChaCha20Rng::seed_from_u64(seed) with stream i,
 * precode first, then postcode from the same stream.  Array sizes: pre has
 * ni*cn entries, post has nip*dn. */
void orc_sdig_gen_level(int fid, uint64_t seed, uint64_t level, const uint64_t pre_dim[3],
                        const uint64_t post_dim[3], uint64_t *pre_indptr, uint64_t *pre_indices,
                        uint64_t *pre_data, uint64_t *post_indptr, uint64_t *post_indices,
                        uint64_t *post_data) {
    const orc_field *F = &FIELDS[fid];
    orc_chacha_rng rng;
    orc_chacha_seed_from_u64(&rng, seed);
    orc_chacha_set_stream(&rng, level);
    gen_code(F, &rng, pre_dim[0], pre_dim[1], pre_dim[2], pre_indptr, pre_indices, pre_data);
    gen_code(F, &rng, post_dim[0], post_dim[1], post_dim[2], post_indptr, post_indices, post_data);
}
