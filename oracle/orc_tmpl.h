/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Field-width-generic bodies, instantiated once per LIMBS by orc_lcpc.c
 * (`#define LIMBS n` + `#define SUF(x) x##_Ln` before each #include).
 * Every function cites the reference lines it restates; paths are relative to
 * /root/reference.
 */

typedef unsigned __int128 u128;

/* ---- field arithmetic (ff_derive 0.13 semantics: Montgomery, fully reduced) ---- */

static inline int SUF(fe_geq_p)(const orc_field *F, const uint64_t *a) {
    for (int i = LIMBS - 1; i >= 0; i--) {
        if (a[i] > F->p[i]) return 1;
        if (a[i] < F->p[i]) return 0;
    }
    return 1;
}

static inline void SUF(fe_sub_p)(const orc_field *F, uint64_t *a) {
    uint64_t borrow = 0;
    for (int i = 0; i < LIMBS; i++) {
        u128 d = (u128)a[i] - F->p[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
}

static inline void SUF(fe_add)(const orc_field *F, uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t t[LIMBS], carry = 0;
    for (int i = 0; i < LIMBS; i++) {
        u128 s = (u128)a[i] + b[i] + carry;
        t[i] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
    }
    if (carry || SUF(fe_geq_p)(F, t)) SUF(fe_sub_p)(F, t);
    for (int i = 0; i < LIMBS; i++) r[i] = t[i];
}

static inline void SUF(fe_sub)(const orc_field *F, uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t t[LIMBS], borrow = 0;
    for (int i = 0; i < LIMBS; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        t[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    if (borrow) {
        uint64_t carry = 0;
        for (int i = 0; i < LIMBS; i++) {
            u128 s = (u128)t[i] + F->p[i] + carry;
            t[i] = (uint64_t)s;
            carry = (uint64_t)(s >> 64);
        }
    }
    for (int i = 0; i < LIMBS; i++) r[i] = t[i];
}

/* Montgomery product a*b*R^-1 mod p, coarsely-integrated operand scanning */
static inline void SUF(fe_mul)(const orc_field *F, uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t t[LIMBS + 2];
    for (int i = 0; i < LIMBS + 2; i++) t[i] = 0;
    for (int i = 0; i < LIMBS; i++) {
        u128 c = 0;
        for (int j = 0; j < LIMBS; j++) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[LIMBS];
        t[LIMBS] = (uint64_t)c;
        t[LIMBS + 1] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->inv;
        c = (u128)m * F->p[0] + t[0];
        c >>= 64;
        for (int j = 1; j < LIMBS; j++) {
            c += (u128)m * F->p[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[LIMBS];
        t[LIMBS - 1] = (uint64_t)c;
        t[LIMBS] = t[LIMBS + 1] + (uint64_t)(c >> 64);
    }
    if (t[LIMBS] || SUF(fe_geq_p)(F, t)) SUF(fe_sub_p)(F, t);
    for (int i = 0; i < LIMBS; i++) r[i] = t[i];
}

static inline int SUF(fe_is_zero)(const uint64_t *a) {
    uint64_t o = 0;
    for (int i = 0; i < LIMBS; i++) o |= a[i];
    return o == 0;
}

static inline int SUF(fe_eq)(const uint64_t *a, const uint64_t *b) {
    uint64_t o = 0;
    for (int i = 0; i < LIMBS; i++) o |= a[i] ^ b[i];
    return o == 0;
}

/* canonical value (PrimeField::to_repr, little-endian limbs): one Montgomery reduction */
static inline void SUF(fe_to_canon)(const orc_field *F, uint64_t *r, const uint64_t *a) {
    uint64_t one[LIMBS];
    for (int i = 0; i < LIMBS; i++) one[i] = 0;
    one[0] = 1;
    SUF(fe_mul)(F, r, a, one);
}

/* PrimeField::to_repr(): the canonical value as 8*LIMBS bytes, little-endian for the lcpc-test-fields fields and
 * WriteableFt63, big-endian for Ft253_192 (ff_derive's PrimeFieldReprEndianness attribute).  Host is little-endian. */
static inline void SUF(fe_to_repr)(const orc_field *F, uint8_t *out, const uint64_t *a) {
    uint64_t canon[LIMBS];
    SUF(fe_to_canon)(F, canon, a);
    if (!F->repr_big_endian) {
        memcpy(out, canon, sizeof canon);
    } else {
        const uint8_t *src = (const uint8_t *)canon;
        for (int i = 0; i < 8 * LIMBS; i++) out[i] = src[8 * LIMBS - 1 - i];
    }
}

static inline void SUF(fe_from_canon)(const orc_field *F, uint64_t *r, const uint64_t *a) {
    SUF(fe_mul)(F, r, a, F->r2);
}

/* a^e, e a plain 64-bit exponent (Field::pow_vartime) */
static void SUF(fe_pow)(const orc_field *F, uint64_t *r, const uint64_t *a, uint64_t e) {
    uint64_t acc[LIMBS], base[LIMBS];
    for (int i = 0; i < LIMBS; i++) { acc[i] = F->r[i]; base[i] = a[i]; }
    while (e) {
        if (e & 1) SUF(fe_mul)(F, acc, acc, base);
        SUF(fe_mul)(F, base, base, base);
        e >>= 1;
    }
    for (int i = 0; i < LIMBS; i++) r[i] = acc[i];
}

/* a^(p-2) */
static void SUF(fe_inv)(const orc_field *F, uint64_t *r, const uint64_t *a) {
    uint64_t e[LIMBS], acc[LIMBS], base[LIMBS];
    for (int i = 0; i < LIMBS; i++) { e[i] = F->p[i]; acc[i] = F->r[i]; base[i] = a[i]; }
    for (int i = 0, borrow = 2; i < LIMBS && borrow; i++) { /* e = p - 2 (Ft253_192's low limb is 1: the borrow ripples) */
        uint64_t v = e[i];
        e[i] = v - (uint64_t)borrow;
        borrow = v < (uint64_t)borrow;
    }
    for (int i = 0; i < LIMBS; i++)
        for (int b = 0; b < 64; b++) {
            if ((e[i] >> b) & 1) SUF(fe_mul)(F, acc, acc, base);
            SUF(fe_mul)(F, base, base, base);
        }
    for (int i = 0; i < LIMBS; i++) r[i] = acc[i];
}

/* ---- NTT: fffft::FieldFFT::{fft_io, ifft_oi} ------------------------------------
 * fffft is a path dependency outside the tree (Cargo.toml:16); call sites:
 * lcpc-ligero-pc/src/lib.rs:140 (precomp_fft), :163 (fft_io_pc),
 * lcpc-2d/src/tests.rs:224 and proof-of-storage/src/lcpc_online.rs:572 (ifft_oi).
 * Restated from the published crate (kwantam/fffft): roots = [w^0 .. w^(n/2-1)]
 * with w = ROOT_OF_UNITY^(2^(S-k)); "io" = Gentleman-Sande decimation in
 * frequency, in-order input, bit-reversed output.  PARITY UNPINNED against the
 * Rust crate (no vector in the reference fixes root or permutation).          */

static void SUF(ntt_root)(const orc_field *F, uint64_t *w, int log_n) {
    for (int i = 0; i < LIMBS; i++) w[i] = F->root[i];
    for (int i = 0; i < F->s - log_n; i++) SUF(fe_mul)(F, w, w, w);
}

/* roots[i] = w^i, i < n/2 (n >= 2) */
static void SUF(ntt_roots)(const orc_field *F, uint64_t *roots, int log_n, int inverse) {
    uint64_t w[LIMBS];
    SUF(ntt_root)(F, w, log_n);
    if (inverse) SUF(fe_inv)(F, w, w);
    size_t half = ((size_t)1 << log_n) / 2;
    for (int i = 0; i < LIMBS; i++) roots[i] = F->r[i];
    for (size_t k = 1; k < half; k++)
        SUF(fe_mul)(F, roots + k * LIMBS, roots + (k - 1) * LIMBS, w);
}

static void SUF(fft_io)(const orc_field *F, uint64_t *x, int log_n, const uint64_t *roots) {
    size_t n = (size_t)1 << log_n;
    for (size_t gap = n / 2; gap > 0; gap /= 2) {
        size_t nchunks = n / (2 * gap);
        for (size_t c = 0; c < nchunks; c++) {
            uint64_t *lo = x + 2 * c * gap * LIMBS, *hi = lo + gap * LIMBS;
            for (size_t i = 0; i < gap; i++) {
                uint64_t neg[LIMBS];
                SUF(fe_sub)(F, neg, lo + i * LIMBS, hi + i * LIMBS);
                SUF(fe_add)(F, lo + i * LIMBS, lo + i * LIMBS, hi + i * LIMBS);
                SUF(fe_mul)(F, hi + i * LIMBS, neg, roots + nchunks * i * LIMBS);
            }
        }
    }
}

/* inverse: bit-reversed input, in-order output, scaled by 1/n; iroots = inverse roots */
static void SUF(ifft_oi)(const orc_field *F, uint64_t *x, int log_n, const uint64_t *iroots) {
    size_t n = (size_t)1 << log_n;
    for (size_t gap = 1; gap < n; gap *= 2) {
        size_t nchunks = n / (2 * gap);
        for (size_t c = 0; c < nchunks; c++) {
            uint64_t *lo = x + 2 * c * gap * LIMBS, *hi = lo + gap * LIMBS;
            for (size_t i = 0; i < gap; i++) {
                uint64_t t[LIMBS], neg[LIMBS];
                SUF(fe_mul)(F, t, hi + i * LIMBS, iroots + nchunks * i * LIMBS);
                SUF(fe_sub)(F, neg, lo + i * LIMBS, t);
                SUF(fe_add)(F, lo + i * LIMBS, lo + i * LIMBS, t);
                for (int l = 0; l < LIMBS; l++) hi[i * LIMBS + l] = neg[l];
            }
        }
    }
    uint64_t ninv[LIMBS], nn[LIMBS];
    for (int l = 0; l < LIMBS; l++) nn[l] = 0;
    nn[0] = (uint64_t)n;
    SUF(fe_from_canon)(F, nn, nn);
    SUF(fe_inv)(F, ninv, nn);
    for (size_t i = 0; i < n; i++) SUF(fe_mul)(F, x + i * LIMBS, x + i * LIMBS, ninv);
}

/* ---- commit: lcpc-2d/src/lib.rs:651-700 ------------------------------------------ */

/* lib.rs:665-674: zero-padded local copy of the coefficients */
static void SUF(pad_coeffs)(const uint64_t *in, size_t len, uint64_t *coeffs, size_t n_rows,
                            size_t n_per_row) {
    memset(coeffs, 0, n_rows * n_per_row * LIMBS * sizeof(uint64_t));
    memcpy(coeffs, in, len * LIMBS * sizeof(uint64_t));
}

/* lib.rs:677-682 with E = LigeroEncodingRho (lcpc-ligero-pc/src/lib.rs:162-164):
 * copy each padded row into the wider row and run fft_io on it; rayon over rows
 * becomes an OpenMP loop over rows. */
static void SUF(encode_rows_ligero)(const orc_field *F, const uint64_t *coeffs, uint64_t *comm,
                                    size_t n_rows, size_t n_per_row, size_t n_cols) {
    int log_n = 0;
    while (((size_t)1 << log_n) < n_cols) log_n++;
    uint64_t *roots = (uint64_t *)malloc((n_cols / 2 + 1) * LIMBS * sizeof(uint64_t));
    SUF(ntt_roots)(F, roots, log_n, 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t r = 0; r < n_rows; r++) {
        uint64_t *row = comm + r * n_cols * LIMBS;
        memset(row, 0, n_cols * LIMBS * sizeof(uint64_t));
        memcpy(row, coeffs + r * n_per_row * LIMBS, n_per_row * LIMBS * sizeof(uint64_t));
        SUF(fft_io)(F, row, log_n, roots);
    }
    free(roots);
}

/* lib.rs:736-775 hash_columns: leaf_j = D(0^32 || repr(M[0][j]) || ... ); the
 * recursion bottoms out at blocks of <= 32 columns (LOG_MIN_NCOLS = 5, :648),
 * processed row-major inside the block (:757-761). */
static void SUF(hash_columns)(const orc_field *F, const uint64_t *comm, uint8_t *hashes,
                              size_t n_rows, size_t n_cols, size_t row_stride) {
    size_t n_blocks = (n_cols + 31) / 32;
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t b = 0; b < n_blocks; b++) {
        size_t c0 = b * 32, nc = n_cols - c0 < 32 ? n_cols - c0 : 32;
        orc_b3_hasher dig[32];
        uint8_t zeros[32] = {0};
        for (size_t c = 0; c < nc; c++) {
            orc_b3_init(&dig[c]);
            orc_b3_update(&dig[c], zeros, 32);
        }
        for (size_t r = 0; r < n_rows; r++)
            for (size_t c = 0; c < nc; c++) {
                uint8_t repr[8 * LIMBS];
                SUF(fe_to_repr)(F, repr, comm + (r * row_stride + c0 + c) * LIMBS);
                orc_b3_update(&dig[c], repr, sizeof repr);
            }
        for (size_t c = 0; c < nc; c++) orc_b3_finalize(&dig[c], hashes + (c0 + c) * 32);
    }
}

/* lib.rs:1126-1154 collapse_columns: poly[j] += coeffs[r*n_per_row + j] * tensor[r];
 * same 32-column blocking. `poly` is accumulated into (caller zeroes it). */
static void SUF(collapse_columns)(const orc_field *F, const uint64_t *coeffs, const uint64_t *tensor,
                                  uint64_t *poly, size_t n_rows, size_t n_per_row) {
    size_t n_blocks = (n_per_row + 31) / 32;
#pragma omp parallel for schedule(static)
    for (size_t b = 0; b < n_blocks; b++) {
        size_t c0 = b * 32, nc = n_per_row - c0 < 32 ? n_per_row - c0 : 32;
        for (size_t r = 0; r < n_rows; r++)
            for (size_t c = 0; c < nc; c++) {
                uint64_t t[LIMBS];
                SUF(fe_mul)(F, t, coeffs + (r * n_per_row + c0 + c) * LIMBS, tensor + r * LIMBS);
                SUF(fe_add)(F, poly + (c0 + c) * LIMBS, poly + (c0 + c) * LIMBS, t);
            }
    }
}

/* lib.rs:1015-1030 verify_column_value: sum_r tensor[r]*col[r] == poly_eval */
static int SUF(verify_column_value)(const orc_field *F, const uint64_t *col, const uint64_t *tensor,
                                    size_t n_rows, const uint64_t *poly_eval) {
    uint64_t acc[LIMBS], t[LIMBS];
    for (int i = 0; i < LIMBS; i++) acc[i] = 0;
    for (size_t r = 0; r < n_rows; r++) {
        SUF(fe_mul)(F, t, tensor + r * LIMBS, col + r * LIMBS);
        SUF(fe_add)(F, acc, acc, t);
    }
    return SUF(fe_eq)(acc, poly_eval);
}

/* leaf hash of one opened column: lib.rs:990-997 (verify_column_path, first half),
 * identical to proof-of-storage/src/lcpc_online.rs:439-452 hash_field_vec_to_digest */
static void SUF(hash_column)(const orc_field *F, const uint64_t *col, size_t n_rows, uint8_t out[32]) {
    orc_b3_hasher h;
    uint8_t zeros[32] = {0};
    orc_b3_init(&h);
    orc_b3_update(&h, zeros, 32);
    for (size_t r = 0; r < n_rows; r++) {
        uint8_t repr[8 * LIMBS];
        SUF(fe_to_repr)(F, repr, col + r * LIMBS);
        orc_b3_update(&h, repr, sizeof repr);
    }
    orc_b3_finalize(&h, out);
}

/* ---- Brakedown: lcpc-brakedown-pc/src/encode.rs ---------------------------------- */

/* y = A*x for A (rows x cols) in CSC; sprs CsMat::dot on a dense vector
 * (encode.rs:52,66,85).  Exact field arithmetic: summation order is irrelevant. */
static void SUF(csc_matvec)(const orc_field *F, const orc_csc *A, const uint64_t *x, uint64_t *y) {
    for (size_t i = 0; i < A->rows * LIMBS; i++) y[i] = 0;
    for (size_t j = 0; j < A->cols; j++)
        for (uint64_t k = A->indptr[j]; k < A->indptr[j + 1]; k++) {
            uint64_t t[LIMBS];
            uint64_t *yi = y + A->indices[k] * LIMBS;
            SUF(fe_mul)(F, t, A->data + k * LIMBS, x + j * LIMBS);
            SUF(fe_add)(F, yi, yi, t);
        }
}

/* encode.rs:97-109 reed_solomon: xo[r] = sum_j xi[j] * (r+1)^j by Horner */
static void SUF(reed_solomon)(const orc_field *F, const uint64_t *xi, size_t n_in, uint64_t *xo,
                              size_t n_out) {
    uint64_t x[LIMBS];
    for (int i = 0; i < LIMBS; i++) x[i] = F->r[i];
    for (size_t r = 0; r < n_out; r++) {
        uint64_t acc[LIMBS];
        for (int i = 0; i < LIMBS; i++) acc[i] = 0;
        for (size_t j = n_in; j-- > 0;) {
            SUF(fe_mul)(F, acc, acc, x);
            SUF(fe_add)(F, acc, acc, xi + j * LIMBS);
        }
        for (int i = 0; i < LIMBS; i++) xo[r * LIMBS + i] = acc[i];
        SUF(fe_add)(F, x, x, F->r);
    }
}

/* encode.rs:36-94 encode, one row in place.  xi has codeword_length entries, the
 * first precodes[0].cols of them are the message. */
static void SUF(sdig_encode)(const orc_field *F, uint64_t *xi, size_t n_levels, const orc_csc *pre,
                             const orc_csc *post) {
    size_t in_start = 0;
    for (size_t l = 0; l + 1 < n_levels; l++) { /* :46-58 precodes all the way down */
        size_t in_end = in_start + pre[l].cols;
        SUF(csc_matvec)(F, &pre[l], xi + in_start * LIMBS, xi + in_end * LIMBS);
        in_start = in_end;
    }
    /* :61-74 base case: last precode into a temporary, Reed-Solomon of that */
    const orc_csc *lp = &pre[n_levels - 1];
    size_t in_end = in_start + lp->cols;
    uint64_t *tmp = (uint64_t *)malloc((lp->rows + 1) * LIMBS * sizeof(uint64_t));
    SUF(csc_matvec)(F, lp, xi + in_start * LIMBS, tmp);
    size_t out_end = in_end + post[n_levels - 1].cols;
    SUF(reed_solomon)(F, tmp, lp->rows, xi + in_end * LIMBS, out_end - in_end);
    free(tmp);
    in_start = in_end + lp->rows;
    size_t out_start = out_end;
    for (size_t l = n_levels; l-- > 0;) { /* :76-90 postcodes back up */
        in_start -= pre[l].rows;
        SUF(csc_matvec)(F, &post[l], xi + in_start * LIMBS, xi + out_start * LIMBS);
        out_start += post[l].rows;
    }
}

static void SUF(encode_rows_sdig)(const orc_field *F, const uint64_t *coeffs, uint64_t *comm,
                                  size_t n_rows, size_t n_per_row, size_t n_cols, size_t n_levels,
                                  const orc_csc *pre, const orc_csc *post) {
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t r = 0; r < n_rows; r++) {
        uint64_t *row = comm + r * n_cols * LIMBS;
        memset(row, 0, n_cols * LIMBS * sizeof(uint64_t));
        memcpy(row, coeffs + r * n_per_row * LIMBS, n_per_row * LIMBS * sizeof(uint64_t));
        SUF(sdig_encode)(F, row, n_levels, pre, post);
    }
}

/* ff_derive `Field::random`: draw LIMBS x next_u64 (limb 0 first), mask the unused
 * top bits, reject if >= p; the accepted limbs ARE the Montgomery residue.
 * Call sites: lcpc-2d/src/lib.rs:904,1060; lcpc-brakedown-pc/src/matgen.rs:175-177. */
static void SUF(fe_random)(const orc_field *F, orc_chacha_rng *rng, uint64_t *out) {
    for (;;) {
        for (int i = 0; i < LIMBS; i++) out[i] = orc_chacha_next_u64(rng);
        out[LIMBS - 1] &= F->top_mask;
        if (!SUF(fe_geq_p)(F, out)) return;
    }
}
