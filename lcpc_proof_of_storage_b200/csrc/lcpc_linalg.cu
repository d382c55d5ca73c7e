// Row folds, column gathers, byte packing and the Brakedown expander encoder.
//
//  - fold: collapse_columns (lcpc-2d/src/lib.rs:1126-1154) out[j] = sum_r t[r]*M[r][j];
//    one thread per column (coalesced row segments), rows split across CTAs to fill the
//    machine, several tensors per pass so the matrix is streamed once.
//  - gather_columns: the strided column read of open_column (lib.rs:833-839).
//  - pack_bytes7: WriteableFt63::from_data_bytes over a whole file
//    (proof-of-storage/src/fields/writable_ft63.rs:35-40, data_field.rs:38-46).
//  - sdig_encode: lcpc-brakedown-pc/src/encode.rs:36-109 on all rows of the matrix at
//    once, CSR sparse-matrix x dense-batch products level by level.
#include "lcpc_field.cuh"
#include "lcpc_kernels.h"

namespace lcpc {

#define LCPC_FIELD_SWITCH(fid, CALL)                  \
    switch (fid) {                                    \
    case FT63: return CALL(FT63);                     \
    case FT127: return CALL(FT127);                   \
    case FT191: return CALL(FT191);                   \
    case FT255: return CALL(FT255);                   \
    case FT253_192: return CALL(FT253_192);           \
    default: return cudaErrorInvalidValue;            \
    }

// ------------------------------------------------------------------ fold

// VEC adjacent columns per thread (2 for the one-limb field: 16-byte loads) and UNR rows per trip with all
// loads issued before the arithmetic: the fold is a stream over the matrix (8 B/coefficient) and what
// bounds it is the number of bytes in flight per SM, not the integer pipes.
template <int FID, int NT, int VEC>
__global__ void __launch_bounds__(128)
k_fold(const uint64_t *__restrict__ mat, size_t n_rows, size_t width, size_t row_stride,
       const uint64_t *__restrict__ tensors, uint64_t *__restrict__ out, size_t rows_per_split) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    constexpr int UNR = L == 1 ? 8 : (L == 2 ? 4 : 2);
    const size_t j = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (j >= width) return;
    const size_t r0 = (size_t)blockIdx.y * rows_per_split;
    const size_t r1 = r0 + rows_per_split < n_rows ? r0 + rows_per_split : n_rows;
    const bool full = j + VEC <= width;  // the last thread of an odd-width matrix owns one column only
    typename F::Dot acc[NT][VEC];  // unreduced sums for the multi-limb fields: one reduction per output, not per row
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
        for (int v = 0; v < VEC; v++) F::dot_init(acc[t][v]);
    size_t r = r0;
    for (; r + UNR <= r1; r += UNR) {
        E c[UNR][VEC];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            const uint64_t *p = mat + ((r + u) * row_stride + j) * L;
            if constexpr (L == 1 && VEC == 2) {
                if (full && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
                    const ulonglong2 q = *reinterpret_cast<const ulonglong2 *>(p);
                    c[u][0].v[0] = q.x;
                    c[u][1].v[0] = q.y;
                } else {
                    c[u][0].v[0] = p[0];
                    c[u][1].v[0] = full ? p[1] : 0;
                }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; v++) c[u][v] = (v == 0 || full) ? ld_fe<L>(p + v * L) : F::zero();
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; u++)
#pragma unroll
            for (int t = 0; t < NT; t++) {
                const E tv = ld_fe<L>(tensors + ((size_t)t * n_rows + r + u) * L);  // warp-uniform: broadcast
#pragma unroll
                for (int v = 0; v < VEC; v++) F::dot_mac(acc[t][v], c[u][v], tv);
            }
    }
    for (; r < r1; r++) {
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            if (v == 0 || full) {
                const E c = ld_fe<L>(mat + (r * row_stride + j + v) * L);
#pragma unroll
                for (int t = 0; t < NT; t++) F::dot_mac(acc[t][v], c, ld_fe<L>(tensors + ((size_t)t * n_rows + r) * L));
            }
        }
    }
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
        for (int v = 0; v < VEC; v++)
            if (v == 0 || full)
                st_fe<L>(out + (((size_t)blockIdx.y * NT + t) * width + j + v) * L, F::dot_finish(acc[t][v]));
}

// out[i] = sum_k parts[k*n + i]
template <int FID>
__global__ void k_add_partials(const uint64_t *__restrict__ parts, size_t n_parts, size_t n, uint64_t *__restrict__ out) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    E acc = F::zero();
    for (size_t k = 0; k < n_parts; k++) acc = F::add(acc, ld_fe<L>(parts + (k * n + i) * L));
    st_fe<L>(out + i * L, acc);
}

static size_t fold_splits(size_t n_rows, size_t width) {
    size_t tiles = (width + 127) / 128;  // an upper bound for the two-columns-per-thread case: more, smaller splits
    size_t want = (148 * 8 + tiles - 1) / tiles;  // aim for ~8 CTAs per SM
    size_t max_splits = (n_rows + 15) / 16;        // at least 16 rows per split
    size_t s = want < max_splits ? want : max_splits;
    return s < 1 ? 1 : s;
}

size_t fold_scratch_bytes(int fid, size_t n_rows, size_t width, size_t n_tensors) {
    size_t nt = n_tensors < 4 ? n_tensors : 4;
    return fold_splits(n_rows, width) * nt * width * field_consts(fid).limbs * sizeof(uint64_t);
}

template <int FID>
static cudaError_t fold_t(const uint64_t *d_mat, size_t n_rows, size_t width, size_t row_stride,
                          const uint64_t *d_tensors, size_t n_tensors, uint64_t *d_out, uint64_t *d_scratch,
                          const Launch &lc) {
    constexpr int L = Field<FID>::LIMBS;
    if (width == 0 || n_tensors == 0) return cudaSuccess;
    const size_t splits = fold_splits(n_rows, width);
    const size_t rps = (n_rows + splits - 1) / splits;
    constexpr int VEC = L == 1 ? 2 : 1;
    const unsigned gx = (unsigned)((width + 128 * VEC - 1) / (128 * VEC));
    size_t done = 0;
    while (done < n_tensors) {
        const size_t nt = n_tensors - done < 4 ? n_tensors - done : 4;
        const uint64_t *tens = d_tensors + done * n_rows * L;
        uint64_t *out = d_out + done * width * L;
        uint64_t *dst = splits == 1 ? out : d_scratch;
        dim3 grid(gx, (unsigned)splits);
        lc.begin("k_fold");
        switch (nt) {
        case 1: k_fold<FID, 1, VEC><<<grid, 128, 0, lc.s>>>(d_mat, n_rows, width, row_stride, tens, dst, rps); break;
        case 2: k_fold<FID, 2, VEC><<<grid, 128, 0, lc.s>>>(d_mat, n_rows, width, row_stride, tens, dst, rps); break;
        case 3: k_fold<FID, 3, VEC><<<grid, 128, 0, lc.s>>>(d_mat, n_rows, width, row_stride, tens, dst, rps); break;
        default: k_fold<FID, 4, VEC><<<grid, 128, 0, lc.s>>>(d_mat, n_rows, width, row_stride, tens, dst, rps); break;
        }
        lc.end();
        if (splits > 1) {
            const size_t n = nt * width;
            lc.begin("k_add_partials");
            k_add_partials<FID><<<(unsigned)((n + 255) / 256), 256, 0, lc.s>>>(d_scratch, splits, n, out);
            lc.end();
        }
        done += nt;
    }
    return cudaGetLastError();
}

cudaError_t fold(int fid, const uint64_t *d_mat, size_t n_rows, size_t width, size_t row_stride,
                 const uint64_t *d_tensors, size_t n_tensors, uint64_t *d_out, uint64_t *d_scratch,
                 const Launch &lc) {
#define CALL(F) fold_t<F>(d_mat, n_rows, width, row_stride, d_tensors, n_tensors, d_out, d_scratch, lc)
    LCPC_FIELD_SWITCH(fid, CALL)
#undef CALL
}

template <int FID>
static cudaError_t add_partials_t(const uint64_t *d_parts, size_t n_parts, size_t n, uint64_t *d_out, const Launch &lc) {
    if (n == 0) return cudaSuccess;
    lc.begin("k_add_partials");
    k_add_partials<FID><<<(unsigned)((n + 255) / 256), 256, 0, lc.s>>>(d_parts, n_parts, n, d_out);
    lc.end();
    return cudaGetLastError();
}

cudaError_t add_partials(int fid, const uint64_t *d_parts, size_t n_parts, size_t n, uint64_t *d_out,
                         const Launch &lc) {
#define CALL(F) add_partials_t<F>(d_parts, n_parts, n, d_out, lc)
    LCPC_FIELD_SWITCH(fid, CALL)
#undef CALL
}

// ------------------------------------------------------------------ column gather

template <int L>
__global__ void k_gather_columns(const uint64_t *__restrict__ mat, size_t n_rows, size_t row_stride,
                                 const uint64_t *__restrict__ cols, size_t n, uint64_t *__restrict__ out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * n_rows) return;
    const size_t i = t / n_rows, r = t % n_rows;
    st_fe<L>(out + t * L, ld_fe<L>(mat + (r * row_stride + (size_t)cols[i]) * L));
}

cudaError_t gather_columns(int fid, const uint64_t *d_mat, size_t n_rows, size_t row_stride,
                           const uint64_t *d_cols, size_t n, uint64_t *d_out, const Launch &lc) {
    const size_t total = n * n_rows;
    if (total == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    lc.begin("k_gather_columns");
    switch (field_consts(fid).limbs) {
    case 1: k_gather_columns<1><<<blocks, 256, 0, lc.s>>>(d_mat, n_rows, row_stride, d_cols, n, d_out); break;
    case 2: k_gather_columns<2><<<blocks, 256, 0, lc.s>>>(d_mat, n_rows, row_stride, d_cols, n, d_out); break;
    case 3: k_gather_columns<3><<<blocks, 256, 0, lc.s>>>(d_mat, n_rows, row_stride, d_cols, n, d_out); break;
    default: k_gather_columns<4><<<blocks, 256, 0, lc.s>>>(d_mat, n_rows, row_stride, d_cols, n, d_out); break;
    }
    lc.end();
    return cudaGetLastError();
}

// ------------------------------------------------------------------ verifier helpers

// PrimeField::to_repr() of each element (canonical value; big-endian bytes for Ft253_192): what
// FieldHash::to_hash_repr feeds the transcript (lcpc-2d/src/lib.rs:48-58)
template <int FID>
__global__ void k_to_canon(const uint64_t *__restrict__ in, size_t n, uint64_t *__restrict__ out) {
    using F = Field<FID>;
    constexpr int L = F::LIMBS;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_fe<L>(out + i * L, F::to_repr(ld_fe<L>(in + i * L)));
}

template <int FID>
static cudaError_t to_canon_t(const uint64_t *d_in, size_t n, uint64_t *d_out, const Launch &lc) {
    if (n == 0) return cudaSuccess;
    lc.begin("k_to_canon");
    k_to_canon<FID><<<(unsigned)((n + 255) / 256), 256, 0, lc.s>>>(d_in, n, d_out);
    lc.end();
    return cudaGetLastError();
}

cudaError_t to_canon(int fid, const uint64_t *d_in, size_t n, uint64_t *d_out, const Launch &lc) {
#define CALL(F) to_canon_t<F>(d_in, n, d_out, lc)
    LCPC_FIELD_SWITCH(fid, CALL)
#undef CALL
}

// verify_column_value for every (opened column i, tensor t) pair (lib.rs:1015-1030):
// out[i*n_tensors + t] = sum_r tensors[t][r] * cols[i][r]
template <int FID>
__global__ void k_column_dots(const uint64_t *__restrict__ cols, size_t n_rows, size_t n_open,
                              const uint64_t *__restrict__ tensors, size_t n_tensors, uint64_t *__restrict__ out) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_open * n_tensors) return;
    const size_t i = idx / n_tensors, t = idx % n_tensors;
    typename F::Dot acc;
    F::dot_init(acc);
    for (size_t r = 0; r < n_rows; r++)
        F::dot_mac(acc, ld_fe<L>(tensors + (t * n_rows + r) * L), ld_fe<L>(cols + (i * n_rows + r) * L));
    st_fe<L>(out + idx * L, F::dot_finish(acc));
}

template <int FID>
static cudaError_t column_dots_t(const uint64_t *d_cols, size_t n_rows, size_t n_open, const uint64_t *d_tensors,
                                 size_t n_tensors, uint64_t *d_out, const Launch &lc) {
    const size_t total = n_open * n_tensors;
    if (total == 0) return cudaSuccess;
    lc.begin("k_column_dots");
    k_column_dots<FID><<<(unsigned)((total + 63) / 64), 64, 0, lc.s>>>(d_cols, n_rows, n_open, d_tensors, n_tensors, d_out);
    lc.end();
    return cudaGetLastError();
}

cudaError_t column_dots(int fid, const uint64_t *d_cols, size_t n_rows, size_t n_open, const uint64_t *d_tensors,
                        size_t n_tensors, uint64_t *d_out, const Launch &lc) {
#define CALL(F) column_dots_t<F>(d_cols, n_rows, n_open, d_tensors, n_tensors, d_out, lc)
    LCPC_FIELD_SWITCH(fid, CALL)
#undef CALL
}

// ------------------------------------------------------------------ 7-byte packing

// element k = little-endian integer of bytes [7k, 7k+7), zero-extended; the value is
// stored as the limb itself (no Montgomery conversion -- writable_ft63.rs:35-40).
__global__ void k_pack_bytes7(const uint8_t *__restrict__ bytes, size_t n_bytes, uint64_t *__restrict__ elems,
                              size_t n_elems) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // group of 8 elements = 56 bytes
    const size_t e0 = g * 8;
    if (e0 >= n_elems) return;
    const size_t b0 = g * 56;
    const bool aligned = (reinterpret_cast<uintptr_t>(bytes) & 7) == 0;
    if (aligned && b0 + 56 <= n_bytes) {
        const uint64_t *w = reinterpret_cast<const uint64_t *>(bytes + b0);
        uint64_t x[7];
#pragma unroll
        for (int i = 0; i < 7; i++) x[i] = w[i];
        const uint64_t M = 0x00ffffffffffffffull;
        uint64_t o[8];
        o[0] = x[0] & M;
#pragma unroll
        for (int i = 1; i < 7; i++) o[i] = ((x[i - 1] >> (64 - 8 * i)) | (x[i] << (8 * i))) & M;
        o[7] = x[6] >> 8;
#pragma unroll
        for (int i = 0; i < 8; i++) elems[e0 + i] = o[i];
    } else {
        for (int i = 0; i < 8 && e0 + i < n_elems; i++) {
            uint64_t v = 0;
            for (int k = 0; k < 7; k++) {
                const size_t b = (e0 + i) * 7 + k;
                if (b < n_bytes) v |= (uint64_t)bytes[b] << (8 * k);
            }
            elems[e0 + i] = v;
        }
    }
}

cudaError_t pack_bytes7(const uint8_t *d_bytes, size_t n_bytes, uint64_t *d_elems, const Launch &lc) {
    const size_t n_elems = (n_bytes + 6) / 7;
    if (n_elems == 0) return cudaSuccess;
    const size_t groups = (n_elems + 7) / 8;
    lc.begin("k_pack_bytes7");
    k_pack_bytes7<<<(unsigned)((groups + 127) / 128), 128, 0, lc.s>>>(d_bytes, n_bytes, d_elems, n_elems);
    lc.end();
    return cudaGetLastError();
}

// ------------------------------------------------------------------ 31-byte packing (Ft253_192)

// Ft253_192::from_data_bytes (proof-of-storage/src/fields/ft253_192.rs:18-30): the 31 bytes are zero-padded to 32 and
// limb i = u64::from_be_bytes(padded[8i .. 8i+8]) -- limb 0 (least significant) comes from the FIRST eight bytes --
// stored directly as the Montgomery limbs.  The top limb is bytes 24..30 shifted left by 8, so any chunk whose byte 24
// exceeds 0x1f gives limbs >= p; the reference then computes on unreduced values (ff_derive's add drops the carry
// out of 2^256: the results are not field elements).  Such input is reported through *bad instead of being encoded.
__global__ void k_pack_bytes31(const uint8_t *__restrict__ bytes, size_t n_bytes, uint64_t *__restrict__ elems,
                               size_t n_elems, uint32_t *__restrict__ bad) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    // Every load is unconditional on a clamped address and the value is masked afterwards, and the limbs are put
    // together from 32-bit words: the first version (predicated byte loads summed into zeroed 64-bit register pairs)
    // returned 0x0c in the low byte of limb 0 whenever byte 6 of the group lay beyond the end of the file
    // (gpurun_out/dbg1.log of round 1; the stale value is the address offset of byte 12, held in the same register a
    // few instructions earlier).
    const size_t last = n_bytes - 1;  // n_elems > 0 implies n_bytes > 0
    uint32_t word[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t w = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int idx = 4 * j + k;  // position inside the zero-padded 32-byte group
            uint32_t byte = 0;
            if (idx < 31) {
                const size_t b = e * 31 + idx;
                const uint32_t v = bytes[b < n_bytes ? b : last];
                byte = b < n_bytes ? v : 0u;
            }
            w = (w << 8) | byte;
        }
        word[j] = w;
    }
    uint64_t limb[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        limb[i] = ((uint64_t)word[2 * i] << 32) | word[2 * i + 1];
        elems[e * 4 + i] = limb[i];
    }
    if (Field<FT253_192>::geq_p(limb)) atomicOr(bad, 1u);
}

cudaError_t pack_bytes31(const uint8_t *d_bytes, size_t n_bytes, uint64_t *d_elems, uint32_t *d_bad, const Launch &lc) {
    const size_t n_elems = (n_bytes + 30) / 31;
    if (n_elems == 0) return cudaSuccess;
    lc.begin("k_pack_bytes31");
    k_pack_bytes31<<<(unsigned)((n_elems + 255) / 256), 256, 0, lc.s>>>(d_bytes, n_bytes, d_elems, n_elems, d_bad);
    lc.end();
    return cudaGetLastError();
}

// ------------------------------------------------------------------ Brakedown

#ifndef LCPC_SPMV_TG_MAX
#define LCPC_SPMV_TG_MAX 8  // most lane groups of matrix rows one thread of k_spmv_tg serves (2 ... 8)
#endif
#ifndef LCPC_SPMV_TG_DEPTH
#define LCPC_SPMV_TG_DEPTH 3  // operand slots of k_spmv_tg (2 / 3 / 4: 0.448 / 0.423 / 0.452 ms at Ft63 2^24)
#endif
#ifndef LCPC_SPMV_DEPTH
#define LCPC_SPMV_DEPTH 0  // 0: three slots for the one-limb field, two otherwise (measured: profiles/r02_spmv.md)
#endif

// The expander levels run on a TRANSPOSED working copy xT[codeword index][matrix row] (leading
// dimension bp = matrix rows rounded up to the lane-group size): every gathered operand
// x[col_k][b .. b+GS) is then one contiguous run, the non-zero a[i,k] is loaded once per lane group
// instead of once per matrix row, and the output run is contiguous too.  A lane group of GS = 8, 16 or
// 32 lanes owns one output index i; a warp covers 32/GS of them.
// KSPLIT (the deep levels, a few hundred rows): the whole warp owns ONE output index and its 32/GS lane groups take
// every (32/GS)-th non-zero, then add up through shuffles -- those levels are a chain of dependent L2 round trips per
// thread and nothing else, so a chain four times shorter is a launch three times shorter.
template <int FID, bool KSPLIT>
__global__ void __launch_bounds__(128)
k_spmv_t(const uint32_t *__restrict__ rowptr, const uint32_t *__restrict__ colidx, const uint64_t *__restrict__ data,
         const uint32_t *__restrict__ order, size_t m_rows, const uint64_t *xT, uint64_t *yT, size_t bp, int log_gs) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    const int gs = 1 << log_gs, slices = 32 >> log_gs;
    const int lane = threadIdx.x & 31, sub = lane >> log_gs, bl = lane & (gs - 1);
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t slot = KSPLIT ? warp : warp * slices + sub;
    if (slot >= m_rows) return;
    const size_t i = order[slot];  // rows by decreasing length: the lane groups of a warp finish together
    const uint32_t k0 = rowptr[i], k1 = rowptr[i + 1];
    const uint32_t kfirst = KSPLIT ? k0 + (uint32_t)sub : k0, kstep = KSPLIT ? (uint32_t)slices : 1u;
    for (size_t b = (size_t)blockIdx.y * gs + bl; b < bp; b += (size_t)gridDim.y * gs) {
        typename F::DotW acc;  // `data` is pre-scaled by 2^32 (scale_csr_data)
        F::dotw_init(acc);
        // A non-zero costs two dependent trips to L2 (column index, then the gathered operand).  Software pipeline with a
        // ring of D slots: while a term is multiplied, the operands of the next D - 1 terms and the column indices of the D
        // terms after those are in flight (without it the kernel waits on the gathers with the integer pipes half idle:
        // long-scoreboard was the top stall, profiles/r02_spmv.md).
        constexpr int D = LCPC_SPMV_DEPTH ? LCPC_SPMV_DEPTH : (L == 1 ? 3 : 2);
        E a[D], x[D];
        uint32_t cin[D];
        uint32_t k = kfirst;
        if (k < k1) {
            // loads past the end of the row are clamped to its last non-zero (loaded, never multiplied): no predicates,
            // no zero fill
            const uint32_t klast = k1 - 1;
#pragma unroll
            for (int d = 0; d < D; d++) cin[d] = colidx[min(k + d * kstep, klast)];
#pragma unroll
            for (int d = 0; d < D; d++) {
                const uint32_t kk = min(k + d * kstep, klast);
                a[d] = ld_fe<L>(data + (size_t)kk * L);
                x[d] = ld_fe<L>(xT + ((size_t)cin[d] * bp + b) * L);
                cin[d] = colidx[min(k + (D + d) * kstep, klast)];
            }
            do {
#pragma unroll
                for (int d = 0; d < D; d++) {
                    if (d == 0 || k + d * kstep < k1) F::dotw_mac(acc, a[d], x[d]);
                    const uint32_t kk = min(k + (D + d) * kstep, klast);
                    a[d] = ld_fe<L>(data + (size_t)kk * L);
                    x[d] = ld_fe<L>(xT + ((size_t)cin[d] * bp + b) * L);
                    cin[d] = colidx[min(k + (2 * D + d) * kstep, klast)];
                }
                k += D * kstep;
            } while (k < k1);
        }
        E r = F::dotw_finish_prescaled(acc);
        if constexpr (KSPLIT) {
            for (int off = gs; off < 32; off <<= 1) {
                E o;
#pragma unroll
                for (int l = 0; l < L; l++) o.v[l] = __shfl_xor_sync(0xffffffffu, r.v[l], off);
                r = F::add(r, o);
            }
            if (sub == 0) st_fe<L>(yT + (i * bp + b) * L, r);
        } else {
            st_fe<L>(yT + (i * bp + b) * L, r);
        }
    }
}

// One-limb field, wide levels: a thread serves NG lane groups of matrix rows (b = (g0 + j) * GS + bl) for its output
// index, so the column index and the non-zero are loaded once per NG products instead of once per product and the NG
// gathers of a term are issued back to back from one address (GS is a template parameter: the offsets are immediates).
// Lane groups past the last one read the start of the next xT row (the working copy is padded by NG * GS elements) and
// are not stored.
template <int FID, int NG, int LOG_GS>
__global__ void __launch_bounds__(128)
k_spmv_tg(const uint32_t *__restrict__ rowptr, const uint32_t *__restrict__ colidx, const uint64_t *__restrict__ data,
          const uint32_t *__restrict__ order, size_t m_rows, const uint64_t *xT, uint64_t *yT, size_t bp) {
    using F = Field<FID>;
    using E = typename F::E;
    static_assert(F::LIMBS == 1, "one-limb fields");
    constexpr int GS = 1 << LOG_GS, SLICES = 32 >> LOG_GS;
    const int lane = threadIdx.x & 31, sub = lane >> LOG_GS, bl = lane & (GS - 1);
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t slot = warp * SLICES + sub;
    if (slot >= m_rows) return;
    const size_t i = order[slot];
    const uint32_t k0 = rowptr[i], k1 = rowptr[i + 1];
    const size_t b0 = (size_t)blockIdx.y * NG * GS + bl;
    typename F::DotW acc[NG];  // `data` is pre-scaled by 2^32 (scale_csr_data)
#pragma unroll
    for (int j = 0; j < NG; j++) F::dotw_init(acc[j]);
    if (k0 < k1) {
        // DG slots: the gathers of the next DG - 1 terms are in flight while this one is multiplied; indices DG terms
        // further.  Loads past the end of the row are clamped to its last non-zero.
        constexpr int DG = LCPC_SPMV_TG_DEPTH;
        const uint32_t klast = k1 - 1;
        const uint64_t *xb = xT + b0;
        E a[DG], x[DG][NG];
        uint32_t cin[DG];
#pragma unroll
        for (int d = 0; d < DG; d++) cin[d] = colidx[min(k0 + d, klast)];
#pragma unroll
        for (int d = 0; d < DG; d++) {
            a[d] = ld_fe<1>(data + min(k0 + d, klast));
            const uint64_t *xp = xb + (size_t)cin[d] * bp;
#pragma unroll
            for (int j = 0; j < NG; j++) x[d][j] = ld_fe<1>(xp + j * GS);
            cin[d] = colidx[min(k0 + DG + d, klast)];
        }
        uint32_t k = k0;
        do {
#pragma unroll
            for (int d = 0; d < DG; d++) {
                if (d == 0 || k + d < k1) {
#pragma unroll
                    for (int j = 0; j < NG; j++) F::dotw_mac(acc[j], a[d], x[d][j]);
                }
                a[d] = ld_fe<1>(data + min(k + DG + d, klast));
                const uint64_t *xp = xb + (size_t)cin[d] * bp;
#pragma unroll
                for (int j = 0; j < NG; j++) x[d][j] = ld_fe<1>(xp + j * GS);
                cin[d] = colidx[min(k + 2 * DG + d, klast)];
            }
            k += DG;
        } while (k < k1);
    }
#pragma unroll
    for (int j = 0; j < NG; j++)
        if (b0 + (size_t)j * GS < bp) st_fe<1>(yT + i * bp + b0 + (size_t)j * GS, F::dotw_finish_prescaled(acc[j]));
}

// data[k] *= 2^32 (every field): the lazy dot product of k_spmv_t reduces by one extra word
template <int FID>
__global__ void k_scale_data(uint64_t *data, size_t n) {
    using F = Field<FID>;
    constexpr int L = F::LIMBS;
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    st_fe<L>(data + k * L, F::mul(ld_fe<L>(data + k * L), F::dotw_scale()));
}

template <int FID>
static cudaError_t scale_csr_data_t(uint64_t *d_data, size_t nnz, cudaStream_t s) {
    if (nnz == 0) return cudaSuccess;
    k_scale_data<FID><<<(unsigned)((nnz + 255) / 256), 256, 0, s>>>(d_data, nnz);
    return cudaGetLastError();
}

cudaError_t scale_csr_data(int fid, uint64_t *d_data, size_t nnz, cudaStream_t s) {
#define CALL(F) scale_csr_data_t<F>(d_data, nnz, s)
    LCPC_FIELD_SWITCH(fid, CALL)
#undef CALL
}

// xoT[r][b] = sum_j xiT[j][b] * (r+1)^j by Horner (encode.rs:97-109), transposed layout
template <int FID>
__global__ void k_reed_solomon_t(const uint64_t *__restrict__ xiT, size_t n_in, uint64_t *xoT, size_t n_out, size_t bp) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bp * n_out) return;
    const size_t r = t / bp, b = t % bp;
    E x = F::zero();  // (r+1) in Montgomery form = sum of (r+1) ones
    const E one = F::one();
    for (size_t i = 0; i <= r; i++) x = F::add(x, one);
    E acc = F::zero();
    for (size_t jj = n_in; jj-- > 0;) acc = F::add(F::mul(acc, x), ld_fe<L>(xiT + (jj * bp + b) * L));
    st_fe<L>(xoT + (r * bp + b) * L, acc);
}

// dst[c][r] = src[r][c] for r < n_r, c < n_c (32x32 element tiles through shared memory); rows
// r in [n_r, dst_ld) of dst are written as zero when zero_pad is set.
template <int L>
__global__ void __launch_bounds__(256)
k_transpose(const uint64_t *__restrict__ src, size_t n_r, size_t n_c, size_t src_ld, uint64_t *__restrict__ dst,
            size_t dst_ld, int zero_pad, uint64_t *__restrict__ copy, size_t copy_ld, const __grid_constant__ ScatterDst sc,
            int sc_mode, size_t sc_col0) {
    // sc_mode 1: the row-major copy, 2: the transposed output itself goes to the column-block matrices of the ranks that
    // own those columns (their own HBM or a peer's over NVLink) instead of a local matrix: element (matrix row, column)
    // lands at base[column >> log_cb][row0 + matrix row][column & (cb - 1)]
    auto peer = [&](size_t row, size_t col) {
        return sc.base[col >> sc.log_cb] + ((((size_t)sc.row0 + row) << sc.log_cb) + (col & (((size_t)1 << sc.log_cb) - 1))) * L;
    };
    __shared__ uint64_t tile[L][32][33];
    const size_t c0 = (size_t)blockIdx.x * 32, r0 = (size_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int rr = ty; rr < 32; rr += 8) {
        const size_t r = r0 + rr, c = c0 + tx;
        Fe<L> v;
#pragma unroll
        for (int l = 0; l < L; l++) v.v[l] = 0;
        if (r < n_r && c < n_c) {
            v = ld_fe<L>(src + (r * src_ld + c) * L);
            if (sc_mode == 1) st_fe<L>(peer(r, sc_col0 + c), v);
            else if (copy) st_fe<L>(copy + (r * copy_ld + c) * L, v);  // the same element, row-major, into a wider matrix
        }
#pragma unroll
        for (int l = 0; l < L; l++) tile[l][rr][tx] = v.v[l];
    }
    __syncthreads();
    for (int cc = ty; cc < 32; cc += 8) {
        const size_t c = c0 + cc, r = r0 + tx;
        if (c < n_c && (r < n_r || (zero_pad && r < dst_ld))) {
            Fe<L> v;
#pragma unroll
            for (int l = 0; l < L; l++) v.v[l] = tile[l][tx][cc];
            if (sc_mode == 2) st_fe<L>(peer(c, sc_col0 + r), v);
            else st_fe<L>(dst + (c * dst_ld + r) * L, v);
        }
    }
}

template <int L>
static void transpose_launch(const uint64_t *src, size_t n_r, size_t n_c, size_t src_ld, uint64_t *dst, size_t dst_ld,
                             int zero_pad, const Launch &lc, uint64_t *copy = nullptr, size_t copy_ld = 0,
                             const ScatterDst *sc = nullptr, int sc_mode = 0, size_t sc_col0 = 0) {
    if (n_r == 0 || n_c == 0) return;
    const size_t rows_cov = zero_pad ? dst_ld : n_r;
    dim3 grid((unsigned)((n_c + 31) / 32), (unsigned)((rows_cov + 31) / 32));
    lc.begin("k_transpose");
    k_transpose<L><<<grid, 256, 0, lc.s>>>(src, n_r, n_c, src_ld, dst, dst_ld, zero_pad, copy, copy_ld, sc ? *sc : ScatterDst{},
                                           sc ? sc_mode : 0, sc_col0);
    lc.end();
}

// lane-group size (log2) that wastes the fewest padded lanes for this many matrix rows
static int sdig_log_gs(size_t n_rows) {
    int best = 5;
    size_t best_bp = (n_rows + 31) / 32 * 32;
    for (int lg = 4; lg >= 3; lg--) {
        size_t g = (size_t)1 << lg, bp = (n_rows + g - 1) / g * g;
        if (bp < best_bp) { best_bp = bp; best = lg; }
    }
    return best;
}
static size_t sdig_bp(size_t n_rows) {
    size_t g = (size_t)1 << sdig_log_gs(n_rows);
    return (n_rows + g - 1) / g * g;
}

// transposed working copy: n_cols codeword rows + the (unstored) output of the last precode
size_t sdig_tmp_elems(const SdigPlan &plan, size_t n_rows) {
    // + 8 lane groups of 32: k_spmv_tg reads (and discards) up to NG * GS elements past the row it gathers from
    return plan.pre.empty() ? 0 : (plan.n_cols + plan.pre.back().rows) * sdig_bp(n_rows) + 8 * 32;
}

template <int FID>
static cudaError_t sdig_encode_t(const SdigPlan &plan, const uint64_t *d_msg, size_t msg_ld, uint64_t *d_comm, size_t n_rows,
                                 uint64_t *d_tmp, const Launch &lc, const ScatterDst *sc) {
    constexpr int L = Field<FID>::LIMBS;
    const size_t nl = plan.pre.size();
    if (nl == 0 || n_rows == 0) return cudaSuccess;
    const size_t n_cols = plan.n_cols, npr = plan.n_per_row;
    const int log_gs = sdig_log_gs(n_rows);
    const size_t bp = sdig_bp(n_rows), gs = (size_t)1 << log_gs;
    uint64_t *xT = d_tmp;                        // [n_cols][bp]
    uint64_t *tT = d_tmp + n_cols * bp * L;      // [rows of the last precode][bp]
    // message columns -> transposed copy (padded lanes zero, so every derived lane stays zero); when the message is not
    // in place yet (commit: the coefficient matrix), the same pass writes it into the first n_per_row columns of comm --
    // the computed columns are all written by the last transpose, so comm needs no zero fill and no separate widening pass
    if (sc) transpose_launch<L>(d_msg, n_rows, npr, msg_ld, xT, bp, 1, lc, nullptr, 0, sc, 1, 0);
    else if (d_msg == d_comm) transpose_launch<L>(d_comm, n_rows, npr, n_cols, xT, bp, 1, lc);
    else transpose_launch<L>(d_msg, n_rows, npr, msg_ld, xT, bp, 1, lc, d_comm, n_cols);
    auto spmv = [&](const DevCsr &m, size_t x_off, uint64_t *y) {
        if (m.rows == 0) return;
        const size_t slices = (size_t)32 >> log_gs, groups = bp / gs;
        // too few warps to fill the machine: one row per warp, its non-zeros split over the lane groups
        const bool ksplit = slices > 1 && (m.rows / slices + 1) * groups < (size_t)148 * 16;
        const size_t rows_per_cta = ksplit ? 4 : 4 * slices;
        dim3 grid((unsigned)((m.rows + rows_per_cta - 1) / rows_per_cta), (unsigned)(groups < 65535 ? groups : 65535));
        if constexpr (L == 1) {
            // wide levels of the one-limb field: NG lane groups per thread (the NG that wastes the fewest group slots)
            if (!ksplit && groups > 1) {
                // the largest NG that leaves at most 15 % of the group slots empty, else the one with the fewest empty slots
                int ng = 0, best_ng = 2;
                double best_eff = 0.0;
                for (int c = LCPC_SPMV_TG_MAX; c >= 2 && !ng; c--) {
                    const double eff = (double)groups / (double)((groups + c - 1) / c * c);
                    if (eff >= 0.85) ng = c;
                    if (eff > best_eff) { best_eff = eff; best_ng = c; }
                }
                if (!ng) ng = best_ng;
                const size_t ychunks = (groups + ng - 1) / ng;
                if ((m.rows / slices + 1) * ychunks >= (size_t)148 * 16) {
                    dim3 gridg((unsigned)((m.rows + 4 * slices - 1) / (4 * slices)), (unsigned)ychunks);
                    lc.begin("k_spmv_t");
#define LCPC_TG2(NGV, LG) k_spmv_tg<FID, NGV, LG><<<gridg, 128, 0, lc.s>>>(m.d_rowptr, m.d_colidx, m.d_data, m.d_order, m.rows, xT + x_off * bp * L, y, bp)
#define LCPC_TG(NGV) case NGV: if (log_gs == 3) LCPC_TG2(NGV, 3); else if (log_gs == 4) LCPC_TG2(NGV, 4); else LCPC_TG2(NGV, 5); break;
                    switch (ng) { LCPC_TG(2) LCPC_TG(3) LCPC_TG(4) LCPC_TG(5) LCPC_TG(6) LCPC_TG(7) LCPC_TG(8) }
#undef LCPC_TG2
#undef LCPC_TG
                    lc.end();
                    return;
                }
            }
        }
        lc.begin("k_spmv_t");
        if (ksplit) k_spmv_t<FID, true><<<grid, 128, 0, lc.s>>>(m.d_rowptr, m.d_colidx, m.d_data, m.d_order, m.rows, xT + x_off * bp * L, y, bp, log_gs);
        else k_spmv_t<FID, false><<<grid, 128, 0, lc.s>>>(m.d_rowptr, m.d_colidx, m.d_data, m.d_order, m.rows, xT + x_off * bp * L, y, bp, log_gs);
        lc.end();
    };
    // encode.rs:46-58 precodes all the way down
    size_t in_start = 0;
    for (size_t l = 0; l + 1 < nl; l++) {
        const size_t in_end = in_start + plan.pre[l].cols;
        spmv(plan.pre[l], in_start, xT + in_end * bp * L);
        in_start = in_end;
    }
    // encode.rs:61-74 base case: last precode into the temporary, Reed-Solomon of that
    const DevCsr &lp = plan.pre[nl - 1];
    const size_t in_end = in_start + lp.cols;
    spmv(lp, in_start, tT);
    const size_t n_rs = plan.post[nl - 1].cols;
    if (n_rs) {
        const size_t total = bp * n_rs;
        lc.begin("k_reed_solomon_t");
        k_reed_solomon_t<FID><<<(unsigned)((total + 127) / 128), 128, 0, lc.s>>>(tT, lp.rows, xT + in_end * bp * L, n_rs, bp);
        lc.end();
    }
    in_start = in_end + lp.rows;
    size_t out_start = in_end + n_rs;
    // encode.rs:76-90 postcodes back up
    for (size_t l = nl; l-- > 0;) {
        in_start -= plan.pre[l].rows;
        spmv(plan.post[l], in_start, xT + out_start * bp * L);
        out_start += plan.post[l].rows;
    }
    // computed part of the codeword back to the row-major matrix (the message columns are in place)
    if (sc) transpose_launch<L>(xT + npr * bp * L, n_cols - npr, n_rows, bp, nullptr, 0, 0, lc, nullptr, 0, sc, 2, npr);
    else transpose_launch<L>(xT + npr * bp * L, n_cols - npr, n_rows, bp, d_comm + npr * L, n_cols, 0, lc);
    return cudaGetLastError();
}

cudaError_t sdig_encode(const SdigPlan &plan, const uint64_t *d_msg, size_t msg_ld, uint64_t *d_comm, size_t n_rows,
                        uint64_t *d_tmp, const Launch &lc, const ScatterDst *sc) {
#define CALL(F) sdig_encode_t<F>(plan, d_msg, msg_ld, d_comm, n_rows, d_tmp, lc, sc)
    LCPC_FIELD_SWITCH(plan.fid, CALL)
#undef CALL
}

}  // namespace lcpc
