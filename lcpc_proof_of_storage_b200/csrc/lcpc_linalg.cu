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
    default: return cudaErrorInvalidValue;            \
    }

// ------------------------------------------------------------------ fold

template <int FID, int NT>
__global__ void __launch_bounds__(128)
k_fold(const uint64_t *__restrict__ mat, size_t n_rows, size_t width, size_t row_stride,
       const uint64_t *__restrict__ tensors, uint64_t *__restrict__ out, size_t rows_per_split) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= width) return;
    const size_t r0 = (size_t)blockIdx.y * rows_per_split;
    const size_t r1 = r0 + rows_per_split < n_rows ? r0 + rows_per_split : n_rows;
    E acc[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) acc[t] = F::zero();
    for (size_t r = r0; r < r1; r++) {
        const E c = ld_fe<L>(mat + (r * row_stride + j) * L);
#pragma unroll
        for (int t = 0; t < NT; t++) {
            const E tv = ld_fe<L>(tensors + ((size_t)t * n_rows + r) * L);  // warp-uniform: broadcast
            acc[t] = F::add(acc[t], F::mul(c, tv));
        }
    }
#pragma unroll
    for (int t = 0; t < NT; t++) st_fe<L>(out + (((size_t)blockIdx.y * NT + t) * width + j) * L, acc[t]);
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
    size_t tiles = (width + 127) / 128;
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
    const unsigned gx = (unsigned)((width + 127) / 128);
    size_t done = 0;
    while (done < n_tensors) {
        const size_t nt = n_tensors - done < 4 ? n_tensors - done : 4;
        const uint64_t *tens = d_tensors + done * n_rows * L;
        uint64_t *out = d_out + done * width * L;
        uint64_t *dst = splits == 1 ? out : d_scratch;
        dim3 grid(gx, (unsigned)splits);
        lc.begin("k_fold");
        switch (nt) {
        case 1: k_fold<FID, 1><<<grid, 128, 0, lc.s>>>(d_mat, n_rows, width, row_stride, tens, dst, rps); break;
        case 2: k_fold<FID, 2><<<grid, 128, 0, lc.s>>>(d_mat, n_rows, width, row_stride, tens, dst, rps); break;
        case 3: k_fold<FID, 3><<<grid, 128, 0, lc.s>>>(d_mat, n_rows, width, row_stride, tens, dst, rps); break;
        default: k_fold<FID, 4><<<grid, 128, 0, lc.s>>>(d_mat, n_rows, width, row_stride, tens, dst, rps); break;
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

// canonical (de-Montgomery) limbs of each element: what FieldHash::to_hash_repr feeds the
// transcript (lcpc-2d/src/lib.rs:48-58)
template <int FID>
__global__ void k_to_canon(const uint64_t *__restrict__ in, size_t n, uint64_t *__restrict__ out) {
    using F = Field<FID>;
    constexpr int L = F::LIMBS;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_fe<L>(out + i * L, F::to_canon(ld_fe<L>(in + i * L)));
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
    E acc = F::zero();
    for (size_t r = 0; r < n_rows; r++)
        acc = F::add(acc, F::mul(ld_fe<L>(tensors + (t * n_rows + r) * L), ld_fe<L>(cols + (i * n_rows + r) * L)));
    st_fe<L>(out + idx * L, acc);
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

// ------------------------------------------------------------------ Brakedown

// y[b][i] = sum_k data[k] * x[b][colidx[k]] over CSR row i, for every matrix row b of the batch
template <int FID>
__global__ void __launch_bounds__(128)
k_spmv_batch(const uint32_t *__restrict__ rowptr, const uint32_t *__restrict__ colidx, const uint64_t *__restrict__ data,
             size_t m_rows, const uint64_t *x_base, size_t x_stride, uint64_t *y_base, size_t y_stride, size_t batch) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m_rows) return;
    const uint32_t k0 = rowptr[i], k1 = rowptr[i + 1];
    for (size_t b = blockIdx.y; b < batch; b += gridDim.y) {
        E acc = F::zero();
        for (uint32_t k = k0; k < k1; k++) {
            const E a = ld_fe<L>(data + (size_t)k * L);
            const E x = ld_fe<L>(x_base + (b * x_stride + colidx[k]) * L);
            acc = F::add(acc, F::mul(a, x));
        }
        st_fe<L>(y_base + (b * y_stride + i) * L, acc);
    }
}

// xo[b][r] = sum_j xi[b][j] * (r+1)^j by Horner (encode.rs:97-109)
template <int FID>
__global__ void k_reed_solomon(const uint64_t *__restrict__ xi, size_t xi_stride, size_t n_in, uint64_t *xo,
                               size_t xo_stride, size_t n_out, size_t batch) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= batch * n_out) return;
    const size_t b = t / n_out, r = t % n_out;
    // x = (r+1) in Montgomery form = sum of (r+1) ones
    E x = F::zero();
    const E one = F::one();
    for (size_t i = 0; i <= r; i++) x = F::add(x, one);
    E acc = F::zero();
    for (size_t j = n_in; j-- > 0;) acc = F::add(F::mul(acc, x), ld_fe<L>(xi + (b * xi_stride + j) * L));
    st_fe<L>(xo + (b * xo_stride + r) * L, acc);
}

template <int L>
__global__ void k_widen_rows(const uint64_t *__restrict__ coeffs, size_t n_per_row, uint64_t *__restrict__ comm,
                             size_t n_cols, size_t n_rows) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cols) return;
    for (size_t r = blockIdx.y; r < n_rows; r += gridDim.y) {
        Fe<L> v;
#pragma unroll
        for (int l = 0; l < L; l++) v.v[l] = 0;
        if (j < n_per_row) v = ld_fe<L>(coeffs + (r * n_per_row + j) * L);
        st_fe<L>(comm + (r * n_cols + j) * L, v);
    }
}

cudaError_t widen_rows(int fid, const uint64_t *d_coeffs, size_t n_per_row, uint64_t *d_comm, size_t n_cols,
                       size_t n_rows, const Launch &lc) {
    if (n_rows == 0 || n_cols == 0) return cudaSuccess;
    dim3 grid((unsigned)((n_cols + 255) / 256), (unsigned)(n_rows < 65535 ? n_rows : 65535));
    lc.begin("k_widen_rows");
    switch (field_consts(fid).limbs) {
    case 1: k_widen_rows<1><<<grid, 256, 0, lc.s>>>(d_coeffs, n_per_row, d_comm, n_cols, n_rows); break;
    case 2: k_widen_rows<2><<<grid, 256, 0, lc.s>>>(d_coeffs, n_per_row, d_comm, n_cols, n_rows); break;
    case 3: k_widen_rows<3><<<grid, 256, 0, lc.s>>>(d_coeffs, n_per_row, d_comm, n_cols, n_rows); break;
    default: k_widen_rows<4><<<grid, 256, 0, lc.s>>>(d_coeffs, n_per_row, d_comm, n_cols, n_rows); break;
    }
    lc.end();
    return cudaGetLastError();
}

size_t sdig_tmp_elems(const SdigPlan &plan, size_t n_rows) {
    return plan.pre.empty() ? 0 : n_rows * plan.pre.back().rows;
}

template <int FID>
static cudaError_t sdig_encode_t(const SdigPlan &plan, uint64_t *d_comm, size_t n_rows, uint64_t *d_tmp,
                                 const Launch &lc) {
    constexpr int L = Field<FID>::LIMBS;
    const size_t nl = plan.pre.size();
    if (nl == 0 || n_rows == 0) return cudaSuccess;
    const size_t stride = plan.n_cols;
    const unsigned gy = (unsigned)(n_rows < 65535 ? n_rows : 65535);
    auto spmv = [&](const DevCsr &m, const uint64_t *x, size_t xs, uint64_t *y, size_t ys) {
        if (m.rows == 0) return;
        dim3 grid((unsigned)((m.rows + 127) / 128), gy);
        lc.begin("k_spmv_batch");
        k_spmv_batch<FID><<<grid, 128, 0, lc.s>>>(m.d_rowptr, m.d_colidx, m.d_data, m.rows, x, xs, y, ys, n_rows);
        lc.end();
    };
    // encode.rs:46-58 precodes all the way down
    size_t in_start = 0;
    for (size_t l = 0; l + 1 < nl; l++) {
        const size_t in_end = in_start + plan.pre[l].cols;
        spmv(plan.pre[l], d_comm + in_start * L, stride, d_comm + in_end * L, stride);
        in_start = in_end;
    }
    // encode.rs:61-74 base case
    const DevCsr &lp = plan.pre[nl - 1];
    const size_t in_end = in_start + lp.cols;
    spmv(lp, d_comm + in_start * L, stride, d_tmp, lp.rows);
    const size_t n_rs = plan.post[nl - 1].cols;
    {
        const size_t total = n_rows * n_rs;
        if (total) {
            lc.begin("k_reed_solomon");
            k_reed_solomon<FID><<<(unsigned)((total + 127) / 128), 128, 0, lc.s>>>(d_tmp, lp.rows, lp.rows,
                                                                                 d_comm + in_end * L, stride, n_rs, n_rows);
            lc.end();
        }
    }
    in_start = in_end + lp.rows;
    size_t out_start = in_end + n_rs;
    // encode.rs:76-90 postcodes back up
    for (size_t l = nl; l-- > 0;) {
        in_start -= plan.pre[l].rows;
        spmv(plan.post[l], d_comm + in_start * L, stride, d_comm + out_start * L, stride);
        out_start += plan.post[l].rows;
    }
    return cudaGetLastError();
}

cudaError_t sdig_encode(const SdigPlan &plan, uint64_t *d_comm, size_t n_rows, uint64_t *d_tmp, const Launch &lc) {
#define CALL(F) sdig_encode_t<F>(plan, d_comm, n_rows, d_tmp, lc)
    LCPC_FIELD_SWITCH(plan.fid, CALL)
#undef CALL
}

}  // namespace lcpc
