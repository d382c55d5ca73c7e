// Column-leaf hashing fused with the column gather, and the Merkle reduction.
//
// Leaf j = BLAKE3(0^32 || repr(M[0][j]) || ... || repr(M[n_rows-1][j])) with repr = the
// canonical little-endian bytes of the element (lcpc-2d/src/lib.rs:736-775, FieldHash
// :35-59); internal node = BLAKE3(left || right) (:777-815); tree array
// [np2 leaves | np2/2 | ... | root] with all-zero padding leaves (:685-695, :720-734).
//
// B200 mapping: the matrix stays row-major.  One thread owns one (column, BLAKE3 chunk)
// pair: adjacent threads read adjacent columns of the same row, so the strided "column
// gather" is a sequence of fully coalesced row-segment loads, and the de-Montgomery
// reduction happens in registers on the way into the message block.  BLAKE3's own chunk
// tree gives a second grid dimension (1024-byte chunks are independent), which is what
// fills 148 SMs when n_cols alone is only 2^16; a second small kernel folds the chunk
// chaining values per column.  Merkle levels are reduced 9 at a time inside one CTA.
#include <cstdlib>

#include "lcpc_blake3.cuh"
#include "lcpc_field.cuh"
#include "lcpc_kernels.h"

#ifndef LCPC_HASH_CTAS1
#define LCPC_HASH_CTAS1 7  // resident 128-thread CTAs per SM for the one-limb field (register budget 72: measured best of 5/7/8, profiles/r02_hash_tail.md)
#endif

namespace lcpc {

// ------------------------------------------------------------------ leaf hashing

// Message words of BLAKE3 block `B` (global block index over the column's byte stream)
// for column `col`.  Stream = 8 zero words, then 2L words per row.
template <int FID>
__device__ __forceinline__ void load_block_words(uint32_t m[16], const uint64_t *__restrict__ mat, size_t n_rows,
                                                 size_t row_stride, size_t col, uint64_t B, int64_t row_base) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    constexpr int WPE = 2 * L;                  // 32-bit words per element
    constexpr int K = (8 + WPE - 1) / WPE;      // virtual zero elements covering the prefix
    constexpr int O = WPE * K - 8;              // word offset of the prefix inside them
    constexpr int NE = (16 + WPE - 1) / WPE + 1;  // elements a block can touch (upper bound)
    // virtual word index of m[0]
    const uint64_t vw0 = 16 * B + O;
    const uint64_t ve0 = vw0 / WPE;
    const int sw0 = (int)(vw0 % WPE);
    if constexpr (16 % WPE == 0 && O == 0) {
        // whole elements per block (L = 1, 2, 4)
        constexpr int EPB = 16 / WPE;
#pragma unroll
        for (int e = 0; e < EPB; e++) {
            const int64_t row = (int64_t)(ve0 + e) - K;
            E c = F::zero();
            if (row >= 0 && (uint64_t)row < n_rows) c = F::to_repr(ld_fe<L>(mat + ((size_t)(row - row_base) * row_stride + col) * L));
#pragma unroll
            for (int l = 0; l < L; l++) {
                m[e * WPE + 2 * l] = (uint32_t)c.v[l];
                m[e * WPE + 2 * l + 1] = (uint32_t)(c.v[l] >> 32);
            }
        }
    } else {
        // elements straddle blocks (L = 3): three phases of sw0, resolved by a switch so
        // that every message index stays a compile-time constant
        uint32_t w[NE * WPE];
#pragma unroll
        for (int e = 0; e < NE; e++) {
            const int64_t row = (int64_t)(ve0 + e) - K;
            E c = F::zero();
            if (row >= 0 && (uint64_t)row < n_rows) c = F::to_repr(ld_fe<L>(mat + ((size_t)(row - row_base) * row_stride + col) * L));
#pragma unroll
            for (int l = 0; l < L; l++) {
                w[e * WPE + 2 * l] = (uint32_t)c.v[l];
                w[e * WPE + 2 * l + 1] = (uint32_t)(c.v[l] >> 32);
            }
        }
        switch (sw0) {
        case 0:
#pragma unroll
            for (int i = 0; i < 16; i++) m[i] = w[i];
            break;
        case 2:
#pragma unroll
            for (int i = 0; i < 16; i++) m[i] = w[i + 2];
            break;
        default:
#pragma unroll
            for (int i = 0; i < 16; i++) m[i] = w[i + 4];
            break;
        }
    }
}

// Raw (Montgomery) elements of BLAKE3 block `B` for fields whose elements do not straddle blocks
// (LIMBS 1, 2, 4): rows outside [0, n_rows) -- the 32-byte zero prefix and the tail -- read as zero.
template <int FID>
struct BlockElems {
    static constexpr int L = Field<FID>::LIMBS;
    static constexpr int WPE = 2 * L;
    static constexpr bool WHOLE = (16 % WPE == 0);
    static constexpr int EPB = WHOLE ? 16 / WPE : 1;
    static constexpr int K = 8 / WPE;  // prefix length in elements
    typename Field<FID>::E e[EPB];

    __device__ __forceinline__ void load(const uint64_t *__restrict__ mat, size_t n_rows, size_t row_stride, size_t col,
                                         uint64_t B, int64_t row_base) {
#pragma unroll
        for (int i = 0; i < EPB; i++) {
            const int64_t row = (int64_t)(B * EPB + i) - K;
            e[i] = (row >= 0 && (uint64_t)row < n_rows) ? ld_fe<L>(mat + ((size_t)(row - row_base) * row_stride + col) * L)
                                                        : Field<FID>::zero();
        }
    }
    // to_repr() bytes of the block as little-endian message words (the de-Montgomery reduction of zero is zero)
    __device__ __forceinline__ void words(uint32_t m[16]) const {
#pragma unroll
        for (int i = 0; i < EPB; i++) {
            const typename Field<FID>::E c = Field<FID>::to_repr(e[i]);
#pragma unroll
            for (int l = 0; l < L; l++) {
                m[i * WPE + 2 * l] = (uint32_t)c.v[l];
                m[i * WPE + 2 * l + 1] = (uint32_t)(c.v[l] >> 32);
            }
        }
    }
};

// One thread per (column, chunk): chaining value of that chunk, or the leaf itself when
// the whole message is a single chunk.  The loads of block b+1 are issued before block b is
// compressed, so the L2/HBM latency hides behind ~900 ALU instructions.
// `mat` may be a window of the matrix: its first row is global row `row_base`, rows in [0, n_rows) are
// valid (the rest of the byte stream reads as zero), and only chunks [chunk0, chunk_end) are computed
// (streaming commits hash chunks as their rows arrive; row edits re-hash only the chunks they touch).
// Chaining value of chunk `c` of column `col` (the leaf itself when the whole message is a single chunk).
template <int FID>
__device__ __forceinline__ void chunk_cv(uint32_t cv[8], const uint64_t *__restrict__ mat, size_t n_rows, size_t row_stride,
                                         size_t col, uint64_t total_bytes, uint64_t n_chunks, int64_t row_base, uint64_t c) {
    const uint64_t chunk_bytes = (c + 1 == n_chunks) ? total_bytes - c * b3::CHUNK_BYTES : b3::CHUNK_BYTES;
    const uint32_t nb = (uint32_t)((chunk_bytes + b3::BLOCK_BYTES - 1) / b3::BLOCK_BYTES);
    if constexpr (BlockElems<FID>::WHOLE) {
        // Interior fast path (a full chunk of a multi-chunk leaf whose rows all exist): no per-row bounds checks, flags
        // and lengths are literals, and two blocks per trip through two buffers used alternately, so that the
        // prefetched block is never copied.  The issue port binds this kernel (one warp instruction per cycle per
        // sub-partition), so every instruction that is not a G function counts.
        using E = typename Field<FID>::E;
        constexpr int L = Field<FID>::LIMBS, WPE = 2 * L, EPB = BlockElems<FID>::EPB, K = BlockElems<FID>::K;
        const int64_t row0 = (int64_t)(c * 16 * EPB) - K;  // row of element 0 of the chunk's first block
        if (n_chunks > 1 && nb == 16 && (uint64_t)(row0 + 16 * EPB) <= n_rows) {
            const uint64_t *base = mat + ((row0 - row_base) * (int64_t)row_stride + (int64_t)col) * L;
            const size_t rs = row_stride * L;
            auto words = [](uint32_t(&m)[16], const E(&e)[EPB]) {
#pragma unroll
                for (int i = 0; i < EPB; i++) {
                    const E r = Field<FID>::to_repr(e[i]);
#pragma unroll
                    for (int l = 0; l < L; l++) {
                        m[i * WPE + 2 * l] = (uint32_t)r.v[l];
                        m[i * WPE + 2 * l + 1] = (uint32_t)(r.v[l] >> 32);
                    }
                }
            };
            E bufA[EPB], bufB[EPB];
            if (c == 0) {  // the 32 zero bytes in front of row 0
#pragma unroll
                for (int i = 0; i < EPB; i++) bufA[i] = i < K ? Field<FID>::zero() : ld_fe<L>(base + (size_t)i * rs);
            } else {
#pragma unroll
                for (int i = 0; i < EPB; i++) bufA[i] = ld_fe<L>(base + (size_t)i * rs);
            }
            b3::set_iv(cv);
#pragma unroll 1
            for (int b = 0; b < 16; b += 2) {
                const uint64_t *pb = base + (size_t)(b + 1) * EPB * rs;
#pragma unroll
                for (int i = 0; i < EPB; i++) bufB[i] = ld_fe<L>(pb + (size_t)i * rs);
                uint32_t m[16];
                words(m, bufA);
                b3::compress(cv, m, c, b3::BLOCK_BYTES, b == 0 ? b3::CHUNK_START : 0u);
                if (b + 2 < 16) {
#pragma unroll
                    for (int i = 0; i < EPB; i++) bufA[i] = ld_fe<L>(pb + (size_t)(EPB + i) * rs);
                }
                words(m, bufB);
                b3::compress(cv, m, c, b3::BLOCK_BYTES, b + 2 == 16 ? b3::CHUNK_END : 0u);
            }
            return;
        }
    }
    b3::set_iv(cv);
    BlockElems<FID> cur, nxt;
    if constexpr (BlockElems<FID>::WHOLE) cur.load(mat, n_rows, row_stride, col, c * 16, row_base);
    for (uint32_t b = 0; b < nb; b++) {
        uint32_t m[16];
        if constexpr (BlockElems<FID>::WHOLE) {
            if (b + 1 < nb) nxt.load(mat, n_rows, row_stride, col, c * 16 + b + 1, row_base);
            cur.words(m);
        } else {
            load_block_words<FID>(m, mat, n_rows, row_stride, col, c * 16 + b, row_base);
        }
        uint32_t flags = (b == 0 ? b3::CHUNK_START : 0u);
        uint32_t len = b3::BLOCK_BYTES;
        if (b + 1 == nb) {
            flags |= b3::CHUNK_END;
            if (n_chunks == 1) flags |= b3::ROOT;
            len = (uint32_t)(chunk_bytes - (uint64_t)b * b3::BLOCK_BYTES);
        }
        b3::compress(cv, m, n_chunks == 1 ? 0 : c, len, flags);
        if constexpr (BlockElems<FID>::WHOLE) cur = nxt;
    }
}

template <int FID, class Store>
__device__ __forceinline__ void hash_chunks_body(const uint64_t *__restrict__ mat, size_t n_rows, size_t row_stride, size_t n_cols,
                                                 const uint64_t *__restrict__ col_idx, uint64_t total_bytes, uint64_t n_chunks,
                                                 int64_t row_base, uint64_t chunk0, uint64_t chunk_end, Store store) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cols) return;
    const size_t col = col_idx ? (size_t)col_idx[j] : j;
    for (uint64_t c = chunk0 + blockIdx.y; c < chunk_end; c += gridDim.y) {
        uint32_t cv[8];
        chunk_cv<FID>(cv, mat, n_rows, row_stride, col, total_bytes, n_chunks, row_base, c);
        uint4 *o = reinterpret_cast<uint4 *>(store(c, j));
        o[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
        o[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
    }
}

template <int FID>
__global__ void __launch_bounds__(128, Field<FID>::LIMBS == 1 ? LCPC_HASH_CTAS1 : 5)
k_hash_chunks(const uint64_t *__restrict__ mat, size_t n_rows, size_t row_stride, size_t n_cols,
              const uint64_t *__restrict__ col_idx, uint64_t total_bytes, uint64_t n_chunks, uint32_t *__restrict__ out,
              int64_t row_base, uint64_t chunk0, uint64_t chunk_end) {
    hash_chunks_body<FID>(mat, n_rows, row_stride, n_cols, col_idx, total_bytes, n_chunks, row_base, chunk0, chunk_end,
                          [=](uint64_t c, size_t j) { return out + (c * n_cols + j) * 8; });
}

// Row-sharded hashing across GPUs, exchange fused into the hash: the chaining value of (chunk c, column j) is stored
// straight into the chaining-value store of the rank that owns column j -- local HBM for this rank's own column block,
// peer HBM over NVLink for the others (32 bytes per chunk and column: 3 % of the encoded matrix for 8-byte elements).
// Rank g owns columns [g << log_cb, (g + 1) << log_cb); its store is [n_chunks][1 << log_cb][32 B].
template <int FID>
__global__ void __launch_bounds__(128, Field<FID>::LIMBS == 1 ? LCPC_HASH_CTAS1 : 5)
k_hash_chunks_scatter(const uint64_t *__restrict__ mat, size_t n_rows, size_t row_stride, size_t n_cols, uint64_t total_bytes,
                      uint64_t n_chunks, int64_t row_base, uint64_t chunk0, uint64_t chunk_end,
                      const __grid_constant__ CvScatter sc) {
    hash_chunks_body<FID>(mat, n_rows, row_stride, n_cols, nullptr, total_bytes, n_chunks, row_base, chunk0, chunk_end,
                          [&](uint64_t c, size_t j) {
                              const size_t in_block = j & (((size_t)1 << sc.log_cb) - 1);
                              return sc.base[j >> sc.log_cb] + ((c << sc.log_cb) + in_block) * 8;
                          });
}

// One thread per column: BLAKE3 tree over the column's chunk chaining values
// (left subtree = largest power of two of chunks strictly below the total).
__global__ void __launch_bounds__(128)
k_hash_merge(const uint32_t *__restrict__ cvs, size_t n_cols, uint64_t n_chunks, uint32_t *__restrict__ leaves) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cols) return;
    uint32_t stack[40][8];
    int sp = 0;
    uint32_t cv[8];
    for (uint64_t c = 0; c < n_chunks; c++) {
        const uint4 *in = reinterpret_cast<const uint4 *>(cvs + (c * n_cols + j) * 8);
        uint4 a = in[0], b = in[1];
        cv[0] = a.x; cv[1] = a.y; cv[2] = a.z; cv[3] = a.w;
        cv[4] = b.x; cv[5] = b.y; cv[6] = b.z; cv[7] = b.w;
        if (c + 1 == n_chunks) break;
        uint64_t total = c + 1;
        while ((total & 1) == 0) {
            sp--;
            b3::parent_cv(stack[sp], cv, 0, cv);
            total >>= 1;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) stack[sp][i] = cv[i];
        sp++;
    }
    while (sp > 0) {
        sp--;
        b3::parent_cv(stack[sp], cv, sp == 0 ? b3::ROOT : 0u, cv);
    }
    uint4 *o = reinterpret_cast<uint4 *>(leaves + j * 8);
    o[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
    o[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
}

// ------------------------------------------------------------------ fused leaf merge + Merkle tree

// BLAKE3 parent tree over the chunk chaining values of column j (n_chunks >= 2), values read through L2 (they may have
// been written by other CTAs of the same launch)
__device__ __forceinline__ void merge_column(uint32_t cv[8], const uint32_t *cvs, size_t n_cols, uint64_t n_chunks, size_t j) {
    uint32_t stack[40][8];
    int sp = 0;
    for (uint64_t c = 0; c < n_chunks; c++) {
        const uint4 *in = reinterpret_cast<const uint4 *>(cvs + (c * n_cols + j) * 8);
        const uint4 a = __ldcg(in), b = __ldcg(in + 1);
        cv[0] = a.x; cv[1] = a.y; cv[2] = a.z; cv[3] = a.w;
        cv[4] = b.x; cv[5] = b.y; cv[6] = b.z; cv[7] = b.w;
        if (c + 1 == n_chunks) break;
        uint64_t total = c + 1;
        while ((total & 1) == 0) {
            sp--;
            b3::parent_cv(stack[sp], cv, 0, cv);
            total >>= 1;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) stack[sp][i] = cv[i];
        sp++;
    }
    while (sp > 0) {
        sp--;
        b3::parent_cv(stack[sp], cv, sp == 0 ? b3::ROOT : 0u, cv);
    }
}

// device-scope release / acquire around the ticket counters (MEMBAR.ALL.GPU instead of __threadfence()'s MEMBAR.SC.GPU)
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// first digest of tree level l in the flat array [np2 | np2/2 | ... | 1]
__device__ __forceinline__ size_t level_offset(size_t np2, int l) { return l == 0 ? 0 : 2 * np2 - (np2 >> (l - 1)); }

// The Merkle levels above one tile of T = blockDim.x adjacent leaves (thread t holds leaf tile*T + t in `leaf`): the
// leaf level and the log2(T) levels inside the tile are written to the flat tree; then the LAST tile of the launch to
// get here (ticket counter, left at zero again) computes the levels above the tile roots.  One launch builds the tree.
// tw: leaves per tile (a power of two <= T; T when 0); the CTA may have more threads than that.
template <int T>
__device__ __forceinline__ void tile_tree(const uint32_t leaf[8], size_t tile, uint8_t *hashes, size_t np2, unsigned *ticket,
                                          unsigned n_tiles, uint32_t (*buf)[T][8], unsigned *s_flag, unsigned tw = 0) {
    const unsigned t = threadIdx.x;
    if (tw == 0) tw = T;
    const size_t tile_n = np2 < (size_t)tw ? np2 : (size_t)tw;
    int lt = 0;
    while (((size_t)1 << lt) < tile_n) lt++;
    if (t < tile_n) {
#pragma unroll
        for (int k = 0; k < 8; k++) buf[0][t][k] = leaf[k];
        uint4 *g = reinterpret_cast<uint4 *>(hashes + (tile * tw + t) * 32);
        g[0] = make_uint4(leaf[0], leaf[1], leaf[2], leaf[3]);
        g[1] = make_uint4(leaf[4], leaf[5], leaf[6], leaf[7]);
    }
    __syncthreads();
    int src = 0;
    for (int l = 1; l <= lt; l++) {
        const size_t n_out = tile_n >> l;
        if (t < n_out) {
            uint32_t o[8];
            b3::hash_pair<true>(buf[src][2 * t], buf[src][2 * t + 1], o);
#pragma unroll
            for (int k = 0; k < 8; k++) buf[src ^ 1][t][k] = o[k];
            uint4 *g = reinterpret_cast<uint4 *>(hashes + (level_offset(np2, l) + tile * n_out + t) * 32);
            g[0] = make_uint4(o[0], o[1], o[2], o[3]);
            g[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
        __syncthreads();
        src ^= 1;
    }
    if (n_tiles <= 1) return;
    // publish this tile's root, then take a ticket: the last tile continues with the top of the tree
    fence_gpu();
    if (t == 0) {
        const unsigned old = atomicAdd(ticket, 1u);
        *s_flag = (old == n_tiles - 1) ? 1u : 0u;
        if (old == n_tiles - 1) *ticket = 0;  // ready for the next launch on this context (stream-ordered)
    }
    __syncthreads();
    if (*s_flag == 0) return;
    fence_gpu();
    int depth = 0;
    while (((size_t)1 << depth) < np2) depth++;
    for (int l = lt + 1; l <= depth; l++) {
        const size_t n_out = np2 >> l;
        const uint4 *in = reinterpret_cast<const uint4 *>(hashes + level_offset(np2, l - 1) * 32);
        uint4 *out = reinterpret_cast<uint4 *>(hashes + level_offset(np2, l) * 32);
        for (size_t i = t; i < n_out; i += blockDim.x) {
            const uint4 a0 = __ldcg(in + 4 * i), a1 = __ldcg(in + 4 * i + 1), b0 = __ldcg(in + 4 * i + 2), b1 = __ldcg(in + 4 * i + 3);
            const uint32_t lft[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const uint32_t rgt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint32_t o[8];
            b3::hash_pair<true>(lft, rgt, o);
            __stcg(out + 2 * i, make_uint4(o[0], o[1], o[2], o[3]));
            __stcg(out + 2 * i + 1, make_uint4(o[4], o[5], o[6], o[7]));
        }
        __syncthreads();  // orders this CTA's own stores before its own loads of the next level: no device-wide fence
    }
}

// Everything after the chunk chaining values in ONE launch: per column the BLAKE3 parent tree over its chunk values
// (-> leaf; with n_chunks == 1 the leaves are already in place), then the whole Merkle tree (merkleize, lib.rs:720-734:
// leaves n_cols..np2 are all-zero).  grid = np2 / T tiles.
constexpr int MT_TILE = 256;
// PAR: the BLAKE3 parent tree over a column's n_chunks chaining values, level by level across the CTA instead of one
// thread walking it: BLAKE3's tree ("the left subtree takes the largest power of two of chunks below the total") is the
// tree obtained by pairing adjacent nodes and carrying an odd last node up unchanged, so level k is
// floor(n_k / 2) independent compressions per column.  A CTA owns only C columns (their chaining values sit in shared
// memory as [chunk][column], two buffers used alternately) and the grid is np2 / C tiles: depth ceil(log2 n_chunks)
// instead of n_chunks - 1, and enough CTAs to fill the machine when few columns meet many chunks -- at 8 GPUs the leaves
// of the weak-scaling bench have 33 chunks and a rank owns 8192 columns: the per-thread walk was a 32-deep chain on 32
// CTAs (0.063 ms); 512 CTAs of 16 columns walk 6 levels.
// cv_stride: columns per chunk row of `cvs` (n_cols, or the column-block width of a row-sharded store).
template <bool PAR>
__global__ void __launch_bounds__(MT_TILE)
k_merge_tree(const uint32_t *cvs, size_t n_cols, size_t cv_stride, uint64_t n_chunks, uint8_t *hashes, size_t np2,
             unsigned *ticket, int C) {
    __shared__ uint32_t buf[2][MT_TILE][8];
    __shared__ unsigned s_flag;
    extern __shared__ uint32_t sm_cv[];  // PAR: [n_chunks][C][8] | [(n_chunks + 1) / 2][C][8]
    const unsigned t = threadIdx.x;
    uint32_t leaf[8];
#pragma unroll
    for (int k = 0; k < 8; k++) leaf[k] = 0;
    if constexpr (PAR) {
        const unsigned nc = (unsigned)n_chunks;
        uint32_t *in = sm_cv, *out = sm_cv + (size_t)nc * C * 8;
        const size_t j0 = (size_t)blockIdx.x * C;
        if (j0 < n_cols) {
            for (unsigned idx = t; idx < nc * C * 2; idx += MT_TILE) {  // 16 bytes per thread per trip, coalesced
                const unsigned c = idx / (2 * C), rem = idx % (2 * C), col = rem >> 1, half = rem & 1;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (j0 + col < n_cols) v = __ldcg(reinterpret_cast<const uint4 *>(cvs + ((size_t)c * cv_stride + j0 + col) * 8) + half);
                *reinterpret_cast<uint4 *>(in + ((size_t)c * C + col) * 8 + half * 4) = v;
            }
            __syncthreads();
            unsigned n = nc;
            while (n > 1) {
                const unsigned pairs = n >> 1;
                const uint32_t flags = n == 2 ? (uint32_t)b3::ROOT : 0u;
                for (unsigned idx = t; idx < pairs * C; idx += MT_TILE) {
                    const unsigned p = idx / C, col = idx % C;
                    uint32_t o[8];
                    b3::parent_cv(in + ((size_t)(2 * p) * C + col) * 8, in + ((size_t)(2 * p + 1) * C + col) * 8, flags, o);
#pragma unroll
                    for (int k = 0; k < 8; k++) out[((size_t)p * C + col) * 8 + k] = o[k];
                }
                if (n & 1) {  // the odd last node moves up unchanged
                    for (unsigned idx = t; idx < (unsigned)C * 8; idx += MT_TILE)
                        out[(size_t)pairs * C * 8 + idx] = in[(size_t)(n - 1) * C * 8 + idx];
                }
                __syncthreads();
                uint32_t *tmp = in;
                in = out;
                out = tmp;
                n = pairs + (n & 1);
            }
            // row 0 of `in` holds the leaves of columns j0 .. j0 + C
            if (t < (unsigned)C && j0 + t < n_cols) {
#pragma unroll
                for (int k = 0; k < 8; k++) leaf[k] = in[(size_t)t * 8 + k];
            }
        }
        tile_tree<MT_TILE>(leaf, blockIdx.x, hashes, np2, ticket, gridDim.x, buf, &s_flag, (unsigned)C);
        return;
    } else {
        const size_t j = (size_t)blockIdx.x * MT_TILE + t;
        if (j < n_cols) {
            if (n_chunks > 1) {
                merge_column(leaf, cvs, cv_stride, n_chunks, j);
            } else {
                const uint4 *in = reinterpret_cast<const uint4 *>(hashes + j * 32);
                const uint4 a = __ldcg(in), b = __ldcg(in + 1);
                leaf[0] = a.x; leaf[1] = a.y; leaf[2] = a.z; leaf[3] = a.w;
                leaf[4] = b.x; leaf[5] = b.y; leaf[6] = b.z; leaf[7] = b.w;
            }
        }
        tile_tree<MT_TILE>(leaf, blockIdx.x, hashes, np2, ticket, gridDim.x, buf, &s_flag);
    }
}

// The whole of merkleize (lib.rs:720-734) in ONE launch: grid = (chunks, tiles of 128 columns over the PADDED leaf range).
// CTA (c, tile) hashes chunk c of its columns; the last of a tile's CTAs to finish (per-tile ticket) merges the chunk
// values into the tile's leaves and builds the tile's 7 tree levels, and the last tile to finish builds the top of the
// tree.  Chunks are the fast grid dimension, so tiles complete one after another while later tiles are still hashing:
// the merge and the lower tree levels hide behind the hashing, and only the top levels remain as a serial tail.
// tickets: 1 + n_tiles zeroed counters, left zeroed.
constexpr int HT_TILE = 128;
template <int FID>
__global__ void __launch_bounds__(HT_TILE, LCPC_HASH_CTAS1)
k_hash_tree(const uint64_t *__restrict__ mat, size_t n_rows, size_t row_stride, size_t n_cols, uint64_t total_bytes,
            uint64_t n_chunks, uint32_t *cvs, uint8_t *hashes, size_t np2, unsigned *tickets, unsigned group) {
    __shared__ uint32_t buf[2][HT_TILE][8];
    __shared__ unsigned s_flag;
    // linear CTA id -> (group of `group` adjacent tiles, chunk, tile inside the group), tile fastest: consecutive CTAs
    // read adjacent 1 KiB segments of the same rows (DRAM pages), and a group's tiles complete together
    const unsigned n_tiles = (unsigned)((np2 + HT_TILE - 1) / HT_TILE);
    const unsigned per_group = group * (unsigned)n_chunks;
    const unsigned g_idx = blockIdx.x / per_group, rem = blockIdx.x % per_group;
    const unsigned g_tiles = (g_idx + 1) * group <= n_tiles ? group : n_tiles - g_idx * group;  // last group may be short
    const uint64_t c = rem / g_tiles;
    const size_t tile = (size_t)g_idx * group + rem % g_tiles;
    if (c >= n_chunks) return;
    const size_t j = tile * HT_TILE + threadIdx.x;
    uint32_t leaf[8];
#pragma unroll
    for (int k = 0; k < 8; k++) leaf[k] = 0;
    if (tile * HT_TILE >= n_cols) {
        // a tile of padding leaves only: nothing to hash, chunk 0's CTA builds its (all-zero-leaf) levels
        if (c != 0) return;
    } else {
        if (j < n_cols) chunk_cv<FID>(leaf, mat, n_rows, row_stride, j, total_bytes, n_chunks, 0, c);
        if (n_chunks > 1) {
            if (j < n_cols) {
                uint4 *o = reinterpret_cast<uint4 *>(cvs + (c * n_cols + j) * 8);
                __stcg(o, make_uint4(leaf[0], leaf[1], leaf[2], leaf[3]));
                __stcg(o + 1, make_uint4(leaf[4], leaf[5], leaf[6], leaf[7]));
            }
            fence_gpu();
            __syncthreads();
            if (threadIdx.x == 0) {
                const unsigned old = atomicAdd(&tickets[1 + tile], 1u);
                s_flag = (old == (unsigned)n_chunks - 1) ? 1u : 0u;
                if (old == (unsigned)n_chunks - 1) tickets[1 + tile] = 0;
            }
            __syncthreads();
            if (s_flag == 0) return;
            fence_gpu();
#pragma unroll
            for (int k = 0; k < 8; k++) leaf[k] = 0;
            if (j < n_cols) merge_column(leaf, cvs, n_cols, n_chunks, j);
            __syncthreads();  // s_flag is reused by tile_tree
        }
    }
    tile_tree<HT_TILE>(leaf, tile, hashes, np2, tickets, (unsigned)((np2 + HT_TILE - 1) / HT_TILE), buf, &s_flag);
}

static uint64_t leaf_bytes(int fid, size_t n_rows) { return 32 + (uint64_t)n_rows * 8 * field_consts(fid).limbs; }
static uint64_t leaf_chunks(int fid, size_t n_rows) {
    return (leaf_bytes(fid, n_rows) + b3::CHUNK_BYTES - 1) / b3::CHUNK_BYTES;
}

size_t hash_scratch_bytes(int fid, size_t n_rows, size_t n_cols) {
    uint64_t nc = leaf_chunks(fid, n_rows);
    return nc <= 1 ? 0 : (size_t)(nc * n_cols * 32);
}

template <int FID>
static cudaError_t hash_columns_t(const uint64_t *d_mat, size_t n_rows, size_t row_stride, size_t n_cols,
                                  const uint64_t *d_col_idx, uint8_t *d_leaves, uint8_t *d_cv_scratch,
                                  const Launch &lc) {
    if (n_cols == 0) return cudaSuccess;
    const uint64_t total = leaf_bytes(FID, n_rows), nc = leaf_chunks(FID, n_rows);
    const unsigned gx = (unsigned)((n_cols + 127) / 128);
    const unsigned gy = (unsigned)(nc < 65535 ? nc : 65535);
    uint32_t *out = nc == 1 ? reinterpret_cast<uint32_t *>(d_leaves) : reinterpret_cast<uint32_t *>(d_cv_scratch);
    lc.begin("k_hash_chunks");
    k_hash_chunks<FID><<<dim3(gx, gy), 128, 0, lc.s>>>(d_mat, n_rows, row_stride, n_cols, d_col_idx, total, nc, out, 0, 0, nc);
    lc.end();
    if (nc > 1) {
        lc.begin("k_hash_merge");
        k_hash_merge<<<gx, 128, 0, lc.s>>>(reinterpret_cast<const uint32_t *>(d_cv_scratch), n_cols, nc,
                                        reinterpret_cast<uint32_t *>(d_leaves));
        lc.end();
    }
    return cudaGetLastError();
}

template <int FID>
static cudaError_t hash_chunk_range_t(const uint64_t *d_mat, int64_t row_base, size_t n_rows_valid, size_t row_stride,
                                      size_t n_cols, uint64_t chunk0, uint64_t chunk_end, uint64_t total_bytes,
                                      uint64_t n_chunks_total, uint8_t *d_cvs, const Launch &lc) {
    if (n_cols == 0 || chunk_end <= chunk0) return cudaSuccess;
    const unsigned gx = (unsigned)((n_cols + 127) / 128);
    const uint64_t todo = chunk_end - chunk0;
    const unsigned gy = (unsigned)(todo < 65535 ? todo : 65535);
    lc.begin("k_hash_chunks");
    k_hash_chunks<FID><<<dim3(gx, gy), 128, 0, lc.s>>>(d_mat, n_rows_valid, row_stride, n_cols, nullptr, total_bytes,
                                                    n_chunks_total, reinterpret_cast<uint32_t *>(d_cvs), row_base, chunk0,
                                                    chunk_end);
    lc.end();
    return cudaGetLastError();
}

cudaError_t hash_chunk_range(int fid, const uint64_t *d_mat, int64_t row_base, size_t n_rows_valid, size_t row_stride,
                             size_t n_cols, uint64_t chunk0, uint64_t chunk_end, uint64_t total_bytes,
                             uint64_t n_chunks_total, uint8_t *d_cvs, const Launch &lc) {
#define LCPC_HCR(F) hash_chunk_range_t<F>(d_mat, row_base, n_rows_valid, row_stride, n_cols, chunk0, chunk_end, total_bytes, n_chunks_total, d_cvs, lc)
    switch (fid) {
    case FT63: return LCPC_HCR(FT63);
    case FT127: return LCPC_HCR(FT127);
    case FT191: return LCPC_HCR(FT191);
    case FT255: return LCPC_HCR(FT255);
    case FT253_192: return LCPC_HCR(FT253_192);
    default: return cudaErrorInvalidValue;
    }
#undef LCPC_HCR
}

cudaError_t hash_chunk_range_scatter(int fid, const uint64_t *d_mat, int64_t row_base, size_t n_rows_valid, size_t row_stride,
                                     size_t n_cols, uint64_t chunk0, uint64_t chunk_end, uint64_t total_bytes,
                                     uint64_t n_chunks_total, const CvScatter &sc, const Launch &lc) {
    if (n_cols == 0 || chunk_end <= chunk0) return cudaSuccess;
    const unsigned gx = (unsigned)((n_cols + 127) / 128);
    const uint64_t todo = chunk_end - chunk0;
    const unsigned gy = (unsigned)(todo < 65535 ? todo : 65535);
    lc.begin("k_hash_chunks_scatter");
#define LCPC_HCS(F) k_hash_chunks_scatter<F><<<dim3(gx, gy), 128, 0, lc.s>>>(d_mat, n_rows_valid, row_stride, n_cols, total_bytes, n_chunks_total, row_base, chunk0, chunk_end, sc)
    switch (fid) {
    case FT63: LCPC_HCS(FT63); break;
    case FT127: LCPC_HCS(FT127); break;
    case FT255: LCPC_HCS(FT255); break;
    case FT253_192: LCPC_HCS(FT253_192); break;
    default: return cudaErrorInvalidValue;  // 24-byte elements straddle chunk boundaries: no row-sharded hashing
    }
#undef LCPC_HCS
    lc.end();
    return cudaGetLastError();
}

cudaError_t hash_merge(const uint8_t *d_cvs, size_t n_cols, uint64_t n_chunks, uint8_t *d_leaves, const Launch &lc) {
    if (n_cols == 0) return cudaSuccess;
    if (n_chunks <= 1) {  // a single chunk is its own (ROOT-flagged) leaf
        if (d_cvs != d_leaves) return cudaMemcpyAsync(d_leaves, d_cvs, n_cols * 32, cudaMemcpyDeviceToDevice, lc.s);
        return cudaSuccess;
    }
    lc.begin("k_hash_merge");
    k_hash_merge<<<(unsigned)((n_cols + 127) / 128), 128, 0, lc.s>>>(reinterpret_cast<const uint32_t *>(d_cvs), n_cols, n_chunks,
                                                                  reinterpret_cast<uint32_t *>(d_leaves));
    lc.end();
    return cudaGetLastError();
}

// cvs (n_chunks >= 2) or leaves already in d_hashes (n_chunks == 1) -> leaves + the whole tree, one launch
cudaError_t merge_tree(const uint8_t *d_cvs, size_t n_cols, uint64_t n_chunks, uint8_t *d_hashes, size_t np2, unsigned *d_ticket,
                       const Launch &lc, size_t cv_stride) {
    if (np2 == 0 || n_cols > np2) return cudaErrorInvalidValue;
    if (cv_stride == 0) cv_stride = n_cols;
    const unsigned tiles = (unsigned)((np2 + MT_TILE - 1) / MT_TILE);
    // Level-wise merge (PAR) or the per-thread walk?  Both do the same n_chunks - 1 compressions per column; the walk keeps
    // every thread busy but is a chain of n_chunks - 1 dependent compressions on n_cols / 256 CTAs, the level-wise form is
    // ceil(log2 n_chunks) deep on n_cols / C CTAs with a third of its threads busy on average.  Measured
    // (tools/bench_merge_tree.py, profiles/r02_hash_tail.md): 33 chunks x 8192 columns 54 -> 33 us, 147 x 8192 191 -> 95 us,
    // 17 x 16384 a tie, and from 32768 columns up the walk wins by 1.5 - 2 x.  So: level-wise when a leaf has at least 8
    // chunks and there are at most 16384 columns (a rank's column block of a sharded commitment, the proof-of-storage
    // shapes); C = the largest power of two <= 32 whose two buffers (1.5 x n_chunks rows of C digests) fit 30 KiB beside
    // the 16 KiB of static tile buffers.
    int C = 0;
    if (n_chunks >= 8 && n_cols <= 16384) {
        for (C = 32; C >= 4; C >>= 1)
            if ((n_chunks + (n_chunks + 1) / 2) * (uint64_t)C * 32 <= 30 * 1024 && (size_t)C <= np2) break;
        if (C < 4) C = 0;
    }
    if (const char *e = getenv("LCPC_MERGE_PAR")) {  // experiments: 0 forces the per-thread walk, n >= 4 that many columns per CTA
        const int v = atoi(e);
        if (v == 0) C = 0;
        else if (n_chunks >= 2 && (n_chunks + (n_chunks + 1) / 2) * (uint64_t)v * 32 <= 30 * 1024 && (size_t)v <= np2) C = v;
    }
    lc.begin("k_merge_tree");
    if (C > 0) {
        const size_t smem = (size_t)(n_chunks + (n_chunks + 1) / 2) * C * 32;
        k_merge_tree<true><<<(unsigned)(np2 / C), MT_TILE, smem, lc.s>>>(reinterpret_cast<const uint32_t *>(d_cvs), n_cols, cv_stride, n_chunks,
                                                          d_hashes, np2, d_ticket, C);
    } else {
        k_merge_tree<false><<<tiles, MT_TILE, 0, lc.s>>>(reinterpret_cast<const uint32_t *>(d_cvs), n_cols, cv_stride, n_chunks,
                                                         d_hashes, np2, d_ticket, 0);
    }
    lc.end();
    return cudaGetLastError();
}

size_t hash_tree_tickets(size_t np2) { return 1 + (np2 + HT_TILE - 1) / HT_TILE; }

bool hash_tree_supported(int fid, size_t n_rows, size_t np2) {
    return leaf_chunks(fid, n_rows) * ((np2 + HT_TILE - 1) / HT_TILE + 64) < ((uint64_t)1 << 31);
}

// One launch pays when the whole grid is resident at once (small commitments: launch- and latency-bound); with several
// waves of CTAs the per-tile tails run on a few warps while their CTA slots could hash, and two launches win
// (2^24 Ft63: 0.175 ms against 0.188 ms; profiles/r02_hash_tail.md).
bool hash_tree_preferred(int fid, size_t n_rows, size_t np2) {
    return hash_tree_supported(fid, n_rows, np2) && leaf_chunks(fid, n_rows) * ((np2 + HT_TILE - 1) / HT_TILE) <= 148 * 7;
}

template <int FID>
static cudaError_t hash_tree_t(const uint64_t *d_mat, size_t n_rows, size_t row_stride, size_t n_cols, size_t np2,
                               uint8_t *d_hashes, uint8_t *d_cvs, unsigned *d_tickets, const Launch &lc) {
    const uint64_t total = leaf_bytes(FID, n_rows), nc = leaf_chunks(FID, n_rows);
    const unsigned n_tiles = (unsigned)((np2 + HT_TILE - 1) / HT_TILE);
    // tile-fastest over the whole grid measured best (profiles/r02_hash_tail.md); LCPC_HT_GROUP overrides for experiments
    unsigned group = n_tiles;
    if (const char *e = getenv("LCPC_HT_GROUP")) group = (unsigned)atoi(e);
    if (group == 0 || group > n_tiles) group = n_tiles;
    const unsigned n_groups = (n_tiles + group - 1) / group;
    lc.begin("k_hash_tree");
    k_hash_tree<FID><<<n_groups * group * (unsigned)nc, HT_TILE, 0, lc.s>>>(d_mat, n_rows, row_stride, n_cols, total, nc,
                                                                         reinterpret_cast<uint32_t *>(d_cvs), d_hashes, np2,
                                                                         d_tickets, group);
    lc.end();
    return cudaGetLastError();
}

cudaError_t hash_tree(int fid, const uint64_t *d_mat, size_t n_rows, size_t row_stride, size_t n_cols, size_t np2,
                      uint8_t *d_hashes, uint8_t *d_cvs, unsigned *d_tickets, const Launch &lc) {
    if (n_cols == 0 || np2 == 0 || n_cols > np2 || !hash_tree_supported(fid, n_rows, np2)) return cudaErrorInvalidValue;
    switch (fid) {
    case FT63: return hash_tree_t<FT63>(d_mat, n_rows, row_stride, n_cols, np2, d_hashes, d_cvs, d_tickets, lc);
    case FT127: return hash_tree_t<FT127>(d_mat, n_rows, row_stride, n_cols, np2, d_hashes, d_cvs, d_tickets, lc);
    case FT191: return hash_tree_t<FT191>(d_mat, n_rows, row_stride, n_cols, np2, d_hashes, d_cvs, d_tickets, lc);
    case FT255: return hash_tree_t<FT255>(d_mat, n_rows, row_stride, n_cols, np2, d_hashes, d_cvs, d_tickets, lc);
    case FT253_192: return hash_tree_t<FT253_192>(d_mat, n_rows, row_stride, n_cols, np2, d_hashes, d_cvs, d_tickets, lc);
    default: return cudaErrorInvalidValue;
    }
}

uint64_t hash_leaf_bytes(int fid, size_t n_rows) { return leaf_bytes(fid, n_rows); }
uint64_t hash_leaf_chunks(int fid, size_t n_rows) { return leaf_chunks(fid, n_rows); }

// ------------------------------------------------------------------ column-major emit

// out[c][r] = canonical little-endian repr of mat[r][c]: the layout of proof-of-storage's encoded file
// (lcpc_online/encoded_file_writer.rs:327-349: column c at byte c*row_capacity*w, row r at +r*w;
// fields/data_field.rs:62-70: to_repr bytes).  32 x 32 element tiles through shared memory so that both the
// row-major loads and the column-major stores are full lines.
template <int FID>
__global__ void __launch_bounds__(256) k_emit_colmajor(const uint64_t *__restrict__ mat, size_t n_rows, size_t row_stride,
                                                       size_t n_cols, uint64_t *__restrict__ out, size_t out_col_stride) {
    using F = Field<FID>;
    constexpr int L = F::LIMBS;
    __shared__ uint64_t tile[L][32][33];
    const size_t c0 = (size_t)blockIdx.x * 32, r0 = (size_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const size_t r = r0 + ty + 8 * k, c = c0 + tx;
        typename F::E v = F::zero();
        if (r < n_rows && c < n_cols) v = F::to_repr(ld_fe<L>(mat + (r * row_stride + c) * L));
#pragma unroll
        for (int l = 0; l < L; l++) tile[l][ty + 8 * k][tx] = v.v[l];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const size_t c = c0 + ty + 8 * k, r = r0 + tx;
        if (r < n_rows && c < n_cols) {
#pragma unroll
            for (int l = 0; l < L; l++) out[(c * out_col_stride + r) * L + l] = tile[l][tx][ty + 8 * k];
        }
    }
}

cudaError_t emit_colmajor(int fid, const uint64_t *d_mat, size_t n_rows, size_t row_stride, size_t n_cols, uint64_t *d_out,
                          size_t out_col_stride, const Launch &lc) {
    if (n_rows == 0 || n_cols == 0) return cudaSuccess;
    const dim3 grid((unsigned)((n_cols + 31) / 32), (unsigned)((n_rows + 31) / 32));
    lc.begin("k_emit_colmajor");
    switch (fid) {
    case FT63: k_emit_colmajor<FT63><<<grid, 256, 0, lc.s>>>(d_mat, n_rows, row_stride, n_cols, d_out, out_col_stride); break;
    case FT127: k_emit_colmajor<FT127><<<grid, 256, 0, lc.s>>>(d_mat, n_rows, row_stride, n_cols, d_out, out_col_stride); break;
    case FT191: k_emit_colmajor<FT191><<<grid, 256, 0, lc.s>>>(d_mat, n_rows, row_stride, n_cols, d_out, out_col_stride); break;
    case FT255: k_emit_colmajor<FT255><<<grid, 256, 0, lc.s>>>(d_mat, n_rows, row_stride, n_cols, d_out, out_col_stride); break;
    case FT253_192: k_emit_colmajor<FT253_192><<<grid, 256, 0, lc.s>>>(d_mat, n_rows, row_stride, n_cols, d_out, out_col_stride); break;
    default: return cudaErrorInvalidValue;
    }
    lc.end();
    return cudaGetLastError();
}

cudaError_t hash_columns(int fid, const uint64_t *d_mat, size_t n_rows, size_t row_stride, size_t n_cols,
                         const uint64_t *d_col_idx, uint8_t *d_leaves, uint8_t *d_cv_scratch, const Launch &lc) {
    switch (fid) {
    case FT63: return hash_columns_t<FT63>(d_mat, n_rows, row_stride, n_cols, d_col_idx, d_leaves, d_cv_scratch, lc);
    case FT127: return hash_columns_t<FT127>(d_mat, n_rows, row_stride, n_cols, d_col_idx, d_leaves, d_cv_scratch, lc);
    case FT191: return hash_columns_t<FT191>(d_mat, n_rows, row_stride, n_cols, d_col_idx, d_leaves, d_cv_scratch, lc);
    case FT255: return hash_columns_t<FT255>(d_mat, n_rows, row_stride, n_cols, d_col_idx, d_leaves, d_cv_scratch, lc);
    case FT253_192: return hash_columns_t<FT253_192>(d_mat, n_rows, row_stride, n_cols, d_col_idx, d_leaves, d_cv_scratch, lc);
    default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------ Merkle tree

constexpr int MERKLE_TILE = 512;  // input digests per CTA (256 threads)

// Reduces up to 9 levels of the tree over a tile of `MERKLE_TILE` inputs in shared memory.
// level_in points at a full level of n_in digests; the following levels are contiguous
// after it in the flat array (lib.rs:784-789 split_at_mut layout).
__global__ void __launch_bounds__(256) k_merkle_levels(uint8_t *level_in, size_t n_in, int n_levels) {
    __shared__ uint32_t buf[2][MERKLE_TILE][8];
    const size_t tile0 = (size_t)blockIdx.x * MERKLE_TILE;
    const size_t tile_n = n_in - tile0 < MERKLE_TILE ? n_in - tile0 : MERKLE_TILE;
    const uint4 *in = reinterpret_cast<const uint4 *>(level_in + tile0 * 32);
    for (size_t i = threadIdx.x; i < tile_n * 2; i += blockDim.x) reinterpret_cast<uint4 *>(&buf[0][0][0])[i] = in[i];
    __syncthreads();
    uint8_t *level_out = level_in + n_in * 32;
    size_t level_n = n_in / 2;   // digests in the output level (whole tree)
    size_t cur = tile_n;         // digests in this tile at the current level
    size_t out0 = tile0 / 2;
    int src = 0;
    for (int l = 0; l < n_levels; l++) {
        const size_t n_out = cur / 2;
        for (size_t i = threadIdx.x; i < n_out; i += blockDim.x) {
            uint32_t o[8];
            b3::hash_pair(buf[src][2 * i], buf[src][2 * i + 1], o);
#pragma unroll
            for (int k = 0; k < 8; k++) buf[src ^ 1][i][k] = o[k];
            uint4 *g = reinterpret_cast<uint4 *>(level_out + (out0 + i) * 32);
            g[0] = make_uint4(o[0], o[1], o[2], o[3]);
            g[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
        __syncthreads();
        src ^= 1;
        cur = n_out;
        level_out += level_n * 32;
        level_n /= 2;
        out0 /= 2;
    }
}

cudaError_t merkle_tree(uint8_t *d_hashes, size_t n_leaves, const Launch &lc) {
    uint8_t *level = d_hashes;
    size_t n_in = n_leaves;
    while (n_in > 1) {
        int levels = 0;
        size_t tile = n_in < (size_t)MERKLE_TILE ? n_in : (size_t)MERKLE_TILE;
        while (((size_t)1 << levels) < tile) levels++;
        unsigned blocks = (unsigned)((n_in + MERKLE_TILE - 1) / MERKLE_TILE);
        lc.begin("k_merkle_levels");
        k_merkle_levels<<<blocks, 256, 0, lc.s>>>(level, n_in, levels);
        lc.end();
        for (int l = 0; l < levels; l++) {
            level += n_in * 32;
            n_in /= 2;
        }
    }
    return cudaGetLastError();
}

// paths[i][l] = level_l[(col >> l) ^ 1]   (open_column, lib.rs:836-851)
__global__ void k_gather_paths(const uint8_t *__restrict__ hashes, size_t np2, int depth, const uint64_t *__restrict__ cols,
                               size_t n, uint8_t *__restrict__ paths) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * (size_t)depth) return;
    const size_t i = t / depth;
    const int l = (int)(t % depth);
    // offset of level l: sum_{k<l} np2 >> k = 2*np2 - (np2 >> (l-1)) for l >= 1
    const size_t off = l == 0 ? 0 : 2 * np2 - (np2 >> (l - 1));
    const size_t node = ((size_t)cols[i] >> l) ^ 1;
    const uint4 *srcp = reinterpret_cast<const uint4 *>(hashes + (off + node) * 32);
    uint4 *dstp = reinterpret_cast<uint4 *>(paths + t * 32);
    dstp[0] = srcp[0];
    dstp[1] = srcp[1];
}

cudaError_t gather_paths(const uint8_t *d_hashes, size_t np2, const uint64_t *d_cols, size_t n, uint8_t *d_paths,
                         const Launch &lc) {
    int depth = 0;
    while (((size_t)1 << depth) < np2) depth++;
    if (n == 0 || depth == 0) return cudaSuccess;
    size_t total = n * (size_t)depth;
    lc.begin("k_gather_paths");
    k_gather_paths<<<(unsigned)((total + 255) / 256), 256, 0, lc.s>>>(d_hashes, np2, depth, d_cols, n, d_paths);
    lc.end();
    return cudaGetLastError();
}

// verify_column_path's climb (lib.rs:999-1011): from each leaf digest up its sibling path;
// ok[i] = 1 when the recomputed root equals `root`.
__global__ void k_verify_paths(const uint8_t *__restrict__ leaves, const uint8_t *__restrict__ paths, int depth,
                               const uint64_t *__restrict__ cols, size_t n, const uint8_t *__restrict__ root,
                               uint32_t *__restrict__ ok) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h[8], sib[8];
    const uint32_t *lp = reinterpret_cast<const uint32_t *>(leaves + i * 32);
#pragma unroll
    for (int k = 0; k < 8; k++) h[k] = lp[k];
    uint64_t col = cols[i];
    for (int l = 0; l < depth; l++) {
        const uint32_t *sp = reinterpret_cast<const uint32_t *>(paths + (i * (size_t)depth + l) * 32);
#pragma unroll
        for (int k = 0; k < 8; k++) sib[k] = sp[k];
        uint32_t o[8];
        if ((col & 1) == 0) b3::hash_pair(h, sib, o);
        else b3::hash_pair(sib, h, o);
#pragma unroll
        for (int k = 0; k < 8; k++) h[k] = o[k];
        col >>= 1;
    }
    const uint32_t *rp = reinterpret_cast<const uint32_t *>(root);
    uint32_t diff = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) diff |= h[k] ^ rp[k];
    ok[i] = diff == 0 ? 1u : 0u;
}

cudaError_t verify_paths(const uint8_t *d_leaves, const uint8_t *d_paths, int depth, const uint64_t *d_cols, size_t n,
                         const uint8_t *d_root, uint32_t *d_ok, const Launch &lc) {
    if (n == 0) return cudaSuccess;
    lc.begin("k_verify_paths");
    k_verify_paths<<<(unsigned)((n + 63) / 64), 64, 0, lc.s>>>(d_leaves, d_paths, depth, d_cols, n, d_root, d_ok);
    lc.end();
    return cudaGetLastError();
}

}  // namespace lcpc
