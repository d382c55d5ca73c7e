// Ligero Reed-Solomon encoder: batched multi-row NTT over the lcpc prime fields (kernels and per-field host templates;
// instantiated once per field in lcpc_ntt_f*.cu so that the four fields compile in parallel).
//
// Computes what the reference's LigeroEncodingRho::encode does per row
// (lcpc-ligero-pc/src/lib.rs:162-164 -> fffft fft_io_pc): the length-n decimation-in-
// frequency transform, in-order input, bit-reversed output,
//     out[bitrev(i)] = sum_j in[j] * w^(i*j),   w = ROOT_OF_UNITY^(2^(S-k)), n = 2^k,
// on zero-padded rows (lcpc-2d/src/lib.rs:665-682: the copy of the n_per_row message
// into the n_cols-wide row is fused into the first pass's loads).
//
// Structure (B200): the in-place DIF is cut into passes.  A pass that covers r bits
// is a radix-2^r transform held in registers (constant small twiddles from the kernel
// parameter bank), followed by one multiplication per element by a per-pass twiddle
// table T[m][lo] = w_sub^(lo*bitrev(m)) laid out so that a warp reads it coalesced.
// Leading passes run over global memory with stride n/2^r ("strided" passes: adjacent
// threads own adjacent columns, so every load/store is a full line); the trailing
// 2^LB-point transform of each contiguous block runs out of shared memory
// (limb planes, padded 1-in-16 against bank conflicts) in register-radix sub-steps.
// Because an in-place DIF leaves element i holding X[bitrev(i)], no permutation pass
// is needed for fffft's "io" order.
#pragma once
#ifndef LCPC_NTT_FULL_BLOCK_PATH
#define LCPC_NTT_FULL_BLOCK_PATH 1
#endif
#include "lcpc_field.cuh"
#include "lcpc_kernels.h"

#ifndef LCPC_NTT_CTAS1
#define LCPC_NTT_CTAS1 3  // resident 256-thread CTAs per SM for the one-limb field (register budget 80)
#endif

namespace lcpc {

template <int FID>
struct SmallTw {
    typename Field<FID>::E w[8];
};

__device__ __forceinline__ unsigned bitrev_bits(unsigned x, int bits) {
    return bits == 0 ? 0u : (__brev(x) >> (32 - bits));
}

// ------------------------------------------------------------------ plan-time kernels

// inverse transform constants: w^-1 = w^(n-1) and n^-1 = (2^-1)^log_n, 2^-1 = (p + 1) / 2 brought into Montgomery form
template <int FID>
__global__ void k_init_inverse(const uint64_t *w_in, int log_n, uint64_t *winv_out, uint64_t *ninv_out) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const E w = ld_fe<L>(w_in);
    st_fe<L>(winv_out, F::pow(w, ((uint64_t)1 << log_n) - 1));
    E half, r2;  // (p + 1) / 2: p is odd, so (p >> 1) + 1
    uint64_t carry = 0;
#pragma unroll
    for (int i = L - 1; i >= 0; i--) {
        const uint64_t pi = F::P(i);
        half.v[i] = (pi >> 1) | (carry << 63);
        carry = pi & 1;
    }
    uint64_t c = 1;
#pragma unroll
    for (int i = 0; i < L; i++) {
        const uint64_t v = half.v[i] + c;
        c = v < c ? 1 : 0;
        half.v[i] = v;
    }
#pragma unroll
    for (int i = 0; i < L; i++) r2.v[i] = field_consts(FID).r2[i];
    const E half_m = F::mul(half, r2);
    E acc = F::one();
    for (int i = 0; i < log_n; i++) acc = F::mul(acc, half_m);
    st_fe<L>(ninv_out, acc);
}

template <int FID>
__global__ void k_init_root(const uint64_t *root_in, int log_n, uint64_t *w_out, uint64_t *stw_out) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    E w;
    if (root_in != nullptr) {
        w = ld_fe<L>(root_in);
    } else {
#pragma unroll
        for (int i = 0; i < L; i++) w.v[i] = field_consts(FID).root[i];
        for (int i = 0; i < F::TWO_ADICITY - log_n; i++) w = F::mul(w, w);
    }
    st_fe<L>(w_out, w);
    const uint64_t n = 1ull << log_n;
    for (uint64_t e = 0; e < 8; e++) {
        E v = F::one();
        if ((e * n) % 16 == 0) v = F::pow(w, e * n / 16);
        st_fe<L>(stw_out + e * L, v);
    }
}

// T[m][lo] = w^(mult * lo * bitrev_R(m)), m < 2^R, lo < 2^log_n2
template <int FID>
__global__ void k_build_twiddles(uint64_t *out, const uint64_t *w_n, uint64_t mult, int R, int log_n2) {
    using F = Field<FID>;
    constexpr int L = F::LIMBS;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)1 << (R + log_n2);
    if (idx >= total) return;
    uint64_t m = idx >> log_n2, lo = idx & (((size_t)1 << log_n2) - 1);
    uint64_t e = mult * lo * bitrev_bits((unsigned)m, R);
    st_fe<L>(out + idx * L, F::pow(ld_fe<L>(w_n), e));
}

// ------------------------------------------------------------------ register radix

// 2^R-point DIF on registers; x[j] ends up holding the output of index bitrev_R(j).
// ZB > 0: the inputs x[m], m >= 2^(R - ZB), are known to be zero (the zero padding of a rate-2^-ZB Reed-Solomon
// row in the first pass).  In the first ZB stages every butterfly then has a zero lower input: the sum is the
// upper input itself and the difference is the upper input, so both the addition and the subtraction disappear.
template <int FID, int R, int ZB = 0>
__device__ __forceinline__ void radix_dif(typename Field<FID>::E (&x)[1 << R], const SmallTw<FID> &tw) {
    using F = Field<FID>;
    using E = typename F::E;
#pragma unroll
    for (int t = 0; t < R; t++) {
        const int gap = 1 << (R - 1 - t);
#pragma unroll
        for (int j = 0; j < (1 << R); j++) {
            if ((j & gap) == 0) {
                const int e16 = ((j & (gap - 1)) << t) << (4 - R);  // exponent of w16
                if (t < ZB) {
                    x[j + gap] = (e16 == 0) ? x[j] : F::mul(x[j], tw.w[e16]);
                } else {
                    E a = x[j], b = x[j + gap];
                    x[j] = F::add(a, b);
                    x[j + gap] = (e16 == 0) ? F::sub(a, b) : F::mul(F::sub_for_mul(a, b), tw.w[e16]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------ strided pass

// LN2 >= 0: log_sub - R is that literal (the last strided pass always leaves blocks of 2^LBMAX elements), so every
// element and twiddle address is one base pointer plus a literal offset.  ZB > 0: first pass of a rate-2^-ZB code,
// src_valid == n >> ZB: the zero half (three quarters) is neither loaded nor bounds-checked, see radix_dif.
template <int FID, int R, int LN2 = -1, int ZB = 0>
__global__ void __launch_bounds__(256, Field<FID>::LIMBS == 1 ? LCPC_NTT_CTAS1 : 2)
k_ntt_strided(const uint64_t *src, size_t src_stride, size_t src_valid, uint64_t *dst, size_t n, size_t n_rows,
              int log_sub, const uint64_t *__restrict__ tw, const __grid_constant__ SmallTw<FID> stw) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    const int log_n2 = LN2 >= 0 ? LN2 : log_sub - R;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (n >> R)) return;
    const size_t hi = g >> log_n2, lo = g & (((size_t)1 << log_n2) - 1);
    const size_t base = (hi << log_sub) + lo;
    for (size_t row = blockIdx.y; row < n_rows; row += gridDim.y) {
        E x[1 << R];
        const uint64_t *in = src + (row * src_stride + base) * L;
#pragma unroll
        for (int m = 0; m < (1 << R); m++) {
            if constexpr (ZB > 0) {
                x[m] = m < (1 << (R - ZB)) ? ld_fe<L>(in + ((size_t)m << log_n2) * L) : F::zero();
            } else {
                x[m] = base + ((size_t)m << log_n2) < src_valid ? ld_fe<L>(in + ((size_t)m << log_n2) * L) : F::zero();
            }
        }
        radix_dif<FID, R, ZB>(x, stw);
        const uint64_t *twp = tw + lo * L;
        if constexpr (L == 1) {
            E t[1 << R];
#pragma unroll
            for (int m = 1; m < (1 << R); m++) t[m] = ld_fe<L>(twp + ((size_t)m << log_n2) * L);
#pragma unroll
            for (int m = 1; m < (1 << R); m++) x[m] = F::mul(x[m], t[m]);
        } else {
#pragma unroll
            for (int m = 1; m < (1 << R); m++) x[m] = F::mul(x[m], ld_fe<L>(twp + ((size_t)m << log_n2) * L));
        }
        uint64_t *out = dst + (row * n + base) * L;
#pragma unroll
        for (int m = 0; m < (1 << R); m++) st_fe<L>(out + ((size_t)m << log_n2) * L, x[m]);
    }
}

// ------------------------------------------------------------------ two strided passes in one trip

// For rows longer than 16 * 4096 the leading RA + RB bits take two strided passes, i.e. two full trips through HBM;
// at 2^28 coefficients those two kernels are bandwidth-bound (4.5 TB/s) while everything else is bound by the integer
// pipes.  This kernel does both passes on a tile that stays in shared memory: the CTA owns all 2^(RA+RB) elements at
// stride 4096 for 32 adjacent columns, runs pass A (radix 2^RA over the top bits, table twA) into shared memory and
// pass B (radix 2^RB over the next bits, table twB) out of it.  Same arithmetic, same tables, half the HBM traffic.
// One-limb field, first passes of the transform (log_sub = log_n), log_n - RA - RB = 12.
template <int FID, int RA, int RB, int ZB>
__global__ void __launch_bounds__(256, 3)
k_ntt_strided_fused(const uint64_t *src, size_t src_stride, size_t src_valid, uint64_t *dst, size_t n, size_t n_rows,
                    const uint64_t *__restrict__ twA, const uint64_t *__restrict__ twB,
                    const __grid_constant__ SmallTw<FID> stw) {
    using F = Field<FID>;
    using E = typename F::E;
    static_assert(F::LIMBS == 1, "one-limb field only");
    constexpr int R = RA + RB, LN2 = 12, C = 32;
    __shared__ uint64_t tile[1 << R][C];
    const unsigned c = threadIdx.x & (C - 1), u = threadIdx.x >> 5;  // column inside the tile, phase-specific index
    const size_t col = (size_t)blockIdx.x * C + c;                   // lo2: position inside the final 4096-block
    for (size_t row = blockIdx.y; row < n_rows; row += gridDim.y) {
        // pass A: thread (sB = u, c) owns the 2^RA elements s = a * 2^RB + sB; lo1 = sB * 4096 + col
        for (unsigned sB = u; sB < (1u << RB); sB += blockDim.x / C) {
            const size_t lo1 = ((size_t)sB << LN2) + col;
            const uint64_t *in = src + (row * src_stride + lo1);
            E x[1 << RA];
#pragma unroll
            for (int a = 0; a < (1 << RA); a++) {
                const size_t off = (size_t)a << (LN2 + RB);
                if constexpr (ZB > 0) x[a].v[0] = a < (1 << (RA - ZB)) ? in[off] : 0;
                else x[a].v[0] = lo1 + off < src_valid ? in[off] : 0;
            }
            radix_dif<FID, RA, ZB>(x, stw);
            E t[1 << RA];
#pragma unroll
            for (int a = 1; a < (1 << RA); a++) t[a].v[0] = twA[((size_t)a << (LN2 + RB)) + lo1];
#pragma unroll
            for (int a = 1; a < (1 << RA); a++) x[a] = F::mul(x[a], t[a]);
#pragma unroll
            for (int a = 0; a < (1 << RA); a++) tile[(a << RB) + sB][c] = x[a].v[0];
        }
        __syncthreads();
        // pass B: thread (a = u, c) owns the 2^RB elements of sub-block a; lo2 = col
        for (unsigned a = u; a < (1u << RA); a += blockDim.x / C) {
            E x[1 << RB];
#pragma unroll
            for (int b = 0; b < (1 << RB); b++) x[b].v[0] = tile[(a << RB) + b][c];
            radix_dif<FID, RB, 0>(x, stw);
            E t[1 << RB];
#pragma unroll
            for (int b = 1; b < (1 << RB); b++) t[b].v[0] = twB[((size_t)b << LN2) + col];
#pragma unroll
            for (int b = 1; b < (1 << RB); b++) x[b] = F::mul(x[b], t[b]);
            uint64_t *out = dst + (row * n + ((size_t)a << (LN2 + RB)) + col);
#pragma unroll
            for (int b = 0; b < (1 << RB); b++) out[(size_t)b << LN2] = x[b].v[0];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ shared-memory block pass

// shared-memory indices are 32-bit on purpose: 64-bit index arithmetic doubles the SHF/IADD3 count on the
// ALU pipe, which is the pipe that bounds these kernels
__device__ __forceinline__ unsigned sm_phys(unsigned i) { return i + (i >> 4); }

// GSRC: the sub-step takes its inputs straight from global memory (the first sub-step of a block: adjacent
// threads own adjacent elements, so the loads are full lines and the staging copy into shared memory, one
// STS + LDS per element and a barrier, is skipped); elements at or beyond `valid` read as zero.
// LN2 >= 0 fixes log_sub - R at compile time (the 4096-point block of the one-limb field): every shared-memory and
// twiddle address is then one base plus a literal offset, instead of five integer instructions per access (shift,
// add, two LEAs for the 1-in-16 padding and the byte address) -- 14 % of the executed instructions, mostly on the
// ALU pipe that bounds the kernel.
// FULL (with GSRC): every element of the block exists (the block pass behind a strided pass: valid == block size), so the
// loads are unconditional -- no index, compare and zero default per element.
template <int FID, int R, bool TW, bool GSRC, int LN2 = -1, bool FULL = false>
__device__ __forceinline__ void block_substep(uint64_t *sm, unsigned plane, int LB, int log_sub,
                                              const uint64_t *__restrict__ tw, const SmallTw<FID> &stw,
                                              const uint64_t *__restrict__ gsrc, unsigned valid) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    const int log_n2 = LN2 >= 0 ? LN2 : log_sub - R;
    // padded index of element base + (m << log_n2), given p0 = sm_phys(base)
    auto phys = [&](unsigned base, unsigned p0, int m) -> unsigned {
        if constexpr (LN2 >= 4) return p0 + (unsigned)m * ((1u << LN2) + (1u << (LN2 - 4)));  // m << LN2 is a multiple of 16
        else if constexpr (LN2 == 0) return p0 + (unsigned)m;                                  // base is a multiple of 2^R = 16
        else return sm_phys(base + ((unsigned)m << log_n2));
    };
    const unsigned groups = (1u << LB) >> R;
    for (unsigned g = threadIdx.x; g < groups; g += blockDim.x) {
        const unsigned hi = g >> log_n2, lo = g & ((1u << log_n2) - 1);
        const unsigned base = (hi << log_sub) + lo;
        const unsigned p0 = sm_phys(base);
        E x[1 << R];
#pragma unroll
        for (int m = 0; m < (1 << R); m++) {
            const unsigned i = base + ((unsigned)m << log_n2);
            if constexpr (GSRC && FULL) {
                x[m] = ld_fe<L>(gsrc + (size_t)i * L);
            } else if constexpr (GSRC) {
                x[m] = i < valid ? ld_fe<L>(gsrc + (size_t)i * L) : F::zero();
            } else {
                const unsigned p = phys(base, p0, m);
#pragma unroll
                for (int l = 0; l < L; l++) x[m].v[l] = sm[l * plane + p];
            }
        }
        radix_dif<FID, R>(x, stw);
        if constexpr (TW) {
            if constexpr (L == 1) {
                // all pass twiddles of this group in flight together (one exposed L1/L2 latency, not 2^R - 1)
                E t[1 << R];
#pragma unroll
                for (int m = 1; m < (1 << R); m++) t[m] = ld_fe<L>(tw + (((unsigned)m << log_n2) + lo) * L);
#pragma unroll
                for (int m = 1; m < (1 << R); m++) x[m] = F::mul(x[m], t[m]);
            } else {
                // wide elements: keep one twiddle live at a time (register pressure decides occupancy here)
#pragma unroll
                for (int m = 1; m < (1 << R); m++) x[m] = F::mul(x[m], ld_fe<L>(tw + (size_t)(((unsigned)m << log_n2) + lo) * L));
            }
        }
#pragma unroll
        for (int m = 0; m < (1 << R); m++) {
            const unsigned p = phys(base, p0, m);
#pragma unroll
            for (int l = 0; l < L; l++) sm[l * plane + p] = x[m].v[l];
        }
    }
}

template <int FID, int R, bool GSRC>
__device__ __forceinline__ void block_substep_any(uint64_t *sm, unsigned plane, int LB, int log_sub,
                                                  const uint64_t *__restrict__ tw, const SmallTw<FID> &stw,
                                                  const uint64_t *__restrict__ gsrc, unsigned valid) {
    if (log_sub - R > 0) block_substep<FID, R, true, GSRC>(sm, plane, LB, log_sub, tw, stw, gsrc, valid);
    else block_substep<FID, R, false, GSRC>(sm, plane, LB, log_sub, tw, stw, gsrc, valid);
}

template <int FID, int RMAX, bool GSRC>
__device__ __forceinline__ void block_substep_r(int R, uint64_t *sm, unsigned plane, int LB, int log_sub,
                                                const uint64_t *__restrict__ tw, const SmallTw<FID> &stw,
                                                const uint64_t *__restrict__ gsrc, unsigned valid) {
    if (R == 4) {
        if constexpr (RMAX >= 4) block_substep_any<FID, 4, GSRC>(sm, plane, LB, log_sub, tw, stw, gsrc, valid);
    } else if (R == 3) {
        block_substep_any<FID, 3, GSRC>(sm, plane, LB, log_sub, tw, stw, gsrc, valid);
    } else if (R == 2) {
        block_substep_any<FID, 2, GSRC>(sm, plane, LB, log_sub, tw, stw, gsrc, valid);
    } else {
        block_substep_any<FID, 1, GSRC>(sm, plane, LB, log_sub, tw, stw, gsrc, valid);
    }
}

template <int FID, int RMAX, bool SCATTER>
__global__ void __launch_bounds__(256, Field<FID>::LIMBS == 1 ? LCPC_NTT_CTAS1 : 2)
k_ntt_block(const uint64_t *src, size_t src_stride, size_t src_valid, uint64_t *dst, size_t n, size_t n_rows, int LB,
            const uint64_t *__restrict__ tw, const __grid_constant__ SmallTw<FID> stw,
            const __grid_constant__ ScatterDst sc) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    extern __shared__ uint64_t sm[];
    const unsigned NB = 1u << LB;
    const unsigned plane = NB + (NB >> 4) + 1;
    const size_t col0 = (size_t)blockIdx.x << LB;
    const unsigned valid = src_valid > col0 ? (unsigned)(src_valid - col0 < NB ? src_valid - col0 : NB) : 0u;
    for (size_t row = blockIdx.y; row < n_rows; row += gridDim.y) {
        if constexpr (L == 1 && RMAX == 4) {
            if (LB == 12) {  // the full-size block: three radix-16 sub-steps with literal strides 256, 16, 1
#if LCPC_NTT_FULL_BLOCK_PATH
                if (valid == NB) block_substep<FID, 4, true, true, 8, true>(sm, plane, 12, 12, tw, stw, src + (row * src_stride + col0) * L, valid);
                else
#endif
                block_substep<FID, 4, true, true, 8>(sm, plane, 12, 12, tw, stw, src + (row * src_stride + col0) * L, valid);
                __syncthreads();
                block_substep<FID, 4, true, false, 4>(sm, plane, 12, 8, tw + ((size_t)1 << 12) * L, stw, nullptr, 0u);
                __syncthreads();
                block_substep<FID, 4, false, false, 0>(sm, plane, 12, 4, nullptr, stw, nullptr, 0u);
                __syncthreads();
            }
        }
        int log_sub = (L == 1 && RMAX == 4 && LB == 12) ? 0 : LB;
        size_t tw_off = 0;
        bool first = true;
        while (log_sub > 0) {
            const int R = log_sub < RMAX ? log_sub : RMAX;
            const uint64_t *t = tw + tw_off * L;
            if (first) block_substep_r<FID, RMAX, true>(R, sm, plane, LB, log_sub, t, stw, src + (row * src_stride + col0) * L, valid);
            else block_substep_r<FID, RMAX, false>(R, sm, plane, LB, log_sub, t, stw, nullptr, 0u);
            first = false;
            if (log_sub - R > 0) tw_off += (size_t)1 << log_sub;
            log_sub -= R;
            __syncthreads();
        }
        for (unsigned i = threadIdx.x; i < NB; i += blockDim.x) {
            E v;
            const unsigned p = sm_phys(i);
#pragma unroll
            for (int l = 0; l < L; l++) v.v[l] = sm[l * plane + p];
            if constexpr (SCATTER) {
                // whole block lands in one rank's column-block matrix (1 << log_cb >= 1 << LB)
                uint64_t *base = sc.base[col0 >> sc.log_cb];
                const size_t off = ((sc.row0 + row) << sc.log_cb) + (col0 & (((size_t)1 << sc.log_cb) - 1)) + i;
                st_fe<L>(base + off * L, v);
            } else {
                st_fe<L>(dst + (row * n + col0 + i) * L, v);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ inverse transform (fffft ifft_oi)

// Undoes radix_dif: x[j] holds the entry of index bitrev_R(j) on entry, natural order on exit; the factor 1/2 of every
// butterfly is left out (the caller scales by n^-1 once).  twi = powers of w16^-1.
template <int FID, int R>
__device__ __forceinline__ void radix_dit_inv(typename Field<FID>::E (&x)[1 << R], const SmallTw<FID> &twi) {
    using F = Field<FID>;
    using E = typename F::E;
#pragma unroll
    for (int t = R - 1; t >= 0; t--) {
        const int gap = 1 << (R - 1 - t);
#pragma unroll
        for (int j = 0; j < (1 << R); j++) {
            if ((j & gap) == 0) {
                const int e16 = ((j & (gap - 1)) << t) << (4 - R);
                const E a = x[j];
                const E b = (e16 == 0) ? x[j + gap] : F::mul(x[j + gap], twi.w[e16]);
                x[j] = F::add(a, b);
                x[j + gap] = F::sub(a, b);
            }
        }
    }
}

// Inverse of one strided pass, in place: divide by the pass twiddles (the table built from w^-1), then the inverse
// register radix.  SCALE: this is the last inverse pass, multiply by n^-1 on the way out.
template <int FID, int R>
__global__ void __launch_bounds__(256, 2)
k_intt_strided(uint64_t *data, size_t n, size_t n_rows, int log_sub, const uint64_t *__restrict__ twi,
               const __grid_constant__ SmallTw<FID> stwi, const __grid_constant__ typename Field<FID>::E ninv, int scale) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    const int log_n2 = log_sub - R;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (n >> R)) return;
    const size_t hi = g >> log_n2, lo = g & (((size_t)1 << log_n2) - 1);
    const size_t base = (hi << log_sub) + lo;
    for (size_t row = blockIdx.y; row < n_rows; row += gridDim.y) {
        E x[1 << R];
        uint64_t *io = data + (row * n + base) * L;
#pragma unroll
        for (int m = 0; m < (1 << R); m++) x[m] = ld_fe<L>(io + ((size_t)m << log_n2) * L);
        if (log_n2 > 0) {
#pragma unroll
            for (int m = 1; m < (1 << R); m++) x[m] = F::mul(x[m], ld_fe<L>(twi + (((size_t)m << log_n2) + lo) * L));
        }
        radix_dit_inv<FID, R>(x, stwi);
        if (scale) {
#pragma unroll
            for (int m = 0; m < (1 << R); m++) x[m] = F::mul(x[m], ninv);
        }
#pragma unroll
        for (int m = 0; m < (1 << R); m++) st_fe<L>(io + ((size_t)m << log_n2) * L, x[m]);
    }
}

template <int FID, int R>
__device__ __forceinline__ void block_substep_inv(uint64_t *sm, unsigned plane, int LB, int log_sub,
                                                  const uint64_t *__restrict__ twi, const SmallTw<FID> &stwi) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    const int log_n2 = log_sub - R;
    const unsigned groups = (1u << LB) >> R;
    for (unsigned g = threadIdx.x; g < groups; g += blockDim.x) {
        const unsigned hi = g >> log_n2, lo = g & ((1u << log_n2) - 1);
        const unsigned base = (hi << log_sub) + lo;
        E x[1 << R];
#pragma unroll
        for (int m = 0; m < (1 << R); m++) {
            const unsigned p = sm_phys(base + ((unsigned)m << log_n2));
#pragma unroll
            for (int l = 0; l < L; l++) x[m].v[l] = sm[l * plane + p];
        }
        if (log_n2 > 0) {
#pragma unroll
            for (int m = 1; m < (1 << R); m++) x[m] = F::mul(x[m], ld_fe<L>(twi + (size_t)(((unsigned)m << log_n2) + lo) * L));
        }
        radix_dit_inv<FID, R>(x, stwi);
#pragma unroll
        for (int m = 0; m < (1 << R); m++) {
            const unsigned p = sm_phys(base + ((unsigned)m << log_n2));
#pragma unroll
            for (int l = 0; l < L; l++) sm[l * plane + p] = x[m].v[l];
        }
    }
}

// Inverse of the block pass: the sub-steps of k_ntt_block in reverse order on a block held in shared memory.
template <int FID, int RMAX>
__global__ void __launch_bounds__(256, 2)
k_intt_block(uint64_t *data, size_t n, size_t n_rows, int LB, const uint64_t *__restrict__ twi,
             const __grid_constant__ SmallTw<FID> stwi, const __grid_constant__ typename Field<FID>::E ninv, int scale) {
    using F = Field<FID>;
    using E = typename F::E;
    constexpr int L = F::LIMBS;
    extern __shared__ uint64_t sm[];
    const unsigned NB = 1u << LB;
    const unsigned plane = NB + (NB >> 4) + 1;
    const size_t col0 = (size_t)blockIdx.x << LB;
    // forward sub-steps (log_sub, R, table offset), replayed backwards
    int ls_list[16], r_list[16];
    size_t off_list[16];
    int n_steps = 0;
    {
        int log_sub = LB;
        size_t off = 0;
        while (log_sub > 0) {
            const int R = log_sub < RMAX ? log_sub : RMAX;
            ls_list[n_steps] = log_sub;
            r_list[n_steps] = R;
            off_list[n_steps] = off;
            n_steps++;
            if (log_sub - R > 0) off += (size_t)1 << log_sub;
            log_sub -= R;
        }
    }
    for (size_t row = blockIdx.y; row < n_rows; row += gridDim.y) {
        uint64_t *io = data + (row * n + col0) * L;
        for (unsigned i = threadIdx.x; i < NB; i += blockDim.x) {
            const E v = ld_fe<L>(io + (size_t)i * L);
            const unsigned p = sm_phys(i);
#pragma unroll
            for (int l = 0; l < L; l++) sm[l * plane + p] = v.v[l];
        }
        __syncthreads();
        for (int k = n_steps - 1; k >= 0; k--) {
            const uint64_t *t = twi + off_list[k] * L;
            switch (r_list[k]) {
            case 4:
                if constexpr (RMAX >= 4) block_substep_inv<FID, 4>(sm, plane, LB, ls_list[k], t, stwi);
                break;
            case 3: block_substep_inv<FID, 3>(sm, plane, LB, ls_list[k], t, stwi); break;
            case 2: block_substep_inv<FID, 2>(sm, plane, LB, ls_list[k], t, stwi); break;
            default: block_substep_inv<FID, 1>(sm, plane, LB, ls_list[k], t, stwi); break;
            }
            __syncthreads();
        }
        for (unsigned i = threadIdx.x; i < NB; i += blockDim.x) {
            E v;
            const unsigned p = sm_phys(i);
#pragma unroll
            for (int l = 0; l < L; l++) v.v[l] = sm[l * plane + p];
            if (scale) v = F::mul(v, ninv);
            st_fe<L>(io + (size_t)i * L, v);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ host side

static int block_bits_max(int limbs) { return limbs == 1 ? 12 : (limbs == 2 ? 11 : 10); }
static int radix_max(int limbs) { return limbs <= 2 ? 4 : 3; }

template <int FID>
cudaError_t plan_build_t(NttPlan &plan, int log_n, const uint64_t *root_mont, const Launch &lc) {
    using F = Field<FID>;
    constexpr int L = F::LIMBS;
    plan.fid = FID;
    plan.log_n = log_n;
    plan.n = (size_t)1 << log_n;
    plan.passes.clear();
    const int rmax = radix_max(L);
    const int lb = log_n < block_bits_max(L) ? log_n : block_bits_max(L);
    // leading strided passes
    int rem = log_n - lb;
    int n_strided = (rem + rmax - 1) / rmax;
    size_t tw_elems = 0;
    int log_sub = log_n;
    for (int i = 0; i < n_strided; i++) {
        int bits = rem / (n_strided - i) + ((rem % (n_strided - i)) ? 1 : 0);
        plan.passes.push_back({0, bits, log_sub, tw_elems});
        tw_elems += (size_t)1 << log_sub;
        log_sub -= bits;
        rem -= bits;
    }
    // trailing block pass and its sub-step tables
    if (lb > 0) {
        plan.passes.push_back({1, lb, lb, tw_elems});
        int ls = lb;
        while (ls > 0) {
            int r = ls < rmax ? ls : rmax;
            if (ls - r > 0) tw_elems += (size_t)1 << ls;
            ls -= r;
        }
    }
    plan.tw_elems = tw_elems;
    cudaError_t e;
    uint64_t *d_consts = nullptr;  // [w_n | stw(8)] | optional root_in | [w_n^-1 | stw_inv(8)] | n^-1
    if ((e = cudaMalloc(&d_consts, (size_t)(1 + 8 + 1 + 1 + 8 + 1) * L * sizeof(uint64_t))) != cudaSuccess) return e;
    uint64_t *d_root_in = nullptr;
    if (root_mont != nullptr) {
        d_root_in = d_consts + 9 * L;
        if ((e = cudaMemcpyAsync(d_root_in, root_mont, L * sizeof(uint64_t), cudaMemcpyHostToDevice, lc.s)) != cudaSuccess) {
            cudaFree(d_consts);
            return e;
        }
    }
    lc.begin("k_init_root");
    k_init_root<FID><<<1, 1, 0, lc.s>>>(d_root_in, log_n, d_consts, d_consts + L);
    lc.end();
    // the inverse transform (ifft_oi) uses the same tables built from w^-1, stored behind the forward ones
    uint64_t *d_inv = d_consts + 10 * L;  // [w^-1 | stw_inv(8)], then n^-1 at d_consts + 19 L
    lc.begin("k_init_inverse");
    k_init_inverse<FID><<<1, 1, 0, lc.s>>>(d_consts, log_n, d_inv, d_consts + 19 * L);
    lc.end();
    lc.begin("k_init_root");
    k_init_root<FID><<<1, 1, 0, lc.s>>>(d_inv, log_n, d_inv, d_inv + L);
    lc.end();
    if (tw_elems > 0) {
        if ((e = cudaMalloc(&plan.d_tw, 2 * tw_elems * L * sizeof(uint64_t))) != cudaSuccess) {
            cudaFree(d_consts);
            return e;
        }
        plan.d_tw_inv = plan.d_tw + tw_elems * L;
    }
    auto build = [&](size_t off, int ls, int r) {
        size_t total = (size_t)1 << ls;
        unsigned blocks = (unsigned)((total + 255) / 256);
        lc.begin("k_build_twiddles");
        k_build_twiddles<FID><<<blocks, 256, 0, lc.s>>>(plan.d_tw + off * L, d_consts, (uint64_t)1 << (log_n - ls), r, ls - r);
        lc.end();
        lc.begin("k_build_twiddles");
        k_build_twiddles<FID><<<blocks, 256, 0, lc.s>>>(plan.d_tw_inv + off * L, d_inv, (uint64_t)1 << (log_n - ls), r, ls - r);
        lc.end();
    };
    for (const NttPass &p : plan.passes) {
        if (p.kind == 0) {
            build(p.tw_off, p.log_sub, p.bits);
        } else {
            int ls = p.bits;
            size_t off = p.tw_off;
            while (ls > 0) {
                int r = ls < rmax ? ls : rmax;
                if (ls - r > 0) {
                    build(off, ls, r);
                    off += (size_t)1 << ls;
                }
                ls -= r;
            }
        }
    }
    uint64_t h_stw[8 * L], h_inv[10 * L];
    e = cudaMemcpyAsync(h_stw, d_consts + L, sizeof h_stw, cudaMemcpyDeviceToHost, lc.s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_inv, d_inv + L, sizeof h_inv - L * sizeof(uint64_t), cudaMemcpyDeviceToHost, lc.s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(lc.s);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d_consts);
    if (e != cudaSuccess) return e;
    for (int i = 0; i < 8; i++)
        for (int l = 0; l < MAX_LIMBS; l++) {
            plan.stw.w[i][l] = l < L ? h_stw[i * L + l] : 0;
            plan.stw_inv.w[i][l] = l < L ? h_inv[i * L + l] : 0;
        }
    for (int l = 0; l < MAX_LIMBS; l++) plan.ninv[l] = l < L ? h_inv[8 * L + l] : 0;  // n^-1 sits right behind stw_inv
    return cudaSuccess;
}

template <int FID>
cudaError_t encode_t(const NttPlan &plan, const uint64_t *src, size_t src_stride, size_t src_valid,
                            uint64_t *dst, size_t n_rows, const Launch &lc, const ScatterDst *scatter) {
    using F = Field<FID>;
    constexpr int L = F::LIMBS;
    constexpr int RMAX = L <= 2 ? 4 : 3;
    if (n_rows == 0) return cudaSuccess;
    SmallTw<FID> stw;
    for (int i = 0; i < 8; i++)
        for (int l = 0; l < L; l++) stw.w[i].v[l] = plan.stw.w[i][l];
    const size_t n = plan.n;
    const unsigned gy = (unsigned)(n_rows < 65535 ? n_rows : 65535);
    if (plan.log_n == 0) {  // length-1 transform: a copy
        if (src != dst || src_stride != n)
            return cudaMemcpy2DAsync(dst, n * L * 8, src, src_stride * L * 8, L * 8, n_rows, cudaMemcpyDeviceToDevice, lc.s);
        return cudaSuccess;
    }
    bool first = true;
    size_t skip = 0;
    if constexpr (L == 1) {
        // two leading strided passes that end on 4096-blocks: one trip through HBM instead of two
        if (plan.passes.size() >= 3 && plan.passes[0].kind == 0 && plan.passes[1].kind == 0 && plan.passes[2].kind == 1 &&
            plan.passes[1].log_sub - plan.passes[1].bits == 12 && plan.passes[0].bits == 3 &&
            (plan.passes[1].bits == 2 || plan.passes[1].bits == 3)) {
            const NttPass &pa = plan.passes[0], &pb = plan.passes[1];
            const uint64_t *twA = plan.d_tw + pa.tw_off, *twB = plan.d_tw + pb.tw_off;
            const bool zb = (src_valid << 1) == n;
            dim3 grid((unsigned)(4096 / 32), gy);
            lc.begin("k_ntt_strided_fused");
            if (pb.bits == 3) {
                if (zb) k_ntt_strided_fused<FID, 3, 3, 1><<<grid, 256, 0, lc.s>>>(src, src_stride, src_valid, dst, n, n_rows, twA, twB, stw);
                else k_ntt_strided_fused<FID, 3, 3, 0><<<grid, 256, 0, lc.s>>>(src, src_stride, src_valid, dst, n, n_rows, twA, twB, stw);
            } else {
                if (zb) k_ntt_strided_fused<FID, 3, 2, 1><<<grid, 256, 0, lc.s>>>(src, src_stride, src_valid, dst, n, n_rows, twA, twB, stw);
                else k_ntt_strided_fused<FID, 3, 2, 0><<<grid, 256, 0, lc.s>>>(src, src_stride, src_valid, dst, n, n_rows, twA, twB, stw);
            }
            lc.end();
            first = false;
            skip = 2;
        }
    }
    for (const NttPass &p : plan.passes) {
        if (skip) {
            skip--;
            continue;
        }
        const uint64_t *in = first ? src : dst;
        const size_t in_stride = first ? src_stride : n;
        const size_t in_valid = first ? src_valid : n;
        const uint64_t *tw = plan.d_tw + p.tw_off * L;
        if (p.kind == 0) {
            const size_t groups = n >> p.bits;
            dim3 grid((unsigned)((groups + 255) / 256), gy);
            lc.begin("k_ntt_strided");
            // specialised instances for the one-limb field: literal stride when this pass leaves full blocks behind, and
            // zero-aware first stage(s) when it is the first pass over a rate-1/2 or rate-1/4 row
            const int lbmax = block_bits_max(L);
            int zb = 0;
            if (first && in_valid < n) {
                if ((in_valid << 1) == n) zb = 1;
                else if ((in_valid << 2) == n) zb = 2;
            }
            if (zb > p.bits) zb = 0;
            const bool lit = L == 1 && p.log_sub - p.bits == lbmax;
#define LCPC_STRIDED(RR)                                                                                                       \
    do {                                                                                                                       \
        if constexpr (L == 1) {                                                                                                \
            if (lit && zb == 1) { k_ntt_strided<FID, RR, 12, 1><<<grid, 256, 0, lc.s>>>(in, in_stride, in_valid, dst, n, n_rows, p.log_sub, tw, stw); break; } \
            if (lit && zb == 0) { k_ntt_strided<FID, RR, 12, 0><<<grid, 256, 0, lc.s>>>(in, in_stride, in_valid, dst, n, n_rows, p.log_sub, tw, stw); break; } \
            if (!lit && zb == 1) { k_ntt_strided<FID, RR, -1, 1><<<grid, 256, 0, lc.s>>>(in, in_stride, in_valid, dst, n, n_rows, p.log_sub, tw, stw); break; } \
        }                                                                                                                      \
        k_ntt_strided<FID, RR><<<grid, 256, 0, lc.s>>>(in, in_stride, in_valid, dst, n, n_rows, p.log_sub, tw, stw);            \
    } while (0)
            switch (p.bits) {
            case 1: LCPC_STRIDED(1); break;
            case 2: LCPC_STRIDED(2); break;
            case 3: LCPC_STRIDED(3); break;
            default:
                if constexpr (RMAX >= 4) LCPC_STRIDED(4);
                break;
            }
#undef LCPC_STRIDED
        } else {
            const int LB = p.bits;
            const size_t NB = (size_t)1 << LB;
            const size_t plane = NB + (NB >> 4) + 1;
            const size_t smem = plane * L * sizeof(uint64_t);
            size_t thr = NB >> RMAX;
            thr = thr < 32 ? 32 : (thr > 256 ? 256 : thr);
            dim3 grid((unsigned)(n >> LB), gy);
            if (scatter) {
                if (LB > scatter->log_cb) return cudaErrorInvalidValue;
                lc.begin("k_ntt_block_scatter");
                k_ntt_block<FID, RMAX, true><<<grid, (unsigned)thr, smem, lc.s>>>(in, in_stride, in_valid, dst, n, n_rows, LB, tw, stw, *scatter);
            } else {
                lc.begin("k_ntt_block");
                k_ntt_block<FID, RMAX, false><<<grid, (unsigned)thr, smem, lc.s>>>(in, in_stride, in_valid, dst, n, n_rows, LB, tw, stw, ScatterDst{});
            }
        }
        lc.end();
        first = false;
    }
    return cudaGetLastError();
}

// n_rows independent inverse transforms in place (rows of stride n): bit-reversed input, in-order output, scaled by
// 1/n -- fffft's ifft_oi, so decode(encode(x)) == x.  The forward passes are undone last to first.
template <int FID>
cudaError_t decode_t(const NttPlan &plan, uint64_t *data, size_t n_rows, const Launch &lc) {
    using F = Field<FID>;
    constexpr int L = F::LIMBS;
    constexpr int RMAX = L <= 2 ? 4 : 3;
    if (n_rows == 0 || plan.log_n == 0) return cudaSuccess;
    SmallTw<FID> stwi;
    typename F::E ninv;
    for (int i = 0; i < 8; i++)
        for (int l = 0; l < L; l++) stwi.w[i].v[l] = plan.stw_inv.w[i][l];
    for (int l = 0; l < L; l++) ninv.v[l] = plan.ninv[l];
    const size_t n = plan.n;
    const unsigned gy = (unsigned)(n_rows < 65535 ? n_rows : 65535);
    for (size_t k = plan.passes.size(); k-- > 0;) {
        const NttPass &p = plan.passes[k];
        const uint64_t *twi = plan.d_tw_inv + p.tw_off * L;
        const int scale = k == 0 ? 1 : 0;
        if (p.kind == 0) {
            const size_t groups = n >> p.bits;
            dim3 grid((unsigned)((groups + 255) / 256), gy);
            lc.begin("k_intt_strided");
            switch (p.bits) {
            case 1: k_intt_strided<FID, 1><<<grid, 256, 0, lc.s>>>(data, n, n_rows, p.log_sub, twi, stwi, ninv, scale); break;
            case 2: k_intt_strided<FID, 2><<<grid, 256, 0, lc.s>>>(data, n, n_rows, p.log_sub, twi, stwi, ninv, scale); break;
            case 3: k_intt_strided<FID, 3><<<grid, 256, 0, lc.s>>>(data, n, n_rows, p.log_sub, twi, stwi, ninv, scale); break;
            default:
                if constexpr (RMAX >= 4) k_intt_strided<FID, 4><<<grid, 256, 0, lc.s>>>(data, n, n_rows, p.log_sub, twi, stwi, ninv, scale);
                break;
            }
        } else {
            const int LB = p.bits;
            const size_t NB = (size_t)1 << LB;
            const size_t plane = NB + (NB >> 4) + 1;
            const size_t smem = plane * L * sizeof(uint64_t);
            size_t thr = NB >> RMAX;
            thr = thr < 32 ? 32 : (thr > 256 ? 256 : thr);
            dim3 grid((unsigned)(n >> LB), gy);
            lc.begin("k_intt_block");
            k_intt_block<FID, RMAX><<<grid, (unsigned)thr, smem, lc.s>>>(data, n, n_rows, LB, twi, stwi, ninv, scale);
        }
        lc.end();
    }
    return cudaGetLastError();
}

}  // namespace lcpc
