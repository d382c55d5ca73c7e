// Multi-limb Montgomery arithmetic on 32-bit limbs with explicit PTX carry chains, for the
// 127/191/255-bit fields (LIMBS = 2, 3, 4 -> N = 4, 6, 8 words of 32 bits).
//
// Coarsely integrated operand scanning: for each word b[i], t += a*b[i] as two carry chains (the
// low halves land on t[j], the high halves on t[j+1]), then one reduction step t += m*p with
// m = -t[0] (every lcpc modulus is 1 mod 2^32, so -p^-1 mod 2^32 = 0xffffffff and the quotient digit
// is a negation), then a one-word shift.  Each instruction is its own asm statement; the carry flag
// lives between consecutive statements, which nvcc keeps in program order (asm volatile).
#pragma once
#include <cstdint>

namespace lcpc {
namespace m32 {

__device__ __forceinline__ void mad_lo_cc(uint32_t &acc, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(b));
}
__device__ __forceinline__ void madc_lo_cc(uint32_t &acc, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(b));
}
__device__ __forceinline__ void mad_hi_cc(uint32_t &acc, uint32_t a, uint32_t b) {
    asm volatile("mad.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(b));
}
__device__ __forceinline__ void madc_hi_cc(uint32_t &acc, uint32_t a, uint32_t b) {
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(b));
}
__device__ __forceinline__ void addc_cc(uint32_t &acc, uint32_t v) {
    asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(acc) : "r"(v));
}
__device__ __forceinline__ void addc(uint32_t &acc, uint32_t v) {
    asm volatile("addc.u32 %0, %0, %1;" : "+r"(acc) : "r"(v));
}
__device__ __forceinline__ void add_cc(uint32_t &acc, uint32_t v) {
    asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(acc) : "r"(v));
}
__device__ __forceinline__ void sub_cc(uint32_t &acc, uint32_t v) {
    asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(acc) : "r"(v));
}
__device__ __forceinline__ void subc_cc(uint32_t &acc, uint32_t v) {
    asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(acc) : "r"(v));
}
__device__ __forceinline__ void subc(uint32_t &acc, uint32_t v) {
    asm volatile("subc.u32 %0, %0, %1;" : "+r"(acc) : "r"(v));
}

// r = a - p if a >= p else a   (a < 2p, N words); PW(i) = word i of p
template <int N, class PW>
__device__ __forceinline__ void cond_sub_p(uint32_t (&a)[N], PW pw) {
    uint32_t t[N];
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = a[i];
    sub_cc(t[0], pw(0));
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(t[i], pw(i));
    uint32_t borrow = 0;
    subc(borrow, 0);  // 0xffffffff when a < p
#pragma unroll
    for (int i = 0; i < N; i++) a[i] = borrow ? a[i] : t[i];
}

template <int N, class PW>
__device__ __forceinline__ void add_mod(uint32_t (&r)[N], const uint32_t (&a)[N], const uint32_t (&b)[N], PW pw) {
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = a[i];
    add_cc(r[0], b[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r[i], b[i]);
    addc(r[N - 1], b[N - 1]);  // no carry out: 2p < 2^(32N)
    cond_sub_p<N>(r, pw);
}

template <int N, class PW>
__device__ __forceinline__ void sub_mod(uint32_t (&r)[N], const uint32_t (&a)[N], const uint32_t (&b)[N], PW pw) {
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = a[i];
    sub_cc(r[0], b[0]);
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(r[i], b[i]);
    uint32_t borrow = 0;
    subc(borrow, 0);  // 0xffffffff when a < b
    // add p back under the mask
    add_cc(r[0], pw(0) & borrow);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r[i], pw(i) & borrow);
    addc(r[N - 1], pw(N - 1) & borrow);
}

// acc += x * y[0..N) * 2^(32*i), split by the parity of the absolute word index so that every
// (low, high) product pair lands on an aligned word pair of its accumulator: ptxas then emits ONE
// IMAD.WIDE.U32.X (64-bit multiply-accumulate with carry in/out) per 32x32 product instead of an
// IMAD + IMAD.HI + two carry adds.  X takes the pairs that start on even words, Y those on odd words.
template <int N, int I, class YW>
__device__ __forceinline__ void row_mad(uint32_t (&X)[2 * N + 2], uint32_t (&Y)[2 * N + 2], uint32_t x, YW yw) {
    {
        constexpr int j0 = I & 1;  // I + j even
        mad_lo_cc(X[I + j0], x, yw(j0));
        madc_hi_cc(X[I + j0 + 1], x, yw(j0));
#pragma unroll
        for (int j = j0 + 2; j < N; j += 2) {
            madc_lo_cc(X[I + j], x, yw(j));
            madc_hi_cc(X[I + j + 1], x, yw(j));
        }
        addc(X[I + j0 + ((N - j0 + 1) / 2) * 2], 0);  // that word only ever collects carries: no overflow
    }
    {
        constexpr int j0 = 1 - (I & 1);  // I + j odd
        mad_lo_cc(Y[I + j0], x, yw(j0));
        madc_hi_cc(Y[I + j0 + 1], x, yw(j0));
#pragma unroll
        for (int j = j0 + 2; j < N; j += 2) {
            madc_lo_cc(Y[I + j], x, yw(j));
            madc_hi_cc(Y[I + j + 1], x, yw(j));
        }
        addc(Y[I + j0 + ((N - j0 + 1) / 2) * 2], 0);
    }
}

template <int N, int I, class PW>
struct MontRows {
    __device__ __forceinline__ static void product(uint32_t (&X)[2 * N + 2], uint32_t (&Y)[2 * N + 2],
                                                   const uint32_t (&a)[N], const uint32_t (&b)[N]) {
        if constexpr (I < N) {
            row_mad<N, I>(X, Y, b[I], [&](int j) { return a[j]; });
            MontRows<N, I + 1, PW>::product(X, Y, a, b);
        }
    }
    // Montgomery digits: word I of T = X + Y (+ carry c of the words already cleared) is cancelled by
    // adding m*p*2^(32*I) with m = -word (p = 1 mod 2^32); the cleared word is then exactly 0 or 2^32.
    __device__ __forceinline__ static void reduce(uint32_t (&X)[2 * N + 2], uint32_t (&Y)[2 * N + 2], uint32_t &c, PW pw) {
        if constexpr (I < N) {
            const uint32_t m = 0u - (X[I] + Y[I] + c);
            row_mad<N, I>(X, Y, m, pw);
            c = (X[I] | Y[I] | c) != 0 ? 1u : 0u;
            MontRows<N, I + 1, PW>::reduce(X, Y, c, pw);
        }
    }
};

// r = a*b*2^(-32N) mod p, a, b < p
template <int N, class PW>
__device__ __forceinline__ void mont_mul(uint32_t (&r)[N], const uint32_t (&a)[N], const uint32_t (&b)[N], PW pw) {
    uint32_t X[2 * N + 2], Y[2 * N + 2];
#pragma unroll
    for (int i = 0; i < 2 * N + 2; i++) X[i] = Y[i] = 0;
    MontRows<N, 0, PW>::product(X, Y, a, b);
    uint32_t c = 0;
    MontRows<N, 0, PW>::reduce(X, Y, c, pw);
    // result = (X + Y) >> 32N, plus the carry of the last cleared word; < 2p so it fits N words
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = X[N + i];
    add_cc(r[0], Y[N]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r[i], Y[N + i]);
    addc(r[N - 1], Y[2 * N - 1]);
    add_cc(r[0], c);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r[i], 0);
    addc(r[N - 1], 0);
    cond_sub_p<N>(r, pw);
}

// a*2^(-32N) mod p (canonical value of a Montgomery residue): the reduction rows only
template <int N, class PW>
__device__ __forceinline__ void mont_redc(uint32_t (&r)[N], const uint32_t (&a)[N], PW pw) {
    uint32_t X[2 * N + 2], Y[2 * N + 2];
#pragma unroll
    for (int i = 0; i < 2 * N + 2; i++) { X[i] = i < N ? a[i] : 0; Y[i] = 0; }
    uint32_t c = 0;
    MontRows<N, 0, PW>::reduce(X, Y, c, pw);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = X[N + i];
    add_cc(r[0], Y[N]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r[i], Y[N + i]);
    addc(r[N - 1], Y[2 * N - 1]);
    add_cc(r[0], c);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r[i], 0);
    addc(r[N - 1], 0);
    cond_sub_p<N>(r, pw);
}

}  // namespace m32
}  // namespace lcpc
