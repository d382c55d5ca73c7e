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
//
// The carry that leaves the last word of a chain lands one word further up.  In the PRODUCT rows that word has
// so far only collected such carries (row I-1's chains of the other parity end exactly there), so adding into it
// cannot overflow.  In the REDUCTION rows it already holds product data and may be all ones: there the carry goes
// to a separate sink Z (word index - N) that only ever holds carries, and is summed at the very end.  (Adding it
// into the data word drops it with probability ~2^-32 per chain: tests/golden/find_carry_kats.py builds operands
// that make it certain.)
template <int N, int I, bool SINK, class YW>
__device__ __forceinline__ void row_mad_impl(uint32_t (&X)[2 * N + 2], uint32_t (&Y)[2 * N + 2], uint32_t (&Z)[N + 2],
                                             uint32_t x, YW yw) {
    {
        constexpr int j0 = I & 1;  // I + j even
        constexpr int end = I + j0 + ((N - j0 + 1) / 2) * 2;
        mad_lo_cc(X[I + j0], x, yw(j0));
        madc_hi_cc(X[I + j0 + 1], x, yw(j0));
#pragma unroll
        for (int j = j0 + 2; j < N; j += 2) {
            madc_lo_cc(X[I + j], x, yw(j));
            madc_hi_cc(X[I + j + 1], x, yw(j));
        }
        if constexpr (SINK) addc(Z[end - N], 0);
        else addc(X[end], 0);
    }
    {
        constexpr int j0 = 1 - (I & 1);  // I + j odd
        constexpr int end = I + j0 + ((N - j0 + 1) / 2) * 2;
        mad_lo_cc(Y[I + j0], x, yw(j0));
        madc_hi_cc(Y[I + j0 + 1], x, yw(j0));
#pragma unroll
        for (int j = j0 + 2; j < N; j += 2) {
            madc_lo_cc(Y[I + j], x, yw(j));
            madc_hi_cc(Y[I + j + 1], x, yw(j));
        }
        if constexpr (SINK) addc(Z[end - N], 0);
        else addc(Y[end], 0);
    }
}
template <int N, int I, class YW>
__device__ __forceinline__ void row_mad(uint32_t (&X)[2 * N + 2], uint32_t (&Y)[2 * N + 2], uint32_t x, YW yw) {
    uint32_t unused[N + 2];
    row_mad_impl<N, I, false>(X, Y, unused, x, yw);
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
    __device__ __forceinline__ static void reduce(uint32_t (&X)[2 * N + 2], uint32_t (&Y)[2 * N + 2], uint32_t (&Z)[N + 2],
                                                  uint32_t &c, PW pw) {
        if constexpr (I < N) {
            const uint32_t m = 0u - (X[I] + Y[I] + c);
            row_mad_impl<N, I, true>(X, Y, Z, m, pw);
            c = (X[I] | Y[I] | c) != 0 ? 1u : 0u;
            MontRows<N, I + 1, PW>::reduce(X, Y, Z, c, pw);
        }
    }
};

// r = (X + Y) >> 32*SH, N words, plus the carry sink Z (word k of Z weighs 2^(32(N+k))) and the carry c of the
// last cleared word; the value is < 2p < 2^(32N), so nothing above word N-1 survives
template <int N, int SH>
__device__ __forceinline__ void sum_high(uint32_t (&r)[N], const uint32_t (&X)[2 * N + 2], const uint32_t (&Y)[2 * N + 2],
                                         uint32_t (&Z)[N + 2], uint32_t c) {
    constexpr int Z0 = SH - N;  // sink word that lines up with r[0]
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = X[SH + i];
    add_cc(r[0], Y[SH]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r[i], Y[SH + i]);
    addc(r[N - 1], Y[SH + N - 1]);
    Z[Z0] += c;  // both are tiny
    add_cc(r[0], Z[Z0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r[i], Z[Z0 + i]);
    addc(r[N - 1], Z[Z0 + N - 1]);
}

// r = a*b*2^(-32N) mod p, a, b < p
template <int N, class PW>
__device__ __forceinline__ void mont_mul(uint32_t (&r)[N], const uint32_t (&a)[N], const uint32_t (&b)[N], PW pw) {
    uint32_t X[2 * N + 2], Y[2 * N + 2], Z[N + 2];
#pragma unroll
    for (int i = 0; i < 2 * N + 2; i++) X[i] = Y[i] = 0;
#pragma unroll
    for (int i = 0; i < N + 2; i++) Z[i] = 0;
    MontRows<N, 0, PW>::product(X, Y, a, b);
    uint32_t c = 0;
    MontRows<N, 0, PW>::reduce(X, Y, Z, c, pw);
    sum_high<N, N>(r, X, Y, Z, c);
    cond_sub_p<N>(r, pw);
}

// a*2^(-32N) mod p (canonical value of a Montgomery residue): the reduction rows only
template <int N, class PW>
__device__ __forceinline__ void mont_redc(uint32_t (&r)[N], const uint32_t (&a)[N], PW pw) {
    uint32_t X[2 * N + 2], Y[2 * N + 2], Z[N + 2];
#pragma unroll
    for (int i = 0; i < 2 * N + 2; i++) { X[i] = i < N ? a[i] : 0; Y[i] = 0; }
#pragma unroll
    for (int i = 0; i < N + 2; i++) Z[i] = 0;
    uint32_t c = 0;
    MontRows<N, 0, PW>::reduce(X, Y, Z, c, pw);
    sum_high<N, N>(r, X, Y, Z, c);
    cond_sub_p<N>(r, pw);
}

// ---- lazy dot products -----------------------------------------------------------------------
// A sum of K products a_k*b_k (all operands < p) is accumulated UNREDUCED in 2N+2 words and reduced once:
// per term only the N^2 products of the multiplication remain, the N^2 products of the Montgomery
// reduction, the conditional subtraction and the modular addition are paid once per dot product
// (sparse matrix rows, folds and column checks are all dot products).
// The single reduction clears N+1 words instead of N, which tolerates S < 2^32 * p * 2^(32N), i.e. any
// K < 2^32, and still lands below 2p; it therefore returns S * 2^(-32(N+1)) mod p, and the caller
// pre-scales one operand class by 2^32 (a constant matrix, or the short tensor of a fold).

// S += a*b
template <int N, class PW>
__device__ __forceinline__ void wide_mac(uint32_t (&S)[2 * N + 2], const uint32_t (&a)[N], const uint32_t (&b)[N]) {
    uint32_t X[2 * N + 2], Y[2 * N + 2];
#pragma unroll
    for (int i = 0; i < 2 * N + 2; i++) X[i] = Y[i] = 0;
    MontRows<N, 0, PW>::product(X, Y, a, b);
    add_cc(S[0], X[0]);
#pragma unroll
    for (int i = 1; i < 2 * N + 1; i++) addc_cc(S[i], X[i]);
    addc(S[2 * N + 1], 0);
    add_cc(S[1], Y[1]);  // Y[0] is never written (odd-aligned pairs)
#pragma unroll
    for (int i = 2; i < 2 * N + 1; i++) addc_cc(S[i], Y[i]);
    addc(S[2 * N + 1], 0);
}

// r = S * 2^(-32(N+1)) mod p, fully reduced; S < 2^32 * p * 2^(32N).  N+1 reduction rows; the last one clears
// word N, which the carry sink may already have reached, so its digit and its carry count Z[0] in.
template <int N, class PW>
__device__ __forceinline__ void wide_redc(uint32_t (&r)[N], const uint32_t (&S)[2 * N + 2], PW pw) {
    uint32_t X[2 * N + 2], Y[2 * N + 2], Z[N + 2];
#pragma unroll
    for (int i = 0; i < 2 * N + 2; i++) { X[i] = S[i]; Y[i] = 0; }
#pragma unroll
    for (int i = 0; i < N + 2; i++) Z[i] = 0;
    uint32_t c = 0;
    MontRows<N, 0, PW>::reduce(X, Y, Z, c, pw);
    const uint32_t m = 0u - (X[N] + Y[N] + Z[0] + c);
    row_mad_impl<N, N, true>(X, Y, Z, m, pw);
    // X[N] + Y[N] + Z[0] + c is now a multiple of 2^32: 0, 2^32 or 2*2^32
    const uint64_t t = (uint64_t)X[N] + Y[N] + Z[0] + c;
    c = (uint32_t)(t >> 32);
    sum_high<N, N + 1>(r, X, Y, Z, c);
    cond_sub_p<N>(r, pw);
}

}  // namespace m32
}  // namespace lcpc
