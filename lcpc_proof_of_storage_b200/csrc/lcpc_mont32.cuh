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

// ---- running accumulators held as 64-bit register PAIRS -------------------------------------------------------
// The same sum kept as THREE running accumulators (even-aligned pairs, odd-aligned pairs, chain-end carries) so that a
// term costs its N^2 IMAD.WIDE and 2N carry catches and nothing else: wide_mac above builds every product from zero
// and then adds 2(2N+1) words into S, which is 40 % of its instructions at N = 8.  The sink words only count carries
// (at most two per term and word), so any K < 2^31 is safe.
// The accumulators are uint64_t (word pair 2i, 2i+1 of X; word pair 2i+1, 2i+2 of Y) and every carry chain is ONE asm
// statement (volatile like every other statement of this file that touches the carry flag: a chain that floated between
// two statements of an add_cc / addc_cc sequence would clobber the flag they hand over) that unpacks its pairs, runs mad.lo.cc / madc.hi.cc over them and packs them again: ptxas then keeps each
// pair in an aligned register pair for the whole loop.  With one 32-bit variable per word (row_mad above) it moved
// 2.5 registers per product between the loop-carried words and the pairs IMAD.WIDE needs (IMAD.MOV: 13 % of the FMA-pipe
// cycles of the four-limb Brakedown level, a third of all instructions of the one-limb one).
template <int NP>
struct MadChain;
template <>
struct MadChain<1> {
    __device__ __forceinline__ static void run(uint64_t *P, uint32_t x, const uint32_t *y, uint32_t &sink) {
        asm volatile("{ .reg .u32 l0, h0;\n\tmov.b64 {l0, h0}, %0;\n\t"
            "mad.lo.cc.u32 l0, %2, %3, l0;\n\tmadc.hi.cc.u32 h0, %2, %3, h0;\n\t"
            "addc.u32 %1, %1, 0;\n\tmov.b64 %0, {l0, h0}; }"
            : "+l"(P[0]), "+r"(sink) : "r"(x), "r"(y[0]));
    }
};
template <>
struct MadChain<2> {
    __device__ __forceinline__ static void run(uint64_t *P, uint32_t x, const uint32_t *y, uint32_t &sink) {
        asm volatile("{ .reg .u32 l0, h0, l1, h1;\n\tmov.b64 {l0, h0}, %0;\n\tmov.b64 {l1, h1}, %1;\n\t"
            "mad.lo.cc.u32 l0, %3, %4, l0;\n\tmadc.hi.cc.u32 h0, %3, %4, h0;\n\t"
            "madc.lo.cc.u32 l1, %3, %5, l1;\n\tmadc.hi.cc.u32 h1, %3, %5, h1;\n\t"
            "addc.u32 %2, %2, 0;\n\tmov.b64 %0, {l0, h0};\n\tmov.b64 %1, {l1, h1}; }"
            : "+l"(P[0]), "+l"(P[1]), "+r"(sink) : "r"(x), "r"(y[0]), "r"(y[2]));
    }
};
template <>
struct MadChain<3> {
    __device__ __forceinline__ static void run(uint64_t *P, uint32_t x, const uint32_t *y, uint32_t &sink) {
        asm volatile("{ .reg .u32 l0, h0, l1, h1, l2, h2;\n\tmov.b64 {l0, h0}, %0;\n\tmov.b64 {l1, h1}, %1;\n\tmov.b64 {l2, h2}, %2;\n\t"
            "mad.lo.cc.u32 l0, %4, %5, l0;\n\tmadc.hi.cc.u32 h0, %4, %5, h0;\n\t"
            "madc.lo.cc.u32 l1, %4, %6, l1;\n\tmadc.hi.cc.u32 h1, %4, %6, h1;\n\t"
            "madc.lo.cc.u32 l2, %4, %7, l2;\n\tmadc.hi.cc.u32 h2, %4, %7, h2;\n\t"
            "addc.u32 %3, %3, 0;\n\tmov.b64 %0, {l0, h0};\n\tmov.b64 %1, {l1, h1};\n\tmov.b64 %2, {l2, h2}; }"
            : "+l"(P[0]), "+l"(P[1]), "+l"(P[2]), "+r"(sink) : "r"(x), "r"(y[0]), "r"(y[2]), "r"(y[4]));
    }
};
template <>
struct MadChain<4> {
    __device__ __forceinline__ static void run(uint64_t *P, uint32_t x, const uint32_t *y, uint32_t &sink) {
        asm volatile("{ .reg .u32 l0, h0, l1, h1, l2, h2, l3, h3;\n\t"
            "mov.b64 {l0, h0}, %0;\n\tmov.b64 {l1, h1}, %1;\n\tmov.b64 {l2, h2}, %2;\n\tmov.b64 {l3, h3}, %3;\n\t"
            "mad.lo.cc.u32 l0, %5, %6, l0;\n\tmadc.hi.cc.u32 h0, %5, %6, h0;\n\t"
            "madc.lo.cc.u32 l1, %5, %7, l1;\n\tmadc.hi.cc.u32 h1, %5, %7, h1;\n\t"
            "madc.lo.cc.u32 l2, %5, %8, l2;\n\tmadc.hi.cc.u32 h2, %5, %8, h2;\n\t"
            "madc.lo.cc.u32 l3, %5, %9, l3;\n\tmadc.hi.cc.u32 h3, %5, %9, h3;\n\t"
            "addc.u32 %4, %4, 0;\n\t"
            "mov.b64 %0, {l0, h0};\n\tmov.b64 %1, {l1, h1};\n\tmov.b64 %2, {l2, h2};\n\tmov.b64 %3, {l3, h3}; }"
            : "+l"(P[0]), "+l"(P[1]), "+l"(P[2]), "+l"(P[3]), "+r"(sink)
            : "r"(x), "r"(y[0]), "r"(y[2]), "r"(y[4]), "r"(y[6]));
    }
};

// X[i] = words 2i, 2i+1;  Y[i] = words 2i+1, 2i+2;  Z[k] = carries of weight 2^(32(N+k))
template <int N>
struct SplitAcc {
    uint64_t X[N], Y[N];
    uint32_t Z[N + 2];
};
template <int N>
__device__ __forceinline__ void split_init(SplitAcc<N> &s) {
#pragma unroll
    for (int i = 0; i < N; i++) s.X[i] = s.Y[i] = 0;
#pragma unroll
    for (int i = 0; i < N + 2; i++) s.Z[i] = 0;
}
template <int N, int I>
__device__ __forceinline__ void split_rows(SplitAcc<N> &s, const uint32_t (&a)[N], const uint32_t (&b)[N]) {
    if constexpr (I < N) {
        {
            constexpr int j0 = I & 1, np = (N - j0 + 1) / 2, end = I + j0 + 2 * np;  // I + j even
            MadChain<np>::run(&s.X[(I + j0) / 2], b[I], &a[j0], s.Z[end - N]);
        }
        {
            constexpr int j0 = 1 - (I & 1), np = (N - j0 + 1) / 2, end = I + j0 + 2 * np;  // I + j odd
            MadChain<np>::run(&s.Y[(I + j0 - 1) / 2], b[I], &a[j0], s.Z[end - N]);
        }
        split_rows<N, I + 1>(s, a, b);
    }
}
// acc += a*b
template <int N>
__device__ __forceinline__ void split_mac(SplitAcc<N> &s, const uint32_t (&a)[N], const uint32_t (&b)[N]) {
    split_rows<N, 0>(s, a, b);
}
// S = X + Y + Z * 2^(32N), 2N + 2 words
template <int N>
__device__ __forceinline__ void split_sum(uint32_t (&S)[2 * N + 2], const SplitAcc<N> &s) {
    uint32_t Yw[2 * N + 2];
    Yw[0] = 0;
    Yw[2 * N + 1] = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        S[2 * i] = (uint32_t)s.X[i];
        S[2 * i + 1] = (uint32_t)(s.X[i] >> 32);
        Yw[2 * i + 1] = (uint32_t)s.Y[i];
        Yw[2 * i + 2] = (uint32_t)(s.Y[i] >> 32);
    }
    S[2 * N] = S[2 * N + 1] = 0;
    add_cc(S[1], Yw[1]);
#pragma unroll
    for (int i = 2; i < 2 * N + 1; i++) addc_cc(S[i], Yw[i]);
    addc(S[2 * N + 1], 0);
    add_cc(S[N], s.Z[0]);
#pragma unroll
    for (int i = 1; i < N + 1; i++) addc_cc(S[N + i], s.Z[i]);
    addc(S[2 * N + 1], s.Z[N + 1]);
}

// ---- Karatsuba dot products (N = 2H words) ------------------------------------------------------
// The N^2 IMAD.WIDE of a term are what bounds the Brakedown levels over the 255-bit field (FMA pipe; the ALU pipe is half
// idle, profiles/r02_spmv.md), so one Karatsuba level trades a quarter of them for ALU-pipe work: with a = aL + aH*W,
// b = bL + bH*W (W = 2^(32H)),  aL*bH + aH*bL = (aL + aH)(bL + bH) - aL*bL - aH*bH.  The three half-size products of
// every term go to three split accumulators and the subtraction is paid ONCE per dot product.  aL + aH and
// bL + bH can be one bit longer than H words; with sa = sa' + ca*W, sb = sb' + cb*W
//     sa*sb = sa'*sb' + W*(ca*sb' + cb*sa') + W^2*ca*cb,
// and the three correction sums are plain (predicated) additions into Ca, Cb and a counter.
template <int H>
struct KaraAcc {
    SplitAcc<H> p0, p2, pm;         // sum aL*bL, sum aH*bH, sum sa'*sb'
    uint32_t Ca[H + 1], Cb[H + 1];  // sum of sb' over the terms with ca, of sa' over those with cb
    uint32_t Cc;                    // number of terms with ca and cb
};

template <int H>
__device__ __forceinline__ void kara_init(KaraAcc<H> &k) {
    split_init<H>(k.p0);
    split_init<H>(k.p2);
    split_init<H>(k.pm);
#pragma unroll
    for (int i = 0; i < H + 1; i++) k.Ca[i] = k.Cb[i] = 0;
    k.Cc = 0;
}

// C += flag ? v : 0   (H words into H + 1; flag is 0 or 1)
template <int H>
__device__ __forceinline__ void cond_add(uint32_t (&C)[H + 1], const uint32_t (&v)[H], uint32_t flag) {
    if constexpr (H == 4) {
        asm volatile("{ .reg .pred q; setp.ne.u32 q, %5, 0;\n\t"
            "@q add.cc.u32 %0, %0, %6;\n\t@q addc.cc.u32 %1, %1, %7;\n\t@q addc.cc.u32 %2, %2, %8;\n\t"
            "@q addc.cc.u32 %3, %3, %9;\n\t@q addc.u32 %4, %4, 0; }"
            : "+r"(C[0]), "+r"(C[1]), "+r"(C[2]), "+r"(C[3]), "+r"(C[4])
            : "r"(flag), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]));
    } else {
        const uint32_t m = 0u - flag;
        add_cc(C[0], v[0] & m);
#pragma unroll
        for (int i = 1; i < H; i++) addc_cc(C[i], v[i] & m);
        addc(C[H], 0);
    }
}

// acc += a*b
template <int N>
__device__ __forceinline__ void kara_mac(KaraAcc<N / 2> &k, const uint32_t (&a)[N], const uint32_t (&b)[N]) {
    constexpr int H = N / 2;
    static_assert(N % 2 == 0, "even word count");
    uint32_t aL[H], aH[H], bL[H], bH[H], sa[H], sb[H];
#pragma unroll
    for (int i = 0; i < H; i++) { aL[i] = a[i]; aH[i] = a[H + i]; bL[i] = b[i]; bH[i] = b[H + i]; sa[i] = a[i]; sb[i] = b[i]; }
    uint32_t ca = 0, cb = 0;
    add_cc(sa[0], aH[0]);
#pragma unroll
    for (int i = 1; i < H; i++) addc_cc(sa[i], aH[i]);
    addc(ca, 0);
    add_cc(sb[0], bH[0]);
#pragma unroll
    for (int i = 1; i < H; i++) addc_cc(sb[i], bH[i]);
    addc(cb, 0);
    split_mac<H>(k.p0, aL, bL);
    split_mac<H>(k.p2, aH, bH);
    split_mac<H>(k.pm, sa, sb);
    cond_add<H>(k.Ca, sb, ca);
    cond_add<H>(k.Cb, sa, cb);
    k.Cc += ca & cb;
}

// S = the accumulated sum, 2N + 2 words
template <int N>
__device__ __forceinline__ void kara_sum(uint32_t (&S)[2 * N + 2], const KaraAcc<N / 2> &k) {
    constexpr int H = N / 2, W2 = 2 * H + 2;
    uint32_t S0[W2], S2[W2], Sm[W2];
    split_sum<H>(S0, k.p0);
    split_sum<H>(S2, k.p2);
    split_sum<H>(Sm, k.pm);
    // Sm += (Ca + Cb) * W + Cc * W^2
    add_cc(Sm[H], k.Ca[0]);
#pragma unroll
    for (int i = 1; i < H + 1; i++) addc_cc(Sm[H + i], k.Ca[i]);
    addc(Sm[2 * H + 1], 0);
    add_cc(Sm[H], k.Cb[0]);
#pragma unroll
    for (int i = 1; i < H + 1; i++) addc_cc(Sm[H + i], k.Cb[i]);
    addc(Sm[2 * H + 1], 0);
    add_cc(Sm[2 * H], k.Cc);
    addc(Sm[2 * H + 1], 0);
    // Sm -= S0 + S2  (what is left is sum aL*bH + aH*bL >= 0)
    sub_cc(Sm[0], S0[0]);
#pragma unroll
    for (int i = 1; i < W2 - 1; i++) subc_cc(Sm[i], S0[i]);
    subc(Sm[W2 - 1], S0[W2 - 1]);
    sub_cc(Sm[0], S2[0]);
#pragma unroll
    for (int i = 1; i < W2 - 1; i++) subc_cc(Sm[i], S2[i]);
    subc(Sm[W2 - 1], S2[W2 - 1]);
    // S = S0 + S2 * W^2 + Sm * W
#pragma unroll
    for (int i = 0; i < 2 * N + 2; i++) S[i] = i < W2 ? S0[i] : 0;
    add_cc(S[2 * H], S2[0]);
#pragma unroll
    for (int i = 1; i < W2 - 1; i++) addc_cc(S[2 * H + i], S2[i]);
    addc(S[2 * H + W2 - 1], S2[W2 - 1]);
    add_cc(S[H], Sm[0]);
#pragma unroll
    for (int i = 1; i < W2; i++) addc_cc(S[H + i], Sm[i]);
#pragma unroll
    for (int i = H + W2; i < 2 * N + 1; i++) addc_cc(S[i], 0);
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
