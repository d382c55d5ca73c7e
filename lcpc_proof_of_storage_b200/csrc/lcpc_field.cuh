// Prime-field arithmetic for the lcpc fields on sm_100a.
//
// Elements are LIMBS x u64, least-significant limb first, Montgomery form with
// R = 2^(64*LIMBS), always fully reduced on loads/stores to global memory -- the
// in-memory form ff_derive 0.13 gives the reference's fields
// (lcpc-test-fields/src/lib.rs:18-70; WriteableFt63: proof-of-storage/src/fields/
// writable_ft63.rs:8-12; Ft253_192: proof-of-storage/src/fields/ft253_192.rs:6-10).  Every modulus is
// 1 mod 2^40 (Ft253_192: 1 mod 2^192), so -p^-1 mod 2^32 = 0xffffffff: the Montgomery quotient digit
// is a negation.  Ft253_192 is the one field whose to_repr() is BIG-endian (PrimeFieldReprEndianness =
// "big"): Field::to_repr() below returns the element whose memory image is those bytes.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "lcpc_mont32.cuh"

#ifndef LCPC_FT63_DIGIT1_SUM3
#define LCPC_FT63_DIGIT1_SUM3 1
#endif
#ifndef LCPC_KARATSUBA
#define LCPC_KARATSUBA 1
#endif

namespace lcpc {

enum FieldId : int { FT63 = 0, FT127 = 1, FT191 = 2, FT255 = 3, FT253_192 = 4, N_FIELDS = 5 };

template <int L>
struct Fe {
    uint64_t v[L];
};

// Host-visible constant tables (also used by the API for lcpc_field_constants).
struct FieldConsts {
    int limbs, num_bits, two_adicity;
    uint64_t p[4], inv, r[4], r2[4], root[4];
    bool repr_big_endian;
};

__host__ __device__ constexpr FieldConsts field_consts(int fid) {
    switch (fid) {
    case FT63:
        return {1, 63, 41,
                {0x46d0760000000001ull, 0, 0, 0},
                0x46d075ffffffffffull,
                {0x2b8e9dfffffffffdull, 0, 0, 0},
                {0x13085abb0716119eull, 0, 0, 0},
                {0x23bcb75f84213a43ull, 0, 0, 0}, false};
    case FT127:
        return {2, 127, 40,
                {0x7f2bd90000000001ull, 0x6e754097ba20e0bfull, 0, 0},
                0x7f2bd8ffffffffffull,
                {0x01a84dfffffffffeull, 0x23157ed08bbe3e81ull, 0, 0},
                {0x816bd5407cf6dce5ull, 0x2c1637057de6fce8ull, 0, 0},
                {0xf491a1dff39975f8ull, 0x178fd41c0f6a04faull, 0, 0}, false};
    case FT191:
        return {3, 191, 41,
                {0xd246820000000001ull, 0x936888270ceecbcdull, 0x453708aa3fbc8ddaull, 0},
                0xd24681ffffffffffull,
                {0x892c79fffffffffdull, 0x45c6678ad9339c96ull, 0x305ae60140ca5670ull, 0},
                {0x6c25128031d873e2ull, 0xf71a3697a97ffdceull, 0x07ef71ae547daef9ull, 0},
                {0xecd905456df2b092ull, 0x53ce189f0df0a05aull, 0x3f6e6da556ed31d9ull, 0}, false};
    case FT253_192:  // p = (2^61 - 1) * 2^192 + 1, generator 3, 2-adicity 192
        return {4, 253, 192,
                {0x0000000000000001ull, 0, 0, 0x1fffffffffffffffull},
                0xffffffffffffffffull,
                {0xfffffffffffffff8ull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0x0000000000000007ull},
                {0xffffffffffff8040ull, 0xffffffffffffefffull, 0xfffffffffffffdffull, 0x0000000000007f7full},
                {0xe94731e93d73da14ull, 0x0e0f79fb69eec7bfull, 0x246cdb0e8f061ce3ull, 0x029543679ced2616ull}, true};
    default:
        return {4, 255, 41,
                {0x02a4f20000000001ull, 0xef73c79086595f30ull, 0xfda9df04b9575969ull, 0x663c799b6e4d2900ull},
                0x02a4f1ffffffffffull,
                {0xfab61bfffffffffeull, 0x211870def34d419full, 0x04ac41f68d514d2cull, 0x33870cc92365adfeull},
                {0xcf06aad260ab9990ull, 0x12f0d8856156a683ull, 0x5da77ded73588e21ull, 0x38725a1646845639ull},
                {0x9c745ae52a496067ull, 0x95ee9a4091329682ull, 0x854a3ee53365b80eull, 0x16edffae79969e76ull}, false};
    }
}

// 2^32 * R mod p (the Montgomery form of 2^32), limbs; used to undo the extra 2^-32 of the lazy dot-product reduction
struct Limbs4 {
    uint64_t v[4];
};
__host__ __device__ constexpr Limbs4 field_two32_mont(int fid) {
    const FieldConsts fc = field_consts(fid);
    Limbs4 x{{fc.r[0], fc.r[1], fc.r[2], fc.r[3]}};
    for (int it = 0; it < 32; it++) {
        // x = 2x mod p   (2x < 2^(64*limbs) for every lcpc modulus: one spare bit)
        uint64_t carry = 0;
        for (int i = 0; i < 4; i++) {
            const uint64_t nv = (x.v[i] << 1) | carry;
            carry = x.v[i] >> 63;
            x.v[i] = nv;
        }
        bool ge = true;  // x >= p ?
        for (int i = 3; i >= 0; i--) {
            if (x.v[i] != fc.p[i]) {
                ge = x.v[i] > fc.p[i];
                break;
            }
        }
        if (ge) {
            uint64_t borrow = 0;
            for (int i = 0; i < 4; i++) {
                const uint64_t d = x.v[i] - fc.p[i];
                const uint64_t b1 = x.v[i] < fc.p[i];
                const uint64_t d2 = d - borrow;
                const uint64_t b2 = d < borrow;
                x.v[i] = d2;
                borrow = b1 | b2;
            }
        }
    }
    return x;
}

// ---- 63-bit field fast path ----------------------------------------------------------
// p = 0x46d07600_00000001: low word 1, high word P_HI.  A product is four 32x32 -> 64 partial products
// (IMAD.WIDE) plus two Montgomery digits of one IMAD.WIDE each: six IMAD.WIDE in all, and IMAD.WIDE is the
// expensive instruction on this part (4 issue cycles on the FMA pipe and 2 on the ALU side, against 2 for IMAD /
// IADD3 / LOP3 / SHF: profiles/r01c_summary.md).  Everything around them is written to need no compare and no
// select: column sums are three-input carry chains, a digit is `acc + w*Q` with the high word adjusted (see
// redc_digit), and the final correction is fix().
// The digit multiplier is read from the constant bank so that ptxas keeps it one IMAD.WIDE (a literal modulus
// word is split into IMAD + IMAD.HI + a 3-input add).
static __constant__ uint32_t LCPC_FT63_Q = 0xb92f8a00u;   // 2^32 - P_HI

namespace ft63 {
constexpr uint64_t P = 0x46d0760000000001ull;
constexpr uint32_t P_HI = 0x46d07600u;
constexpr uint64_t NEG_P = 0 - P;                         // 2^64 - p

__device__ __forceinline__ uint64_t wmul(uint32_t a, uint32_t b) {
    uint64_t d;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t lo32(uint64_t x) { return (uint32_t)x; }
__device__ __forceinline__ uint32_t hi32(uint64_t x) { return (uint32_t)(x >> 32); }
__device__ __forceinline__ uint64_t pack(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}

// v in [-p, p) as a two's-complement 64-bit value -> v mod p in [0, p).  p < 2^63, so the sign bit of the
// high word alone decides.  Callers fold the "- p" into work they do anyway (a + b + NEG_P is one
// three-input carry chain; the Montgomery reduction leaves v - p for free), and the correction itself is
// SHF + IADD3 on the ALU pipe and one IMAD.X (sign * P_HI + hi + carry) on the FMA pipe: 2 ALU slots
// where compare + select needs 5.  The integer ALU pipe is what bounds the NTT (profiles/r01b_summary.md).
__device__ __forceinline__ uint64_t fix(uint64_t v) {
    uint32_t lo = lo32(v), hi = hi32(v);
    asm("{ .reg .u32 m; shr.u32 m, %1, 31; add.cc.u32 %0, %0, m; madc.lo.u32 %1, m, 0x46d07600, %1; }" : "+r"(lo), "+r"(hi));
    return pack(lo, hi);
}

// One Montgomery digit (2^-32) of the value  w + acc * 2^32  (w a 32-bit word, acc the 64 bits above
// it, P already included in acc): with m = 2^32 - w (also for w = 0) the low word of w + m*p is exactly
// 2^32, so the quotient is acc' = acc_orig + 1 + m*P_HI = acc + w*Q - w*2^32 (mod 2^64; the true value
// fits), Q = 2^32 - P_HI, and "+ 1 + 2^32*P_HI" is "+ P".  No carry flags, no compare.
__device__ __forceinline__ uint64_t redc_digit(uint32_t w, uint64_t acc) {
    const uint64_t u = acc + wmul(w, LCPC_FT63_Q);
    return pack(lo32(u), hi32(u) - w);
}

// Montgomery reduction of p00 + col1*2^32 + p11*2^64 (sums of 32x32 products, not carry-normalised;
// col1 < 2.9 * 2^62).  The second digit's "+ P" is left out, so the digit chain ends on v - p in [-p, p)
// and fix() finishes.  Result in [0, p).
template <bool SUM3 = true>
__device__ __forceinline__ uint64_t redc_cols(uint64_t p00, uint64_t col1, uint64_t p11) {
#if LCPC_FT63_DIGIT1_SUM3
  if constexpr (SUM3) {
    // first digit = redc_digit(lo32(p00), col1 + P + hi32(p00)), spelled as ONE three-operand 64-bit sum
    //     (col1 + w*Q) + {1, P_HI + 1} + {hi32(p00), ~w}        (~w + 1 = -w: the "- w * 2^32" of redc_digit)
    // which ptxas emits as IADD3 + IADD3.X with two carries.  Written the other way it hangs hi32(p00) on the IMAD.WIDE
    // chain as a zero-extended accumulator -- two register moves to build the {hi, 0} pair -- and subtracts w with a
    // third instruction: 2 instructions fewer per product, both off the FMA pipe (profiles/r02_summary.md).
    const uint32_t w = lo32(p00);
    const uint64_t u = (col1 + wmul(w, LCPC_FT63_Q)) + (P + (1ull << 32)) + pack(hi32(p00), ~w);
    // second digit = redc_digit(w2, p11 + hi32(u)) = p11 + hi32(u) + w2*Q - w2*2^32: the two words that join the products
    // form ONE 64-bit addend {hi32(u), -w2}, built in place (no zero-extended pair, no separate "hi - w2")
    const uint32_t w2 = lo32(u);
    const uint64_t v = p11 + pack(hi32(u), 0u - w2) + wmul(w2, LCPC_FT63_Q);
    return fix(v);
  }
#endif
    const uint64_t u = redc_digit(lo32(p00), col1 + P + hi32(p00));
    const uint64_t v = redc_digit(lo32(u), p11 + hi32(u));
    return fix(v);
}

// a*b*2^-64 mod p, a, b < p
__device__ __forceinline__ uint64_t mul(uint64_t a, uint64_t b) {
    const uint32_t a0 = lo32(a), a1 = hi32(a), b0 = lo32(b), b1 = hi32(b);
    return redc_cols(wmul(a0, b0), wmul(a0, b1) + wmul(a1, b0), wmul(a1, b1));
}
// a*2^-64 mod p
// a*2^-64 mod p: the same two digits with no partial products: both words that join a digit product go in as one
// 64-bit addend ({hi32(a), ~w} and {hi32(u), -w2}), so neither needs a zero-extended register pair
__device__ __forceinline__ uint64_t to_canon(uint64_t a) {
#if LCPC_FT63_DIGIT1_SUM3
    const uint32_t w = lo32(a);
    const uint64_t u = wmul(w, LCPC_FT63_Q) + pack(hi32(a), ~w) + (P + (1ull << 32));
    const uint32_t w2 = lo32(u);
    return fix(pack(hi32(u), 0u - w2) + wmul(w2, LCPC_FT63_Q));
#else
    return redc_cols<false>((uint64_t)lo32(a), (uint64_t)hi32(a), 0);
#endif
}
__device__ __forceinline__ uint64_t add(uint64_t a, uint64_t b) { return fix(a + b + NEG_P); }
__device__ __forceinline__ uint64_t sub(uint64_t a, uint64_t b) { return fix(a - b); }
}  // namespace ft63

// Compile-time field description: P(i) etc. are constexpr calls that fold to
// immediates inside fully unrolled loops.
template <int FID>
struct Field {
    static constexpr int LIMBS = field_consts(FID).limbs;
    static constexpr int NUM_BITS = field_consts(FID).num_bits;
    static constexpr int TWO_ADICITY = field_consts(FID).two_adicity;
    using E = Fe<LIMBS>;

    __host__ __device__ static constexpr uint64_t P(int i) { return field_consts(FID).p[i]; }
    __host__ __device__ static constexpr uint64_t INV() { return field_consts(FID).inv; }
    __host__ __device__ static constexpr uint64_t RMODP(int i) { return field_consts(FID).r[i]; }

    // 32-bit word i of the modulus (folds to an immediate)
    struct PWord {
        __host__ __device__ constexpr uint32_t operator()(int i) const { return (uint32_t)(P(i >> 1) >> (32 * (i & 1))); }
    };
    __device__ __forceinline__ static void split(uint32_t (&w)[2 * LIMBS], const E &a) {
#pragma unroll
        for (int i = 0; i < LIMBS; i++) {
            w[2 * i] = (uint32_t)a.v[i];
            w[2 * i + 1] = (uint32_t)(a.v[i] >> 32);
        }
    }
    __device__ __forceinline__ static E join(const uint32_t (&w)[2 * LIMBS]) {
        E r;
#pragma unroll
        for (int i = 0; i < LIMBS; i++) r.v[i] = ((uint64_t)w[2 * i + 1] << 32) | w[2 * i];
        return r;
    }

    __device__ __forceinline__ static E zero() {
        E r;
#pragma unroll
        for (int i = 0; i < LIMBS; i++) r.v[i] = 0;
        return r;
    }
    __device__ __forceinline__ static E one() {
        E r;
#pragma unroll
        for (int i = 0; i < LIMBS; i++) r.v[i] = RMODP(i);
        return r;
    }
    __device__ __forceinline__ static bool is_zero(const E &a) {
        uint64_t o = 0;
#pragma unroll
        for (int i = 0; i < LIMBS; i++) o |= a.v[i];
        return o == 0;
    }

    // a >= p ?
    __device__ __forceinline__ static bool geq_p(const uint64_t *a) {
        if constexpr (LIMBS == 1) {
            return a[0] >= P(0);
        } else {
            // borrow of a - p
            uint64_t borrow = 0;
#pragma unroll
            for (int i = 0; i < LIMBS; i++) {
                uint64_t pi = P(i);
                uint64_t d = a[i] - pi;
                uint64_t b1 = a[i] < pi;
                uint64_t b2 = d < borrow;
                borrow = b1 | b2;
            }
            return borrow == 0;
        }
    }
    __device__ __forceinline__ static void sub_p(uint64_t *a) {
        uint64_t borrow = 0;
#pragma unroll
        for (int i = 0; i < LIMBS; i++) {
            uint64_t pi = P(i);
            uint64_t d = a[i] - pi;
            uint64_t b1 = a[i] < pi;
            uint64_t d2 = d - borrow;
            uint64_t b2 = d < borrow;
            a[i] = d2;
            borrow = b1 | b2;
        }
    }

    __device__ __forceinline__ static E add(const E &a, const E &b) {
        E r;
        if constexpr (LIMBS == 1) {
            r.v[0] = ft63::add(a.v[0], b.v[0]);
        } else {
            uint32_t x[2 * LIMBS], y[2 * LIMBS], z[2 * LIMBS];
            split(x, a);
            split(y, b);
            m32::add_mod<2 * LIMBS>(z, x, y, PWord{});
            r = join(z);
        }
        return r;
    }

    __device__ __forceinline__ static E sub(const E &a, const E &b) {
        E r;
        if constexpr (LIMBS == 1) {
            r.v[0] = ft63::sub(a.v[0], b.v[0]);
        } else {
            uint32_t x[2 * LIMBS], y[2 * LIMBS], z[2 * LIMBS];
            split(x, a);
            split(y, b);
            m32::sub_mod<2 * LIMBS>(z, x, y, PWord{});
            r = join(z);
        }
        return r;
    }

    // difference that is only consumed as the first operand of mul() (hook for lazy reduction)
    __device__ __forceinline__ static E sub_for_mul(const E &a, const E &b) {
        return sub(a, b);
    }

    // Montgomery product a*b*R^-1 mod p, fully reduced.
    __device__ __forceinline__ static E mul(const E &a, const E &b) {
        E r;
        if constexpr (LIMBS == 1) {
            r.v[0] = ft63::mul(a.v[0], b.v[0]);
        } else {
            uint32_t x[2 * LIMBS], y[2 * LIMBS], z[2 * LIMBS];
            split(x, a);
            split(y, b);
            m32::mont_mul<2 * LIMBS>(z, x, y, PWord{});
            r = join(z);
        }
        return r;
    }

    // canonical value (PrimeField::to_repr limbs): a * R^-1 mod p
    __device__ __forceinline__ static E to_canon(const E &a) {
        if constexpr (LIMBS == 1) {
            E r;
            r.v[0] = ft63::to_canon(a.v[0]);
            return r;
        } else {
            uint32_t x[2 * LIMBS], z[2 * LIMBS];
            split(x, a);
            m32::mont_redc<2 * LIMBS>(z, x, PWord{});
            return join(z);
        }
    }

    // PrimeField::to_repr() as an element whose in-memory bytes ARE the repr: the canonical value itself for the
    // little-endian fields; for Ft253_192 the byte-reversed canonical value (limb order and the bytes of every limb)
    static constexpr bool REPR_BE = field_consts(FID).repr_big_endian;
    __device__ __forceinline__ static E to_repr(const E &a) {
        const E c = to_canon(a);
        if constexpr (!REPR_BE) {
            return c;
        } else {
            E r;
#pragma unroll
            for (int i = 0; i < LIMBS; i++) {
                const uint64_t v = c.v[LIMBS - 1 - i];
                r.v[i] = ((uint64_t)__byte_perm((uint32_t)v, 0, 0x0123) << 32) | __byte_perm((uint32_t)(v >> 32), 0, 0x0123);
            }
            return r;
        }
    }

    // ---- lazy dot products: acc = sum_k a_k * b_k with ONE reduction (see lcpc_mont32.cuh) -----------------
    // Every field accumulates unreduced double-width products and pays the Montgomery reduction once per output.
    // Multi-limb fields: one 2N+2-word sum per accumulator, every product built from zero and added (18 registers for
    // four limbs: the fold keeps up to four of them per thread).  One-limb field: the split running accumulator of the
    // Brakedown levels (m32::SplitAcc<2>: 4 IMAD.WIDE + 2 carry catches per term against ~22 instructions for the eager
    // Montgomery multiply-add: the four-tensor fold is a stream over the matrix again, not an integer-pipe kernel).
    struct DotSum {
        uint32_t s[4 * LIMBS + 2];
    };
    struct DotSplit1 {
        m32::SplitAcc<2 * LIMBS> s;
    };
    using Dot = typename std::conditional<LIMBS == 1, DotSplit1, DotSum>::type;
    __device__ __forceinline__ static void dot_init(Dot &d) {
        if constexpr (LIMBS == 1) {
            m32::split_init<2>(d.s);
        } else {
#pragma unroll
            for (int i = 0; i < 4 * LIMBS + 2; i++) d.s[i] = 0;
        }
    }
    __device__ __forceinline__ static void dot_mac(Dot &d, const E &a, const E &b) {
        uint32_t x[2 * LIMBS], y[2 * LIMBS];
        split(x, a);
        split(y, b);
        if constexpr (LIMBS == 1) m32::split_mac<2>(d.s, x, y);
        else m32::wide_mac<2 * LIMBS, PWord>(d.s, x, y);
    }
    // sum * 2^-32 (for callers that pre-scaled one operand class by 2^32)
    __device__ __forceinline__ static E dot_finish_prescaled(const Dot &d) {
        uint32_t z[2 * LIMBS];
        if constexpr (LIMBS == 1) {
            uint32_t s[4 * LIMBS + 2];
            m32::split_sum<2>(s, d.s);
            m32::wide_redc<2 * LIMBS>(z, s, PWord{});
        } else {
            m32::wide_redc<2 * LIMBS>(z, d.s, PWord{});
        }
        return join(z);
    }
    // 2^32 in Montgomery form
    __device__ __forceinline__ static E dot_scale() {
        E r;
#pragma unroll
        for (int i = 0; i < LIMBS; i++) r.v[i] = field_two32_mont(FID).v[i];
        return r;
    }
    // the sum, fully reduced, no pre-scaling needed: one extra product per dot product
    __device__ __forceinline__ static E dot_finish(const Dot &d) { return mul(dot_finish_prescaled(d), dot_scale()); }

    // The same dot product on running accumulators that cost fewer instructions per term and more registers -- for
    // the Brakedown levels, whose constant operand class (the matrix) is pre-scaled by 2^32 at plan time for EVERY field
    // (DOTW prescale; dotw_scale()).  Four-limb fields: one Karatsuba level on three split accumulators (m32::kara_mac,
    // 48 IMAD.WIDE per term instead of 64); one to three limbs: split accumulators (m32::split_mac) -- for the one-limb
    // field that is 4 IMAD.WIDE.X + 4 carry catches per term against 6 IMAD.WIDE + ~20 other instructions for the eager
    // Montgomery multiply-add.
    static constexpr bool DOTW_KARA = LIMBS == 4 && LCPC_KARATSUBA;
    struct DotWSplit {
        m32::SplitAcc<2 * LIMBS> s;
    };
    struct DotWKara {
        m32::KaraAcc<LIMBS> k;
    };
    using DotW = typename std::conditional<DOTW_KARA, DotWKara, DotWSplit>::type;
    __device__ __forceinline__ static void dotw_init(DotW &d) {
        if constexpr (DOTW_KARA) {
            m32::kara_init<LIMBS>(d.k);
        } else {
            m32::split_init<2 * LIMBS>(d.s);
        }
    }
    __device__ __forceinline__ static void dotw_mac(DotW &d, const E &a, const E &b) {
        uint32_t x[2 * LIMBS], y[2 * LIMBS];
        split(x, a);
        split(y, b);
        if constexpr (DOTW_KARA) m32::kara_mac<2 * LIMBS>(d.k, x, y);
        else m32::split_mac<2 * LIMBS>(d.s, x, y);
    }
    // sum * 2^-32 (one operand class pre-scaled by dotw_scale())
    __device__ __forceinline__ static E dotw_finish_prescaled(const DotW &d) {
        uint32_t s[4 * LIMBS + 2], z[2 * LIMBS];
        if constexpr (DOTW_KARA) m32::kara_sum<2 * LIMBS>(s, d.k);
        else m32::split_sum<2 * LIMBS>(s, d.s);
        m32::wide_redc<2 * LIMBS>(z, s, PWord{});
        return join(z);
    }
    // 2^32 in Montgomery form (what the pre-scaled operand class is multiplied by at plan time)
    __device__ __forceinline__ static E dotw_scale() { return dot_scale(); }

    __device__ static E pow(E base, uint64_t e) {
        E acc = one();
        while (e) {
            if (e & 1) acc = mul(acc, base);
            base = mul(base, base);
            e >>= 1;
        }
        return acc;
    }
};

// 16-byte / 32-byte vector loads and stores of one element (global memory is AoS).
template <int L>
__device__ __forceinline__ Fe<L> ld_fe(const uint64_t *p) {
    Fe<L> r;
    if constexpr (L == 1) {
        r.v[0] = p[0];
    } else if constexpr (L == 2) {
        ulonglong2 t = *reinterpret_cast<const ulonglong2 *>(p);
        r.v[0] = t.x; r.v[1] = t.y;
    } else if constexpr (L == 4) {
        ulonglong2 t0 = reinterpret_cast<const ulonglong2 *>(p)[0];
        ulonglong2 t1 = reinterpret_cast<const ulonglong2 *>(p)[1];
        r.v[0] = t0.x; r.v[1] = t0.y; r.v[2] = t1.x; r.v[3] = t1.y;
    } else {
#pragma unroll
        for (int i = 0; i < L; i++) r.v[i] = p[i];
    }
    return r;
}

template <int L>
__device__ __forceinline__ void st_fe(uint64_t *p, const Fe<L> &a) {
    if constexpr (L == 1) {
        p[0] = a.v[0];
    } else if constexpr (L == 2) {
        *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(a.v[0], a.v[1]);
    } else if constexpr (L == 4) {
        reinterpret_cast<ulonglong2 *>(p)[0] = make_ulonglong2(a.v[0], a.v[1]);
        reinterpret_cast<ulonglong2 *>(p)[1] = make_ulonglong2(a.v[2], a.v[3]);
    } else {
#pragma unroll
        for (int i = 0; i < L; i++) p[i] = a.v[i];
    }
}

}  // namespace lcpc
