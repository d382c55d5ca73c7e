// BLAKE3 compression function for sm_100a, fully unrolled: the message schedule is a
// compile-time index table, so the per-round permutation costs no moves.
//
// The reference hashes with blake3::Hasher through digest::Digest (plain hash mode,
// 32-byte output): leaves at lcpc-2d/src/lib.rs:749-764, internal nodes at :800-805.
#pragma once
#include <cstdint>

namespace lcpc {
namespace b3 {

enum : uint32_t { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };
constexpr uint32_t BLOCK_BYTES = 64;
constexpr uint32_t CHUNK_BYTES = 1024;

#define LCPC_B3_IV0 0x6A09E667u
#define LCPC_B3_IV1 0xBB67AE85u
#define LCPC_B3_IV2 0x3C6EF372u
#define LCPC_B3_IV3 0xA54FF53Au
#define LCPC_B3_IV4 0x510E527Fu
#define LCPC_B3_IV5 0x9B05688Cu
#define LCPC_B3_IV6 0x1F83D9ABu
#define LCPC_B3_IV7 0x5BE0CD19u

__device__ __forceinline__ void set_iv(uint32_t cv[8]) {
    cv[0] = LCPC_B3_IV0; cv[1] = LCPC_B3_IV1; cv[2] = LCPC_B3_IV2; cv[3] = LCPC_B3_IV3;
    cv[4] = LCPC_B3_IV4; cv[5] = LCPC_B3_IV5; cv[6] = LCPC_B3_IV6; cv[7] = LCPC_B3_IV7;
}

// schedule[r][i] = index of the original message word used at position i in round r
struct Schedule {
    int s[7][16];
};
__host__ __device__ constexpr Schedule make_schedule() {
    constexpr int perm[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
    Schedule sc{};
    for (int i = 0; i < 16; i++) sc.s[0][i] = i;
    for (int r = 1; r < 7; r++)
        for (int i = 0; i < 16; i++) sc.s[r][i] = sc.s[r - 1][perm[i]];
    return sc;
}

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

// c + d on the FMA pipe (IMAD with a multiplier of 1 read from the constant bank, which ptxas
// cannot strength-reduce back to an IADD3): the G function is otherwise 100 % ALU-pipe work
// (3-input adds, xors, funnel shifts, byte permutes), and that pipe is what bounds the kernel.
static __constant__ uint32_t LCPC_B3_ONE = 1u;
__device__ __forceinline__ uint32_t add_fma(uint32_t c, uint32_t d, uint32_t one) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(d), "r"(one), "r"(c));
    return r;
}

#define LCPC_B3_G(a, b, c, d, mx, my) \
    do {                              \
        a = add_fma(add_fma(a, b, one), (mx), one); \
        d = __byte_perm(d ^ a, 0, 0x1032); \
        c = add_fma(c, d, one);       \
        b = rotr(b ^ c, 12);          \
        a = add_fma(add_fma(a, b, one), (my), one); \
        d = __byte_perm(d ^ a, 0, 0x0321); \
        c = add_fma(c, d, one);       \
        b = rotr(b ^ c, 7);           \
    } while (0)

// The same G with plain three-input adds (one IADD3 where the FMA-pipe form needs two dependent IMADs): shorter
// dependency chain, more ALU-pipe work -- for the Merkle levels, where a handful of warps wait on a serial chain of
// compressions and no pipe is anywhere near busy.
#define LCPC_B3_G_LAT(a, b, c, d, mx, my) \
    do {                                  \
        a = a + b + (mx);                 \
        d = __byte_perm(d ^ a, 0, 0x1032); \
        c = c + d;                        \
        b = rotr(b ^ c, 12);              \
        a = a + b + (my);                 \
        d = __byte_perm(d ^ a, 0, 0x0321); \
        c = c + d;                        \
        b = rotr(b ^ c, 7);               \
    } while (0)

// cv <- compress(cv, m, counter, block_len, flags), keeping only the chaining value.  LAT: latency-oriented G.
template <bool LAT = false>
__device__ __forceinline__ void compress(uint32_t cv[8], const uint32_t m[16], uint64_t counter,
                                         uint32_t block_len, uint32_t flags) {
    constexpr Schedule SC = make_schedule();
    const uint32_t one = LCPC_B3_ONE;
    uint32_t s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    uint32_t s8 = LCPC_B3_IV0, s9 = LCPC_B3_IV1, s10 = LCPC_B3_IV2, s11 = LCPC_B3_IV3;
    uint32_t s12 = (uint32_t)counter, s13 = (uint32_t)(counter >> 32), s14 = block_len, s15 = flags;
    if constexpr (LAT) {
#pragma unroll
        for (int r = 0; r < 7; r++) {
            LCPC_B3_G_LAT(s0, s4, s8, s12, m[SC.s[r][0]], m[SC.s[r][1]]);
            LCPC_B3_G_LAT(s1, s5, s9, s13, m[SC.s[r][2]], m[SC.s[r][3]]);
            LCPC_B3_G_LAT(s2, s6, s10, s14, m[SC.s[r][4]], m[SC.s[r][5]]);
            LCPC_B3_G_LAT(s3, s7, s11, s15, m[SC.s[r][6]], m[SC.s[r][7]]);
            LCPC_B3_G_LAT(s0, s5, s10, s15, m[SC.s[r][8]], m[SC.s[r][9]]);
            LCPC_B3_G_LAT(s1, s6, s11, s12, m[SC.s[r][10]], m[SC.s[r][11]]);
            LCPC_B3_G_LAT(s2, s7, s8, s13, m[SC.s[r][12]], m[SC.s[r][13]]);
            LCPC_B3_G_LAT(s3, s4, s9, s14, m[SC.s[r][14]], m[SC.s[r][15]]);
        }
    } else {
#pragma unroll
    for (int r = 0; r < 7; r++) {
        LCPC_B3_G(s0, s4, s8, s12, m[SC.s[r][0]], m[SC.s[r][1]]);
        LCPC_B3_G(s1, s5, s9, s13, m[SC.s[r][2]], m[SC.s[r][3]]);
        LCPC_B3_G(s2, s6, s10, s14, m[SC.s[r][4]], m[SC.s[r][5]]);
        LCPC_B3_G(s3, s7, s11, s15, m[SC.s[r][6]], m[SC.s[r][7]]);
        LCPC_B3_G(s0, s5, s10, s15, m[SC.s[r][8]], m[SC.s[r][9]]);
        LCPC_B3_G(s1, s6, s11, s12, m[SC.s[r][10]], m[SC.s[r][11]]);
        LCPC_B3_G(s2, s7, s8, s13, m[SC.s[r][12]], m[SC.s[r][13]]);
        LCPC_B3_G(s3, s4, s9, s14, m[SC.s[r][14]], m[SC.s[r][15]]);
    }
    }
    cv[0] = s0 ^ s8;  cv[1] = s1 ^ s9;  cv[2] = s2 ^ s10; cv[3] = s3 ^ s11;
    cv[4] = s4 ^ s12; cv[5] = s5 ^ s13; cv[6] = s6 ^ s14; cv[7] = s7 ^ s15;
}

// cv of a parent node (tree mode inside one hash): compress(IV, left || right, PARENT)
__device__ __forceinline__ void parent_cv(const uint32_t left[8], const uint32_t right[8], uint32_t extra_flags,
                                          uint32_t out[8]) {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { m[i] = left[i]; m[8 + i] = right[i]; }
    set_iv(out);
    compress(out, m, 0, BLOCK_BYTES, PARENT | extra_flags);
}

// BLAKE3 hash of a 64-byte message left || right (one chunk, one block, root): the
// Merkle-tree internal node of lcpc-2d/src/lib.rs:800-805.
template <bool LAT = false>
__device__ __forceinline__ void hash_pair(const uint32_t left[8], const uint32_t right[8], uint32_t out[8]) {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { m[i] = left[i]; m[8 + i] = right[i]; }
    set_iv(out);
    compress<LAT>(out, m, 0, BLOCK_BYTES, CHUNK_START | CHUNK_END | ROOT);
}

}  // namespace b3
}  // namespace lcpc
