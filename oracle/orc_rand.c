/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_rand.h). */
#include "orc_rand.h"
#include <string.h>

/* ---------------- ChaCha20 (20 rounds, 64-bit counter + 64-bit stream) ---------------- */

static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

#define QR(a, b, c, d)                     \
    do {                                   \
        a += b; d ^= a; d = rotl32(d, 16); \
        c += d; b ^= c; b = rotl32(b, 12); \
        a += b; d ^= a; d = rotl32(d, 8);  \
        c += d; b ^= c; b = rotl32(b, 7);  \
    } while (0)

static void chacha_block(orc_chacha_rng *r) {
    uint32_t in[16], x[16];
    in[0] = 0x61707865u; in[1] = 0x3320646eu; in[2] = 0x79622d32u; in[3] = 0x6b206574u;
    for (int i = 0; i < 8; i++) in[4 + i] = r->key[i];
    in[12] = (uint32_t)r->counter;
    in[13] = (uint32_t)(r->counter >> 32);
    in[14] = (uint32_t)r->stream;
    in[15] = (uint32_t)(r->stream >> 32);
    memcpy(x, in, sizeof x);
    int dr = (r->rounds ? r->rounds : 20) / 2;
    for (int i = 0; i < dr; i++) {
        QR(x[0], x[4], x[8], x[12]);
        QR(x[1], x[5], x[9], x[13]);
        QR(x[2], x[6], x[10], x[14]);
        QR(x[3], x[7], x[11], x[15]);
        QR(x[0], x[5], x[10], x[15]);
        QR(x[1], x[6], x[11], x[12]);
        QR(x[2], x[7], x[8], x[13]);
        QR(x[3], x[4], x[9], x[14]);
    }
    for (int i = 0; i < 16; i++) r->buf[i] = x[i] + in[i];
    r->counter++;
    r->idx = 0;
}

void orc_chacha_from_seed(orc_chacha_rng *r, const uint8_t seed[32]) {
    for (int i = 0; i < 8; i++)
        r->key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) |
                    ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
    r->counter = 0;
    r->stream = 0;
    r->idx = 16;
    r->rounds = 20;
}

/* rand_core 0.6 SeedableRng::seed_from_u64 default: PCG32 fills the seed 4 bytes at a time */
void orc_chacha_seed_from_u64(orc_chacha_rng *r, uint64_t state) {
    const uint64_t MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
    uint8_t seed[32];
    for (int i = 0; i < 8; i++) {
        state = state * MUL + INC;
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
        seed[4 * i] = (uint8_t)x;
        seed[4 * i + 1] = (uint8_t)(x >> 8);
        seed[4 * i + 2] = (uint8_t)(x >> 16);
        seed[4 * i + 3] = (uint8_t)(x >> 24);
    }
    orc_chacha_from_seed(r, seed);
}

/* only valid before the first draw (the only way the reference uses it, matgen.rs:43-44) */
void orc_chacha_set_stream(orc_chacha_rng *r, uint64_t stream) { r->stream = stream; }

uint32_t orc_chacha_next_u32(orc_chacha_rng *r) {
    if (r->idx >= 16) chacha_block(r);
    return r->buf[r->idx++];
}

/* rand_core BlockRng::next_u64: two consecutive words of the stream, low word first */
uint64_t orc_chacha_next_u64(orc_chacha_rng *r) {
    uint64_t lo = orc_chacha_next_u32(r);
    uint64_t hi = orc_chacha_next_u32(r);
    return (hi << 32) | lo;
}

uint64_t orc_uniform_usize(orc_chacha_rng *r, uint64_t n) {
    /* UniformInt::new(0, n) -> new_inclusive(0, n-1): range = n */
    uint64_t range = n;
    uint64_t ints_to_reject = (UINT64_MAX - range + 1) % range;
    uint64_t zone = UINT64_MAX - ints_to_reject;
    for (;;) {
        uint64_t v = orc_chacha_next_u64(r);
        unsigned __int128 m = (unsigned __int128)v * range;
        uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
        if (lo <= zone) return hi;
    }
}

uint32_t orc_gen_range_u32(orc_chacha_rng *r, uint32_t n) {
    uint32_t zone = (n << __builtin_clz(n)) - 1;
    for (;;) {
        uint64_t m = (uint64_t)orc_chacha_next_u32(r) * n;
        if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
    }
}

size_t orc_pos_choose_columns(uint64_t seed, size_t amount, size_t max_index, uint64_t *out) {
    orc_chacha_rng rng;
    orc_chacha_seed_from_u64(&rng, seed);
    rng.rounds = 8;
    size_t n = 0;
    for (size_t i = 0; i < amount && i < max_index; i++) out[n++] = i; /* reservoir.extend(take(amount)) */
    if (n < amount) return n;
    for (size_t i = 0; amount + i < max_index; i++) { /* gen_index(rng, i + 1 + amount) */
        uint64_t ub = i + 1 + amount;
        uint64_t k = ub <= 0xffffffffull ? orc_gen_range_u32(&rng, (uint32_t)ub) : orc_uniform_usize(&rng, ub);
        if (k < amount) out[k] = amount + i;
    }
    return n;
}

/* ---------------- Keccak-f[1600] ---------------- */

static const uint64_t KRC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
    0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
    0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
    0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
    0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
static const int KROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14,
                             27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KPIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4,
                             15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};

static inline uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

void orc_keccak_f1600(uint64_t st[25]) {
    uint64_t bc[5], t;
    for (int round = 0; round < 24; round++) {
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            t = bc[(i + 4) % 5] ^ rotl64(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        t = st[1];
        for (int i = 0; i < 24; i++) {
            int j = KPIL[i];
            bc[0] = st[j];
            st[j] = rotl64(t, KROT[i]);
            t = bc[0];
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= KRC[round];
    }
}

/* ---------------- STROBE-128 as used by merlin ---------------- */

enum { STROBE_R = 166, FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };

static void strobe_permute(orc_transcript *t) {
    uint64_t lanes[25];
    for (int i = 0; i < 25; i++) {
        uint64_t v = 0;
        for (int b = 7; b >= 0; b--) v = (v << 8) | t->state[8 * i + b];
        lanes[i] = v;
    }
    orc_keccak_f1600(lanes);
    for (int i = 0; i < 25; i++)
        for (int b = 0; b < 8; b++) t->state[8 * i + b] = (uint8_t)(lanes[i] >> (8 * b));
}

static void strobe_run_f(orc_transcript *t) {
    t->state[t->pos] ^= t->pos_begin;
    t->state[t->pos + 1] ^= 0x04;
    t->state[STROBE_R + 1] ^= 0x80;
    strobe_permute(t);
    t->pos = 0;
    t->pos_begin = 0;
}

static void strobe_absorb(orc_transcript *t, const uint8_t *data, size_t len) {
    for (size_t i = 0; i < len; i++) {
        t->state[t->pos] ^= data[i];
        t->pos++;
        if (t->pos == STROBE_R) strobe_run_f(t);
    }
}

static void strobe_squeeze(orc_transcript *t, uint8_t *data, size_t len) {
    for (size_t i = 0; i < len; i++) {
        data[i] = t->state[t->pos];
        t->state[t->pos] = 0;
        t->pos++;
        if (t->pos == STROBE_R) strobe_run_f(t);
    }
}

static void strobe_begin_op(orc_transcript *t, uint8_t flags, int more) {
    if (more) return; /* continuation of the same operation */
    uint8_t old_begin = t->pos_begin;
    t->pos_begin = t->pos + 1;
    t->cur_flags = flags;
    uint8_t hdr[2] = {old_begin, flags};
    strobe_absorb(t, hdr, 2);
    if ((flags & (FLAG_C | FLAG_K)) && t->pos != 0) strobe_run_f(t);
}

static void strobe_meta_ad(orc_transcript *t, const uint8_t *d, size_t n, int more) {
    strobe_begin_op(t, FLAG_M | FLAG_A, more);
    strobe_absorb(t, d, n);
}

static void strobe_ad(orc_transcript *t, const uint8_t *d, size_t n, int more) {
    strobe_begin_op(t, FLAG_A, more);
    strobe_absorb(t, d, n);
}

static void strobe_prf(orc_transcript *t, uint8_t *d, size_t n, int more) {
    strobe_begin_op(t, FLAG_I | FLAG_A | FLAG_C, more);
    strobe_squeeze(t, d, n);
}

static void le32(uint8_t out[4], size_t v) {
    out[0] = (uint8_t)v; out[1] = (uint8_t)(v >> 8); out[2] = (uint8_t)(v >> 16); out[3] = (uint8_t)(v >> 24);
}

void orc_transcript_append_message(orc_transcript *t, const uint8_t *label, size_t label_len,
                                   const uint8_t *msg, size_t msg_len) {
    uint8_t len[4];
    le32(len, msg_len);
    strobe_meta_ad(t, label, label_len, 0);
    strobe_meta_ad(t, len, 4, 1);
    strobe_ad(t, msg, msg_len, 0);
}

void orc_transcript_challenge_bytes(orc_transcript *t, const uint8_t *label, size_t label_len,
                                    uint8_t *dest, size_t dest_len) {
    uint8_t len[4];
    le32(len, dest_len);
    strobe_meta_ad(t, label, label_len, 0);
    strobe_meta_ad(t, len, 4, 1);
    strobe_prf(t, dest, dest_len, 0);
}

void orc_transcript_new(orc_transcript *t, const uint8_t *label, size_t label_len) {
    static const uint8_t init[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
    memset(t, 0, sizeof *t);
    memcpy(t->state, init, 6);
    memcpy(t->state + 6, "STROBEv1.0.2", 12);
    strobe_permute(t);
    strobe_meta_ad(t, (const uint8_t *)"Merlin v1.0", 11, 0);
    orc_transcript_append_message(t, (const uint8_t *)"dom-sep", 7, label, label_len);
}
