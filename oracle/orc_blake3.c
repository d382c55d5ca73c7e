/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_blake3.h). */
#include "orc_blake3.h"
#include <string.h>

enum { F_CHUNK_START = 1, F_CHUNK_END = 2, F_PARENT = 4, F_ROOT = 8 };

static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                               0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

#define G(a, b, c, d, mx, my)                  \
    do {                                       \
        s[a] = s[a] + s[b] + (mx);             \
        s[d] = rotr(s[d] ^ s[a], 16);          \
        s[c] = s[c] + s[d];                    \
        s[b] = rotr(s[b] ^ s[c], 12);          \
        s[a] = s[a] + s[b] + (my);             \
        s[d] = rotr(s[d] ^ s[a], 8);           \
        s[c] = s[c] + s[d];                    \
        s[b] = rotr(s[b] ^ s[c], 7);           \
    } while (0)

/* compression function; writes the 8-word chaining value to out */
static void compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter,
                     uint32_t block_len, uint32_t flags, uint32_t out[8]) {
    uint32_t s[16], m[16], t[16];
    for (int i = 0; i < 8; i++) s[i] = cv[i];
    for (int i = 0; i < 4; i++) s[8 + i] = IV[i];
    s[12] = (uint32_t)counter;
    s[13] = (uint32_t)(counter >> 32);
    s[14] = block_len;
    s[15] = flags;
    memcpy(m, block, sizeof m);
    for (int r = 0; r < 7; r++) {
        G(0, 4, 8, 12, m[0], m[1]);
        G(1, 5, 9, 13, m[2], m[3]);
        G(2, 6, 10, 14, m[4], m[5]);
        G(3, 7, 11, 15, m[6], m[7]);
        G(0, 5, 10, 15, m[8], m[9]);
        G(1, 6, 11, 12, m[10], m[11]);
        G(2, 7, 8, 13, m[12], m[13]);
        G(3, 4, 9, 14, m[14], m[15]);
        if (r < 6) {
            for (int i = 0; i < 16; i++) t[i] = m[PERM[i]];
            memcpy(m, t, sizeof m);
        }
    }
    for (int i = 0; i < 8; i++) out[i] = s[i] ^ s[i + 8];
}

static void load_block(const uint8_t b[64], uint32_t w[16]) {
    for (int i = 0; i < 16; i++)
        w[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) |
               ((uint32_t)b[4 * i + 3] << 24);
}

static void parent_cv(const uint32_t l[8], const uint32_t r[8], uint32_t flags, uint32_t out[8]) {
    uint32_t blk[16];
    memcpy(blk, l, 32);
    memcpy(blk + 8, r, 32);
    compress(IV, blk, 0, ORC_B3_BLOCK, F_PARENT | flags, out);
}

static void chunk_reset(orc_b3_hasher *h, uint64_t counter) {
    memcpy(h->cv, IV, sizeof IV);
    h->chunk_counter = counter;
    memset(h->buf, 0, sizeof h->buf);
    h->buf_len = 0;
    h->blocks_compressed = 0;
}

void orc_b3_init(orc_b3_hasher *h) {
    chunk_reset(h, 0);
    h->cv_stack_len = 0;
}

static size_t chunk_len(const orc_b3_hasher *h) {
    return (size_t)h->blocks_compressed * ORC_B3_BLOCK + h->buf_len;
}

static uint32_t start_flag(const orc_b3_hasher *h) {
    return h->blocks_compressed == 0 ? F_CHUNK_START : 0;
}

static void chunk_update(orc_b3_hasher *h, const uint8_t *in, size_t len) {
    while (len > 0) {
        if (h->buf_len == ORC_B3_BLOCK) {
            uint32_t w[16];
            load_block(h->buf, w);
            compress(h->cv, w, h->chunk_counter, ORC_B3_BLOCK, start_flag(h), h->cv);
            h->blocks_compressed++;
            h->buf_len = 0;
            memset(h->buf, 0, sizeof h->buf);
        }
        size_t want = ORC_B3_BLOCK - h->buf_len;
        size_t take = len < want ? len : want;
        memcpy(h->buf + h->buf_len, in, take);
        h->buf_len += (uint8_t)take;
        in += take;
        len -= take;
    }
}

/* chaining value of the (complete) chunk in progress, not a root */
static void chunk_cv(const orc_b3_hasher *h, uint32_t out[8]) {
    uint32_t w[16];
    load_block(h->buf, w);
    compress(h->cv, w, h->chunk_counter, h->buf_len, start_flag(h) | F_CHUNK_END, out);
}

void orc_b3_update(orc_b3_hasher *h, const void *data, size_t len) {
    const uint8_t *in = (const uint8_t *)data;
    while (len > 0) {
        if (chunk_len(h) == ORC_B3_CHUNK) {
            uint32_t cv[8];
            chunk_cv(h, cv);
            uint64_t total = h->chunk_counter + 1;
            /* merge completed subtrees: one per trailing zero bit of `total` */
            uint64_t t = total;
            while ((t & 1) == 0) {
                h->cv_stack_len--;
                parent_cv(h->cv_stack[h->cv_stack_len], cv, 0, cv);
                t >>= 1;
            }
            memcpy(h->cv_stack[h->cv_stack_len++], cv, 32);
            chunk_reset(h, total);
        }
        size_t want = ORC_B3_CHUNK - chunk_len(h);
        size_t take = len < want ? len : want;
        chunk_update(h, in, take);
        in += take;
        len -= take;
    }
}

void orc_b3_finalize(const orc_b3_hasher *h, uint8_t out[ORC_B3_OUT]) {
    uint32_t cv[8], w[16];
    if (h->cv_stack_len == 0) {
        /* single chunk: its last block is the root */
        load_block(h->buf, w);
        compress(h->cv, w, 0, h->buf_len, start_flag(h) | F_CHUNK_END | F_ROOT, cv);
    } else {
        chunk_cv(h, cv);
        int n = h->cv_stack_len;
        while (n > 1) {
            parent_cv(h->cv_stack[n - 1], cv, 0, cv);
            n--;
        }
        parent_cv(h->cv_stack[0], cv, F_ROOT, cv);
    }
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)cv[i];
        out[4 * i + 1] = (uint8_t)(cv[i] >> 8);
        out[4 * i + 2] = (uint8_t)(cv[i] >> 16);
        out[4 * i + 3] = (uint8_t)(cv[i] >> 24);
    }
}

void orc_blake3(const void *data, size_t len, uint8_t out[ORC_B3_OUT]) {
    orc_b3_hasher h;
    orc_b3_init(&h);
    orc_b3_update(&h, data, len);
    orc_b3_finalize(&h, out);
}

/* ---- pieces of the tree, for checking a row-sharded hash (one rank computes the chaining values of the chunks whose
 * bytes it owns, another joins them): the public BLAKE3 spec's chunk chaining value and parent tree ------------- */

static void cv_bytes(const uint32_t cv[8], uint8_t out[32]) {
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)cv[i];
        out[4 * i + 1] = (uint8_t)(cv[i] >> 8);
        out[4 * i + 2] = (uint8_t)(cv[i] >> 16);
        out[4 * i + 3] = (uint8_t)(cv[i] >> 24);
    }
}

/* chaining value (not a root) of chunk number `chunk_index` holding `len` <= 1024 bytes */
void orc_b3_chunk_cv(const void *data, size_t len, uint64_t chunk_index, uint8_t out[32]) {
    orc_b3_hasher h;
    uint32_t cv[8];
    chunk_reset(&h, chunk_index);
    chunk_update(&h, (const uint8_t *)data, len);
    chunk_cv(&h, cv);
    cv_bytes(cv, out);
}

/* root hash from the chaining values of chunks 0 .. n-1, n >= 2 (the stack discipline of orc_b3_update / finalize) */
void orc_b3_merge_cvs(const uint8_t *cvs, size_t n, uint8_t out[32]) {
    uint32_t stack[ORC_B3_MAX_DEPTH][8], cv[8];
    int sp = 0;
    for (size_t c = 0; c < n; c++) {
        for (int i = 0; i < 8; i++) {
            const uint8_t *b = cvs + 32 * c + 4 * i;
            cv[i] = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
        }
        if (c + 1 == n) break;
        uint64_t t = c + 1;
        while ((t & 1) == 0) {
            sp--;
            parent_cv(stack[sp], cv, 0, cv);
            t >>= 1;
        }
        memcpy(stack[sp++], cv, 32);
    }
    while (sp > 1) {
        parent_cv(stack[sp - 1], cv, 0, cv);
        sp--;
    }
    parent_cv(stack[0], cv, F_ROOT, cv);
    cv_bytes(cv, out);
}
