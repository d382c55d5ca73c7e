/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Host-side randomness the reference consumes on this path, restated from the
 * published crates because none of them is vendored under /root/reference:
 *   - rand_chacha 0.3 `ChaCha20Rng` (Cargo.toml:33): from_seed, seed_from_u64
 *     (rand_core 0.6 PCG32 expansion), set_stream, next_u32/next_u64.
 *     Call sites: lcpc-2d/src/lib.rs:902,935,1058,1105;
 *     lcpc-brakedown-pc/src/matgen.rs:43-44.
 *   - rand 0.8 `Uniform::<usize>::new(0, n).sample` (lib.rs:937-940,1107-1110;
 *     matgen.rs:119,147-158): widening-multiply rejection sampler.
 *   - merlin 2.0 `Transcript` (Cargo.toml:25) = STROBE-128 over Keccak-f[1600]
 *     (lib.rs:49,901,934,1057,1104).
 * Pinned in tests/test_oracle_rand.py: ChaCha20 keystream against the
 * `cryptography` package, Keccak-f against hashlib's SHA3-256, and merlin's
 * own published "simple transcript" known-answer vector.
 */
#ifndef ORC_RAND_H
#define ORC_RAND_H
#include <stddef.h>
#include <stdint.h>

typedef struct {
    uint32_t key[8];
    uint64_t counter; /* 64-bit block counter (state words 12,13) */
    uint64_t stream;  /* 64-bit stream id     (state words 14,15) */
    uint32_t buf[16];
    int idx;          /* next unread word in buf; 16 = empty      */
    int rounds;       /* 20 = ChaCha20Rng, 8 = ChaCha8Rng; 0 is read as 20 */
} orc_chacha_rng;

void orc_chacha_from_seed(orc_chacha_rng *r, const uint8_t seed[32]);
void orc_chacha_seed_from_u64(orc_chacha_rng *r, uint64_t state);
void orc_chacha_set_stream(orc_chacha_rng *r, uint64_t stream);
uint32_t orc_chacha_next_u32(orc_chacha_rng *r);
uint64_t orc_chacha_next_u64(orc_chacha_rng *r);
/* rand 0.8 Uniform::new(0, n).sample(rng) for usize on a 64-bit target */
uint64_t orc_uniform_usize(orc_chacha_rng *r, uint64_t n);
/* rand 0.8 rng.gen_range(0..n) for u32 (UniformInt::sample_single) */
uint32_t orc_gen_range_u32(orc_chacha_rng *r, uint32_t n);
/* ChaCha8Rng::seed_from_u64(seed) + IteratorRandom::choose_multiple over 0..max_index
 * (proof-of-storage/src/networking/client.rs:443-456); returns the number of indices written */
size_t orc_pos_choose_columns(uint64_t seed, size_t amount, size_t max_index, uint64_t *out);

void orc_keccak_f1600(uint64_t st[25]);

typedef struct {
    uint8_t state[200];
    uint8_t pos, pos_begin, cur_flags;
} orc_transcript;

void orc_transcript_new(orc_transcript *t, const uint8_t *label, size_t label_len);
void orc_transcript_append_message(orc_transcript *t, const uint8_t *label, size_t label_len,
                                   const uint8_t *msg, size_t msg_len);
void orc_transcript_challenge_bytes(orc_transcript *t, const uint8_t *label, size_t label_len,
                                    uint8_t *dest, size_t dest_len);

#endif
