/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the
 * shipped product; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call it.
 *
 * Portable BLAKE3 (hash mode, 32-byte output), written from the public BLAKE3
 * specification.  The reference uses the `blake3` crate (v1.5, workspace
 * Cargo.toml:8) through `digest::Digest`; the crate source is not vendored in
 * /root/reference, so this restates the published algorithm and is pinned in
 * tests/test_oracle_blake3.py against the Python `blake3` package (bindings
 * to the same Rust crate) for every length class the path produces.
 */
#ifndef ORC_BLAKE3_H
#define ORC_BLAKE3_H
#include <stddef.h>
#include <stdint.h>

#define ORC_B3_OUT 32
#define ORC_B3_BLOCK 64
#define ORC_B3_CHUNK 1024
#define ORC_B3_MAX_DEPTH 54

typedef struct {
    uint32_t cv[8];          /* chaining value of the chunk in progress      */
    uint64_t chunk_counter;  /* index of the chunk in progress               */
    uint8_t buf[ORC_B3_BLOCK];
    uint8_t buf_len;
    uint8_t blocks_compressed;
    uint32_t cv_stack[ORC_B3_MAX_DEPTH][8];
    uint8_t cv_stack_len;
} orc_b3_hasher;

void orc_b3_init(orc_b3_hasher *h);
void orc_b3_update(orc_b3_hasher *h, const void *data, size_t len);
void orc_b3_finalize(const orc_b3_hasher *h, uint8_t out[ORC_B3_OUT]);
/* one-shot */
void orc_blake3(const void *data, size_t len, uint8_t out[ORC_B3_OUT]);
/* pieces of the tree (row-sharded hashing): non-root chaining value of one chunk; root from n >= 2 chunk values */
void orc_b3_chunk_cv(const void *data, size_t len, uint64_t chunk_index, uint8_t out[32]);
void orc_b3_merge_cvs(const uint8_t *cvs, size_t n, uint8_t out[32]);

#endif
