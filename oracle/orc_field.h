/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_blake3.h for the rule).
 *
 * Prime-field descriptors for the reference's fields.  The reference derives
 * its arithmetic with ff_derive 0.13 (`#[derive(PrimeField)]`,
 * lcpc-test-fields/src/lib.rs:18-70, proof-of-storage/src/fields/
 * writable_ft63.rs:8-12): an element is `[u64; LIMBS]`, least-significant limb
 * first, holding a*R mod p with R = 2^(64*LIMBS) (Montgomery form), always
 * fully reduced; `to_repr()` is the canonical value as 8*LIMBS little-endian
 * bytes (big-endian for Ft253_192, proof-of-storage/src/fields/ft253_192.rs:6-10).  ff_derive is a crates.io dependency (Cargo.toml:15, "0.13"), not in
 * /root/reference; the constants below are recomputed from the moduli and
 * generators in the derive attributes (tests/test_oracle_field.py re-derives
 * every one of them with Python integers).
 */
#ifndef ORC_FIELD_H
#define ORC_FIELD_H
#include <stdint.h>

#define ORC_MAX_LIMBS 4

enum { ORC_FT63 = 0, ORC_FT127 = 1, ORC_FT191 = 2, ORC_FT255 = 3, ORC_FT253_192 = 4, ORC_N_FIELDS = 5 };

typedef struct {
    int limbs;                    /* number of 64-bit limbs                         */
    int num_bits;                 /* PrimeField::NUM_BITS                           */
    int s;                        /* 2-adicity: p - 1 = 2^s * t, t odd              */
    uint64_t p[ORC_MAX_LIMBS];    /* modulus                                        */
    uint64_t inv;                 /* -p^{-1} mod 2^64                               */
    uint64_t r[ORC_MAX_LIMBS];    /* R mod p   (= F::ONE)                           */
    uint64_t r2[ORC_MAX_LIMBS];   /* R^2 mod p                                      */
    uint64_t root[ORC_MAX_LIMBS]; /* ROOT_OF_UNITY = GENERATOR^t, Montgomery form   */
    uint64_t top_mask;            /* 0xff..ff >> REPR_SHAVE_BITS, for F::random     */
    int repr_big_endian;          /* PrimeFieldReprEndianness = "big" (Ft253_192)   */
} orc_field;

const orc_field *orc_get_field(int fid);

#endif
