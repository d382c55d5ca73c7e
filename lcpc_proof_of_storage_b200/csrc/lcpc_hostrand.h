// Host-side randomness and transcript of the scheme (C++).  In the reference these stay
// on the CPU as well: merlin::Transcript (lcpc-2d/src/lib.rs:49,901,934,1057,1104),
// ChaCha20Rng + F::random for the degree-test tensors (:902-907,:1058-1062),
// Uniform(0, n_cols) for the column indices (:935-940,:1105-1110), and the seeded code
// generation of lcpc-brakedown-pc/src/matgen.rs.  No field arithmetic happens here: the
// only element-level operations are comparisons against the modulus (rejection sampling).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

namespace lcpc {
namespace host {

// ---- Keccak-f[1600] / STROBE-128 / merlin ------------------------------------------
void keccak_f1600(uint64_t st[25]);

class Transcript {
public:
    explicit Transcript(const uint8_t *label, size_t len);
    void append_message(const uint8_t *label, size_t label_len, const uint8_t *msg, size_t msg_len);
    void challenge_bytes(const uint8_t *label, size_t label_len, uint8_t *dest, size_t dest_len);

private:
    enum : uint8_t { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
    static constexpr int STROBE_R = 166;
    uint8_t state_[200];
    uint8_t pos_ = 0, pos_begin_ = 0, cur_flags_ = 0;
    void permute();
    void run_f();
    void absorb(const uint8_t *d, size_t n);
    void squeeze(uint8_t *d, size_t n);
    void begin_op(uint8_t flags, bool more);
    void meta_ad(const uint8_t *d, size_t n, bool more);
    void ad(const uint8_t *d, size_t n, bool more);
    void prf(uint8_t *d, size_t n, bool more);
};

// ---- rand_chacha 0.3 ChaCha20Rng ----------------------------------------------------
// rand_chacha 0.3 ChaChaXRng; the round count distinguishes ChaCha20Rng (prove / verify / matgen)
// from ChaCha8Rng (proof-of-storage column choice, networking/client.rs:448).
class ChaCha20Rng {
public:
    static ChaCha20Rng from_seed(const uint8_t seed[32], int rounds = 20);
    static ChaCha20Rng seed_from_u64(uint64_t state, int rounds = 20);  // rand_core 0.6 PCG32 expansion
    void set_stream(uint64_t stream) { stream_ = stream; }  // before the first draw (matgen.rs:43-44)
    uint32_t next_u32();
    uint64_t next_u64();
    // rand 0.8 Uniform::new(0usize, n).sample(rng), 64-bit target
    uint64_t uniform(uint64_t n);
    // rand 0.8 rng.gen_range(0..n) for u32 (UniformInt::sample_single: leading-zeros zone)
    uint32_t gen_range_u32(uint32_t n);

private:
    uint32_t key_[8];
    uint64_t counter_ = 0, stream_ = 0;
    uint32_t buf_[16];
    int idx_ = 16;
    int rounds_ = 20;
    void refill();
};

// ff_derive `Field::random`: LIMBS x next_u64, top limb masked to NUM_BITS, rejected if >= p;
// the accepted limbs are the Montgomery residue.
void field_random(int fid, ChaCha20Rng &rng, uint64_t *out);

// rand 0.8 IteratorRandom::choose_multiple over 0..max (reservoir sampling; gen_index draws u32 when
// the bound fits): proof-of-storage get_column_indicies_from_random_seed (networking/client.rs:443-456)
void choose_multiple_indices(ChaCha20Rng &rng, uint64_t amount, uint64_t max_index, std::vector<uint64_t> &out);

// ---- lcpc-brakedown-pc code generation -------------------------------------------------
struct SdigDims {
    uint64_t n, m, d;  // columns, rows, non-zeros per column
};
// codespec.rs:168-232; code in 1..6.  Returns false for n <= baselen.
bool sdig_get_dims(int code, uint64_t n, double log2p, std::vector<SdigDims> &pre, std::vector<SdigDims> &post);
double sdig_dist(int code);
// matgen.rs:114-188 gen_code into caller-provided CSC arrays (indptr[n+1], indices[n*d], data[n*d*LIMBS])
void sdig_gen_code(int fid, ChaCha20Rng &rng, const SdigDims &dim, uint64_t *indptr, uint64_t *indices, uint64_t *data);

}  // namespace host
}  // namespace lcpc
