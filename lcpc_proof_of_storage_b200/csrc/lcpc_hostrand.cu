// Host-side transcript / PRNG / code generation (see lcpc_hostrand.h).  Restated from the
// published algorithms of the crates the reference depends on (merlin 2.0, rand_chacha 0.3,
// rand 0.8, ff_derive 0.13); none of them is vendored in the reference tree.
#include "lcpc_hostrand.h"

#include <algorithm>
#include <cmath>

#include "lcpc_field.cuh"

namespace lcpc {
namespace host {

// ------------------------------------------------------------------ Keccak-f[1600]

static inline uint64_t rol64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

// The transcript is the host-side cost of prove / verify (every element of every fold result is one
// append_message, 22 absorbed bytes: 4300 permutations per 2^24 fold), so the permutation is written with the 25
// lanes in locals and every index a literal -- theta, rho+pi, chi, iota per round -- instead of table-driven loops.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
// a second clone for CPUs with ANDN / RORX (chosen at load time): 0.92 -> 0.68 us per permutation here
__attribute__((target_clones("default", "arch=x86-64-v3"), optimize("no-tree-vectorize")))
#endif
void keccak_f1600(uint64_t a[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
        0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
        0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
        0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
        0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    uint64_t a00 = a[0], a01 = a[1], a02 = a[2], a03 = a[3], a04 = a[4], a05 = a[5], a06 = a[6], a07 = a[7], a08 = a[8],
             a09 = a[9], a10 = a[10], a11 = a[11], a12 = a[12], a13 = a[13], a14 = a[14], a15 = a[15], a16 = a[16],
             a17 = a[17], a18 = a[18], a19 = a[19], a20 = a[20], a21 = a[21], a22 = a[22], a23 = a[23], a24 = a[24];
    for (int round = 0; round < 24; round++) {
        // theta
        const uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20, c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21, c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22,
                       c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23, c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
        const uint64_t d0 = c4 ^ rol64(c1, 1), d1 = c0 ^ rol64(c2, 1), d2 = c1 ^ rol64(c3, 1), d3 = c2 ^ rol64(c4, 1),
                       d4 = c3 ^ rol64(c0, 1);
        a00 ^= d0; a05 ^= d0; a10 ^= d0; a15 ^= d0; a20 ^= d0;
        a01 ^= d1; a06 ^= d1; a11 ^= d1; a16 ^= d1; a21 ^= d1;
        a02 ^= d2; a07 ^= d2; a12 ^= d2; a17 ^= d2; a22 ^= d2;
        a03 ^= d3; a08 ^= d3; a13 ^= d3; a18 ^= d3; a23 ^= d3;
        a04 ^= d4; a09 ^= d4; a14 ^= d4; a19 ^= d4; a24 ^= d4;
        // rho + pi: b[y + 5*((2x + 3y) mod 5)] = rol(a[x + 5y], r[x][y])
        const uint64_t b00 = a00, b10 = rol64(a01, 1), b20 = rol64(a02, 62), b05 = rol64(a03, 28), b15 = rol64(a04, 27);
        const uint64_t b16 = rol64(a05, 36), b01 = rol64(a06, 44), b11 = rol64(a07, 6), b21 = rol64(a08, 55), b06 = rol64(a09, 20);
        const uint64_t b07 = rol64(a10, 3), b17 = rol64(a11, 10), b02 = rol64(a12, 43), b12 = rol64(a13, 25), b22 = rol64(a14, 39);
        const uint64_t b23 = rol64(a15, 41), b08 = rol64(a16, 45), b18 = rol64(a17, 15), b03 = rol64(a18, 21), b13 = rol64(a19, 8);
        const uint64_t b14 = rol64(a20, 18), b24 = rol64(a21, 2), b09 = rol64(a22, 61), b19 = rol64(a23, 56), b04 = rol64(a24, 14);
        // chi (+ iota on lane 0)
        a00 = b00 ^ (~b01 & b02) ^ RC[round]; a01 = b01 ^ (~b02 & b03); a02 = b02 ^ (~b03 & b04); a03 = b03 ^ (~b04 & b00); a04 = b04 ^ (~b00 & b01);
        a05 = b05 ^ (~b06 & b07); a06 = b06 ^ (~b07 & b08); a07 = b07 ^ (~b08 & b09); a08 = b08 ^ (~b09 & b05); a09 = b09 ^ (~b05 & b06);
        a10 = b10 ^ (~b11 & b12); a11 = b11 ^ (~b12 & b13); a12 = b12 ^ (~b13 & b14); a13 = b13 ^ (~b14 & b10); a14 = b14 ^ (~b10 & b11);
        a15 = b15 ^ (~b16 & b17); a16 = b16 ^ (~b17 & b18); a17 = b17 ^ (~b18 & b19); a18 = b18 ^ (~b19 & b15); a19 = b19 ^ (~b15 & b16);
        a20 = b20 ^ (~b21 & b22); a21 = b21 ^ (~b22 & b23); a22 = b22 ^ (~b23 & b24); a23 = b23 ^ (~b24 & b20); a24 = b24 ^ (~b20 & b21);
    }
    a[0] = a00; a[1] = a01; a[2] = a02; a[3] = a03; a[4] = a04; a[5] = a05; a[6] = a06; a[7] = a07; a[8] = a08; a[9] = a09;
    a[10] = a10; a[11] = a11; a[12] = a12; a[13] = a13; a[14] = a14; a[15] = a15; a[16] = a16; a[17] = a17; a[18] = a18;
    a[19] = a19; a[20] = a20; a[21] = a21; a[22] = a22; a[23] = a23; a[24] = a24;
}

// ------------------------------------------------------------------ merlin over STROBE-128

void Transcript::permute() {
    uint64_t lanes[25];
    const uint16_t probe = 1;
    if (*reinterpret_cast<const uint8_t *>(&probe) == 1) {  // little-endian host: the byte state is the lane array
        std::memcpy(lanes, state_, sizeof lanes);
        keccak_f1600(lanes);
        std::memcpy(state_, lanes, sizeof lanes);
        return;
    }
    for (int i = 0; i < 25; i++) {
        uint64_t v = 0;
        for (int b = 7; b >= 0; b--) v = (v << 8) | state_[8 * i + b];
        lanes[i] = v;
    }
    keccak_f1600(lanes);
    for (int i = 0; i < 25; i++)
        for (int b = 0; b < 8; b++) state_[8 * i + b] = (uint8_t)(lanes[i] >> (8 * b));
}

void Transcript::run_f() {
    state_[pos_] ^= pos_begin_;
    state_[pos_ + 1] ^= 0x04;
    state_[STROBE_R + 1] ^= 0x80;
    permute();
    pos_ = 0;
    pos_begin_ = 0;
}

void Transcript::absorb(const uint8_t *d, size_t n) {
    while (n) {
        size_t take = (size_t)(STROBE_R - pos_);
        if (take > n) take = n;
        uint8_t *s = state_ + pos_;
        for (size_t i = 0; i < take; i++) s[i] ^= d[i];
        pos_ = (uint8_t)(pos_ + take);
        d += take;
        n -= take;
        if (pos_ == STROBE_R) run_f();
    }
}

void Transcript::squeeze(uint8_t *d, size_t n) {
    for (size_t i = 0; i < n; i++) {
        d[i] = state_[pos_];
        state_[pos_++] = 0;
        if (pos_ == STROBE_R) run_f();
    }
}

void Transcript::begin_op(uint8_t flags, bool more) {
    if (more) return;  // continuation of the operation in progress
    const uint8_t old_begin = pos_begin_;
    pos_begin_ = pos_ + 1;
    cur_flags_ = flags;
    const uint8_t hdr[2] = {old_begin, flags};
    absorb(hdr, 2);
    if ((flags & (FLAG_C | FLAG_K)) && pos_ != 0) run_f();
}

void Transcript::meta_ad(const uint8_t *d, size_t n, bool more) {
    begin_op(FLAG_M | FLAG_A, more);
    absorb(d, n);
}
void Transcript::ad(const uint8_t *d, size_t n, bool more) {
    begin_op(FLAG_A, more);
    absorb(d, n);
}
void Transcript::prf(uint8_t *d, size_t n, bool more) {
    begin_op(FLAG_I | FLAG_A | FLAG_C, more);
    squeeze(d, n);
}

static void le32(uint8_t o[4], size_t v) {
    for (int i = 0; i < 4; i++) o[i] = (uint8_t)(v >> (8 * i));
}

void Transcript::append_message(const uint8_t *label, size_t label_len, const uint8_t *msg, size_t msg_len) {
    uint8_t len[4];
    le32(len, msg_len);
    meta_ad(label, label_len, false);
    meta_ad(len, 4, true);
    ad(msg, msg_len, false);
}

void Transcript::challenge_bytes(const uint8_t *label, size_t label_len, uint8_t *dest, size_t dest_len) {
    uint8_t len[4];
    le32(len, dest_len);
    meta_ad(label, label_len, false);
    meta_ad(len, 4, true);
    prf(dest, dest_len, false);
}

Transcript::Transcript(const uint8_t *label, size_t len) {
    std::memset(state_, 0, sizeof state_);
    const uint8_t init[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
    std::memcpy(state_, init, 6);
    std::memcpy(state_ + 6, "STROBEv1.0.2", 12);
    permute();
    meta_ad(reinterpret_cast<const uint8_t *>("Merlin v1.0"), 11, false);
    append_message(reinterpret_cast<const uint8_t *>("dom-sep"), 7, label, len);
}

// ------------------------------------------------------------------ ChaCha20Rng

static inline uint32_t rol32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

void ChaCha20Rng::refill() {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; i++) in[4 + i] = key_[i];
    in[12] = (uint32_t)counter_;
    in[13] = (uint32_t)(counter_ >> 32);
    in[14] = (uint32_t)stream_;
    in[15] = (uint32_t)(stream_ >> 32);
    uint32_t x[16];
    std::memcpy(x, in, sizeof x);
    auto qr = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rol32(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rol32(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rol32(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rol32(x[b] ^ x[c], 7);
    };
    for (int i = 0; i < rounds_ / 2; i++) {
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) buf_[i] = x[i] + in[i];
    counter_++;
    idx_ = 0;
}

ChaCha20Rng ChaCha20Rng::from_seed(const uint8_t seed[32], int rounds) {
    ChaCha20Rng r;
    r.rounds_ = rounds;
    for (int i = 0; i < 8; i++)
        r.key_[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) |
                    ((uint32_t)seed[4 * i + 3] << 24);
    return r;
}

ChaCha20Rng ChaCha20Rng::seed_from_u64(uint64_t state, int rounds) {
    uint8_t seed[32];
    for (int i = 0; i < 8; i++) {
        state = state * 6364136223846793005ull + 11634580027462260723ull;
        const uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
        const uint32_t rot = (uint32_t)(state >> 59);
        const uint32_t v = (xs >> rot) | (xs << ((32 - rot) & 31));
        for (int b = 0; b < 4; b++) seed[4 * i + b] = (uint8_t)(v >> (8 * b));
    }
    return from_seed(seed, rounds);
}

uint32_t ChaCha20Rng::next_u32() {
    if (idx_ >= 16) refill();
    return buf_[idx_++];
}

uint64_t ChaCha20Rng::next_u64() {
    const uint64_t lo = next_u32();
    const uint64_t hi = next_u32();
    return (hi << 32) | lo;
}

uint64_t ChaCha20Rng::uniform(uint64_t n) {
    const uint64_t reject = (UINT64_MAX - n + 1) % n;
    const uint64_t zone = UINT64_MAX - reject;
    for (;;) {
        const unsigned __int128 m = (unsigned __int128)next_u64() * n;
        if ((uint64_t)m <= zone) return (uint64_t)(m >> 64);
    }
}

uint32_t ChaCha20Rng::gen_range_u32(uint32_t n) {
    // UniformInt<u32>::sample_single_inclusive(0, n-1): range = n, zone = (range << lz) - 1
    const uint32_t zone = (n << __builtin_clz(n)) - 1;
    for (;;) {
        const uint64_t m = (uint64_t)next_u32() * n;
        if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
    }
}

void choose_multiple_indices(ChaCha20Rng &rng, uint64_t amount, uint64_t max_index, std::vector<uint64_t> &out) {
    out.clear();
    for (uint64_t i = 0; i < amount && i < max_index; i++) out.push_back(i);
    if (out.size() < amount) return;  // iterator exhausted: everything is chosen
    for (uint64_t i = 0; amount + i < max_index; i++) {
        const uint64_t ubound = i + 1 + amount;
        const uint64_t k = ubound <= 0xffffffffull ? rng.gen_range_u32((uint32_t)ubound) : rng.uniform(ubound);
        if (k < amount) out[k] = amount + i;
    }
}

void field_random(int fid, ChaCha20Rng &rng, uint64_t *out) {
    const FieldConsts fc = field_consts(fid);
    const int L = fc.limbs;
    const int top_bits = fc.num_bits - 64 * (L - 1);
    const uint64_t mask = top_bits >= 64 ? ~0ull : ((1ull << top_bits) - 1);
    for (;;) {
        for (int i = 0; i < L; i++) out[i] = rng.next_u64();
        out[L - 1] &= mask;
        bool lt = false;  // out < p ?
        for (int i = L - 1; i >= 0; i--) {
            if (out[i] != fc.p[i]) {
                lt = out[i] < fc.p[i];
                break;
            }
        }
        if (lt) return;
    }
}

// ------------------------------------------------------------------ Brakedown code generation

namespace {
struct Spec {
    uint64_t an, ad, bn, bd, rn, rd, blen;
};
// codespec.rs:168-232 (alpha, beta, r as fractions; base-case length)
const Spec SPECS[6] = {{239, 2000, 71, 2500, 71, 50, 20},   {69, 500, 111, 2500, 147, 100, 20},
                       {89, 500, 61, 1000, 1521, 1000, 20}, {1, 5, 41, 500, 41, 25, 20},
                       {211, 1000, 97, 1000, 202, 125, 20}, {119, 500, 241, 2000, 43, 25, 20}};

double entropy(double z) { return -z * std::log2(z) - (1.0 - z) * std::log2(1.0 - z); }
uint64_t cmd(uint64_t n, uint64_t num, uint64_t den) { return (n * num + den - 1) / den; }
}  // namespace

double sdig_dist(int code) {
    const Spec &s = SPECS[code - 1];
    return (double)(s.bn * s.rd) / (double)(s.bd * s.rn);
}

bool sdig_get_dims(int code, uint64_t n, double log2p, std::vector<SdigDims> &pre, std::vector<SdigDims> &post) {
    if (code < 1 || code > 6) return false;
    const Spec &s = SPECS[code - 1];
    if (n <= s.blen) return false;
    const double alpha = (double)s.an / (double)s.ad, beta = (double)s.bn / (double)s.bd, r = (double)s.rn / (double)s.rd;
    const double mu = r - 1.0 - r * alpha, nu = beta + alpha * beta + 0.03;
    const double cn1 = entropy(beta) + alpha * entropy(1.28 * beta / alpha);
    const double cn2 = beta * std::log2(alpha / (1.28 * beta));
    const double dn1 = r * alpha * entropy(beta / r) + mu * entropy(nu / mu);
    const double dn2 = alpha * beta * std::log2(mu / nu);
    std::vector<uint64_t> sizes;  // matgen.rs:66-74
    for (uint64_t ni = n; ni > s.blen; ni = cmd(ni, s.an, s.ad)) sizes.push_back(ni);
    sizes.push_back(cmd(sizes.back(), s.an, s.ad));
    pre.clear();
    post.clear();
    for (size_t i = 0; i + 1 < sizes.size(); i++) {
        const uint64_t ni = sizes[i], mi = sizes[i + 1];
        uint64_t cn = std::min(std::max(cmd(ni, 32 * s.bn, 25 * s.bd), 4 + cmd(ni, s.bn, s.bd)),
                               (uint64_t)std::ceil((110.0 / (double)ni + cn1) / cn2));
        cn = std::min(cn, mi);
        pre.push_back({ni, mi, cn});
        const uint64_t nip = cmd(mi, s.rn, s.rd);
        const uint64_t mip = cmd(ni, s.rn, s.rd) - ni - nip;
        const uint64_t t1 = cmd(ni, 2 * s.bn, s.bd);
        const uint64_t t2 = cmd(ni, s.rn, s.rd) - ni + 110;
        uint64_t dn = std::min(t1 + (uint64_t)std::ceil((double)t2 / log2p),
                               (uint64_t)std::ceil((110.0 / (double)ni + dn1) / dn2));
        dn = std::min(dn, mip);
        post.push_back({nip, mip, dn});
    }
    return true;
}

void sdig_gen_code(int fid, ChaCha20Rng &rng, const SdigDims &dim, uint64_t *indptr, uint64_t *indices, uint64_t *data) {
    const int L = field_consts(fid).limbs;
    std::vector<uint64_t> picked;
    picked.reserve(dim.d);
    uint64_t nnz = 0;
    indptr[0] = 0;
    for (uint64_t c = 0; c < dim.n; c++) {
        picked.clear();
        while (picked.size() < dim.d) {  // d distinct row indices by rejection (matgen.rs:145-158)
            const uint64_t x = rng.uniform(dim.m);
            if (std::find(picked.begin(), picked.end(), x) == picked.end()) picked.push_back(x);
        }
        std::sort(picked.begin(), picked.end());
        for (uint64_t row : picked) {  // one non-zero element per index (matgen.rs:166-181)
            uint64_t *v = data + nnz * L;
            bool zero;
            do {
                field_random(fid, rng, v);
                zero = true;
                for (int l = 0; l < L; l++) zero &= (v[l] == 0);
            } while (zero);
            indices[nnz++] = row;
        }
        indptr[c + 1] = nnz;
    }
}

}  // namespace host
}  // namespace lcpc
