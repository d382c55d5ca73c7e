// C++ client of the boundary, written the way the reference's own tests read (lcpc-ligero-pc/src/tests.rs:196-243:
// commit, prove, verify, and check the verifier's value against a direct evaluation) -- through include/lcpc_b200.hpp.
//
//   g++ -std=c++17 -Iinclude examples/host_mirror.cpp -Llcpc_proof_of_storage_b200/_lib -llcpc_b200
//       -Wl,-rpath,$PWD/lcpc_proof_of_storage_b200/_lib -o host_mirror && ./host_mirror
//
// Without a CUDA device only the host-side pieces run (parameters, transcript); the library has no CPU fallback and
// Context construction reports that.
#include <cstdio>
#include <cstring>

#include "lcpc_b200.hpp"

using namespace lcpc_b200;

static int fail(const char *what) {
    std::fprintf(stderr, "FAILED: %s\n", what);
    return 1;
}

int main() {
    // --- host-side: parameters of LigeroEncoding::new(2^16) over the 63-bit field (BASELINE configs[0]) ---
    const FieldInfo f63(LCPC_FT63);
    if (f63.limbs != 1 || f63.num_bits != 63 || f63.two_adicity != 41) return fail("field constants");
    const auto dims = LigeroEncoding::get_dims_for_len(f63, size_t(1) << 16);
    if (dims[0] != 32 || dims[1] != 2048 || dims[2] != 4096) return fail("LigeroEncoding::_get_dims");
    if (LigeroEncoding::n_col_opens(1, 2) != 309) return fail("n_col_opens");
    if (n_degree_tests(128, 4096, f63.flog2()) != 3) return fail("n_degree_tests");
    if (lcpc_b200::log2(size_t(5)) != 3 || lcpc_b200::log2(size_t(4096)) != 12) return fail("log2");

    // --- Brakedown: SdigEncodingS::new(2^24, seed) over the 255-bit field: the shapes of SURVEY.md section 8 ---
    {
        const FieldInfo f255(LCPC_FT255);
        if (SdigEncoding::n_col_opens(3) != 6593) return fail("Sdig n_col_opens");
        const size_t npr = SdigEncoding::n_per_row_for_len(f255, size_t(1) << 24, 3);
        if (npr != 166292) return fail("Sdig n_per_row for 2^24 over Ft255");
        const auto mats = SdigEncoding::generate(LCPC_FT63, 150, /*seed=*/0, 3);  // small code, host-side generation
        if (mats.first.empty() || mats.first.size() != mats.second.size() || mats.first[0].cols != 150) return fail("matgen shapes");
        if (SdigEncoding::codeword_length(mats.first, mats.second) <= 150) return fail("codeword_length");
    }

    // --- merlin's published test vector (merlin 2.0, `equivalence_simple`) through the C ABI transcript ---
    {
        Transcript tr("test protocol");
        const char *msg = "some data";
        tr.append_message("some label", reinterpret_cast<const uint8_t *>(msg), std::strlen(msg));
        uint8_t ch[32];
        tr.challenge_bytes("challenge", ch, sizeof ch);
        static const uint8_t expect[32] = {0xd5, 0xa2, 0x19, 0x72, 0xd0, 0xd5, 0xfe, 0x32, 0x0c, 0x0d, 0x26, 0x3f, 0xac, 0x7f, 0xff, 0xb8,
                                           0x14, 0x5a, 0xa6, 0x40, 0xaf, 0x6e, 0x9b, 0xca, 0x17, 0x7c, 0x03, 0xc7, 0xef, 0xcf, 0x06, 0x15};
        if (std::memcmp(ch, expect, 32) != 0) return fail("merlin test vector");
    }

    // --- device: commit + prove + verify of a 2^12-coefficient polynomial, value checked by Horner on the host ---
    try {
        Context ctx(0);
        const size_t n = size_t(1) << 12;
        const uint64_t p = 0x46d0760000000001ull;
        std::vector<uint64_t> coeffs(n);
        uint8_t key[32] = {7};
        check(lcpc_random_field_vec(LCPC_FT63, key, coeffs.data(), n));  // F::random x n from ChaCha20 (reduced elements)
        auto enc = LigeroEncoding::create(ctx, LCPC_FT63, n);
        auto comm = LcCommit::commit(coeffs, enc);
        const Digest root = comm.get_root();
        // outer[i] = (x^n_per_row)^i, inner[j] = x^j, in Montgomery form: built with the library's own field arithmetic
        // by folding unit vectors is overkill here -- use x = 1 (every power is F::ONE), so the value is sum(coeffs)
        uint64_t one = 0;
        check(lcpc_field_constants(LCPC_FT63, nullptr, &one, nullptr, nullptr, nullptr));
        std::vector<uint64_t> outer(comm.n_rows, one), inner(comm.n_per_row, one);
        Transcript tp("host mirror");
        tp.append_message("polycommit", root);
        auto proof = comm.prove(outer, enc, tp);
        Transcript tv("host mirror");
        tv.append_message("polycommit", root);
        const auto value = proof.verify(root, outer, inner, enc, tv);
        // sum of the coefficients mod p, in Montgomery form (addition is representation-independent)
        unsigned __int128 acc = 0;
        for (uint64_t c : coeffs) acc = (acc + c) % p;
        if (value[0] != (uint64_t)acc) return fail("verifier's value != direct evaluation at x = 1");
        // a tampered column must be rejected with the reference's variant
        proof.columns[0].col[0] ^= 1;
        Transcript tw("host mirror");
        tw.append_message("polycommit", root);
        try {
            proof.verify(root, outer, inner, enc, tw);
            return fail("tampered proof accepted");
        } catch (const Error &e) {
            if (!e.is_verifier_error()) return fail("tamper: not a VerifierError");
            std::printf("tampered column rejected: %s\n", e.variant().c_str());
        }
        std::printf("device path ok: %zu x %zu -> %zu, %zu openings\n", comm.n_rows, comm.n_per_row, comm.n_cols, proof.columns.size());
    } catch (const Error &e) {
        if (e.code != LCPC_ERR_CUDA) {
            std::fprintf(stderr, "unexpected error: %s\n", e.what());
            return 1;
        }
        std::printf("no CUDA device: host-side checks only (%s)\n", e.what());
    }
    std::printf("host mirror ok\n");
    return 0;
}
