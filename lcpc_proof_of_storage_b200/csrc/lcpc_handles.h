// Internal handle definitions and small helpers shared by the C-ABI translation units
// (lcpc_api.cu, lcpc_scheme.cu).  Not part of the public interface.
#pragma once
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/lcpc_b200.h"
#include "lcpc_field.cuh"
#include "lcpc_kernels.h"

namespace lcpc {
// CUDA-event stopwatch around every kernel launch of a context (off by default).
struct KernelTimer {
    struct Rec {
        const char *name;
        cudaEvent_t a, b;
    };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        cudaEvent_t e;
        if (!pool.empty()) {
            e = pool.back();
            pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        return e;
    }
    ~KernelTimer() {
        for (auto &r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        for (auto e : pool) cudaEventDestroy(e);
    }
};
}  // namespace lcpc

struct lcpc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t launches = 0;
    lcpc::KernelTimer *timer = nullptr;
    std::string timing_report;
    std::mutex mu;
    std::atomic<int> refs{1};  // the creator + every live plan
    // copy engines for the pipelined host commit: H2D and D2H run on their own streams so that
    // both PCIe directions and the kernels overlap
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> events;
    // multi-device context (lcpc_ctx_create_multi): one sub-context per listed device, each with its own stream; this
    // object then only carries the lock and the device of the first entry.  Empty for an ordinary context.
    std::vector<lcpc_ctx *> subs;
    // zeroed ticket counters of the one-launch hash + tree kernels (they leave them zeroed); grown on demand
    unsigned *d_tickets = nullptr;
    size_t n_tickets = 0;
    cudaError_t tickets(size_t n, unsigned **out) {
        if (n > n_tickets) {
            const size_t cap = n < 4096 ? 4096 : n;
            unsigned *p = nullptr;
            cudaError_t e = cudaMallocAsync((void **)&p, cap * sizeof(unsigned), stream);
            if (e != cudaSuccess) return e;
            if ((e = cudaMemsetAsync(p, 0, cap * sizeof(unsigned), stream)) != cudaSuccess) return e;
            if (d_tickets) cudaFreeAsync(d_tickets, stream);
            d_tickets = p;
            n_tickets = cap;
        }
        *out = d_tickets;
        return cudaSuccess;
    }
    lcpc::Launch lc() { return lcpc::Launch{stream, &launches, timer}; }
    cudaEvent_t event(size_t i) {
        while (events.size() <= i) {
            cudaEvent_t e;
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            events.push_back(e);
        }
        return events[i];
    }
};

struct lcpc_plan {
    lcpc_ctx *ctx = nullptr;
    int kind = 0;  // 0 = Ligero (NTT), 1 = Brakedown (SpMV chain)
    int fid = 0;
    size_t n_per_row = 0, n_cols = 0;
    lcpc::NttPlan ntt;
    lcpc::SdigPlan sdig;
    std::mutex mu;
    std::atomic<int> refs{1};  // the creator + every live commit
    // plan on a multi-device context: the same encoding built on every device (twiddles / CSR matrices replicated)
    std::vector<lcpc_plan *> subs;
};

// One device's part of a commitment made on a multi-device context: a block of whole rows of the
// coefficient and encoded matrices, and one block of columns of the Merkle tree's leaves.
struct lcpc_shard {
    size_t row0 = 0, rows = 0;      // my rows [row0, row0 + rows), cut on BLAKE3 chunk boundaries of the column leaves
    uint64_t c0 = 0, c1 = 0;        // the chunks of every leaf those rows make up
    size_t cols_local = 0;          // real columns inside my block of the padded leaf range
    uint64_t *d_coeffs = nullptr;   // [rows][n_per_row]
    uint64_t *d_comm = nullptr;     // [rows][n_cols]
    uint8_t *d_cvs = nullptr;       // chaining values of MY column block from EVERY device's rows: [n_chunks][cb][32 B]
    uint8_t *d_subtree = nullptr;   // flat Merkle tree over my cb leaves
    cudaEvent_t ev = nullptr, ev_tree = nullptr;
};

struct lcpc_commit {
    lcpc_plan *plan = nullptr;
    size_t n_rows = 0, n_per_row = 0, n_cols = 0, np2 = 0;
    uint64_t *d_coeffs = nullptr;
    uint64_t *d_comm = nullptr;
    uint8_t *d_hashes = nullptr;
    // BLAKE3 chunk chaining values of every column, [chunk][column][32 B] (null when a leaf is a single chunk):
    // kept so that a row edit re-hashes only the chunks it touches (lcpc_commit_update_rows_host)
    uint8_t *d_cvs = nullptr;
    std::mutex mu;
    // multi-device commitment: shards[g] lives on plan->ctx->subs[g]; d_top = the log2(n) levels above the subtree
    // roots, on the first device.  The single-device members above stay null.
    std::vector<lcpc_shard> shards;
    uint8_t *d_top = nullptr;
    size_t cb = 0;            // leaves per column block = np2 / devices
    uint64_t n_chunks = 0;    // BLAKE3 chunks per leaf
};


namespace lcpc {
namespace abi {

std::string &last_error();

inline int32_t fail(int32_t code, const std::string &msg) {
    last_error() = msg;
    return code;
}

inline int32_t cuda_fail(cudaError_t e, const char *what) {
    cudaGetLastError();  // clear the non-sticky error state
    return fail(e == cudaErrorMemoryAllocation ? LCPC_ERR_NOMEM : LCPC_ERR_CUDA,
                std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU(call)                                                          \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) return ::lcpc::abi::cuda_fail(e__, #call); \
    } while (0)

inline bool valid_field(int32_t f) { return f >= 0 && f < N_FIELDS; }
inline int limbs_of(int fid) { return field_consts(fid).limbs; }

inline size_t next_pow2(size_t v) {
    size_t p = 1;
    while (p < v) {
        if (p > (SIZE_MAX >> 1)) return 0;
        p <<= 1;
    }
    return p;
}

// stream-ordered device buffer that frees itself
struct DevBuf {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        s = stream;
        if (bytes == 0) bytes = 8;
        return cudaMallocAsync(&p, bytes, stream);
    }
    ~DevBuf() {
        if (p) cudaFreeAsync(p, s);
    }
    template <class T>
    T *as() { return reinterpret_cast<T *>(p); }
};

// merkleize (lib.rs:720-734) of a device matrix: leaves of columns [0, n_cols) into d_hashes, padding leaves zero, tree
// above them.  d_cvs_keep: where the chunk chaining values go when the caller keeps them (a commit handle)
int32_t merkleize_dev(lcpc_ctx *ctx, int fid, const uint64_t *d_comm, size_t n_rows, size_t row_stride, size_t n_cols,
                      size_t np2, uint8_t *d_hashes, uint8_t **d_cvs_keep = nullptr);

// multi-device paths (lcpc_multi.cu); all called with the parent plan / context / commit locks held
namespace multi {
bool usable(const lcpc_plan *plan, size_t n_rows);  // a chunk-aligned row partition with rows on every device exists
int32_t commit_host(lcpc_plan *plan, lcpc_commit *c, const uint64_t *coeffs, size_t n_coeffs, const uint8_t *file_bytes,
                    size_t n_bytes, uint64_t *coeffs_out, uint64_t *comm_out, uint8_t *hashes_out);
int32_t root(lcpc_commit *c, uint8_t *root_out);
int32_t download(lcpc_commit *c, uint64_t *coeffs_out, uint64_t *comm_out, uint8_t *hashes_out);
int32_t fold_host(lcpc_commit *c, int32_t which, const uint64_t *tensors, size_t n_tensors, uint64_t *out);
int32_t open_columns_host(lcpc_commit *c, const uint64_t *cols, size_t n, uint64_t *cols_out, uint8_t *paths_out);
int32_t leaves_host(lcpc_commit *c, const uint64_t *cols, size_t n, uint8_t *leaves_out);
void release(lcpc_commit *c);
}  // namespace multi

// the device-0 plan of a plan made on a multi-device context (the plan itself otherwise): entry points that work on one
// device (single rows, verification, streaming) run there
inline lcpc_plan *primary(lcpc_plan *p) { return (p && !p->subs.empty()) ? p->subs[0] : p; }
inline lcpc_ctx *primary(lcpc_ctx *c) { return (c && !c->subs.empty()) ? c->subs[0] : c; }

// encode rows already on the device (Ligero reads d_coeffs with stride n_per_row; Brakedown widens first)
int32_t encode_dev(lcpc_plan *plan, const uint64_t *d_coeffs, size_t n_rows, uint64_t *d_comm);

}  // namespace abi
}  // namespace lcpc
