// Integer-pipe issue rates of the device, measured in place: the denominators of the second roofline bench.py reports
// (the commit kernels are bound by the 32-bit integer pipes, not by HBM -- DESIGN.md section 3).  Pure streams of one
// instruction kind on eight independent dependency chains per thread, four resident 256-thread CTAs per SM, timed with
// clock64 / %globaltimer inside the kernel; the loop bodies were checked with cuobjdump (tools/ubench/int_pipes2.cu is
// the stand-alone form this was taken from, profiles/r01c_int_pipes.txt its output).
#include "lcpc_handles.h"

using namespace lcpc;
using namespace lcpc::abi;

namespace {

constexpr int UB_ITERS = 2048;
__constant__ uint32_t UB_Q = 0xb92f8a00u;

__device__ __forceinline__ long long ub_gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// NW chains of IMAD.WIDE, NI chains of IMAD, NL chains of LOP3
template <int NW, int NI, int NL>
__global__ void __launch_bounds__(256) k_int_pipes(uint32_t *out, long long *tm, uint32_t seed) {
    uint64_t w[NW + 1];
    uint32_t a[NI + 1], l[NL + 1];
    const uint32_t q = UB_Q;
#pragma unroll
    for (int i = 0; i <= NW; i++) w[i] = seed * 77u + threadIdx.x + i;
#pragma unroll
    for (int i = 0; i <= NI; i++) a[i] = seed * 13u + threadIdx.x * 3 + i;
#pragma unroll
    for (int i = 0; i <= NL; i++) l[i] = seed * 5u + threadIdx.x * 7 + i;
    __syncthreads();
    const long long g0 = ub_gtime(), c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < UB_ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < NW; i++)
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32)));
#pragma unroll
            for (int i = 0; i < NI; i++) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(q));
#pragma unroll
            for (int i = 0; i < NL; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i]) : "r"(q), "r"(seed));
        }
    }
    const long long c1 = clock64(), g1 = ub_gtime();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) acc ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
#pragma unroll
    for (int i = 0; i < NI; i++) acc ^= a[i];
#pragma unroll
    for (int i = 0; i < NL; i++) acc ^= l[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) {
        tm[3 * blockIdx.x] = c1 - c0;
        tm[3 * blockIdx.x + 1] = g0;
        tm[3 * blockIdx.x + 2] = g1;
    }
}

// SM sub-partition cycles per warp instruction of the stream (all resident warps issuing), and the SM clock in GHz
template <int NW, int NI, int NL>
cudaError_t run_stream(lcpc_ctx *ctx, int n_sm, uint32_t *d_out, long long *d_tm, std::vector<long long> &h, double *cyc, double *ghz) {
    const int bps = 4, blocks = n_sm * bps;
    for (int rep = 0; rep < 2; rep++) k_int_pipes<NW, NI, NL><<<blocks, 256, 0, ctx->stream>>>(d_out, d_tm, 12345u + rep);
    ctx->launches += 2;
    cudaError_t e = cudaMemcpyAsync(h.data(), d_tm, 3 * (size_t)blocks * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return e;
    long long gmin = h[1], gmax = h[2];
    double g = 0;
    for (int i = 0; i < blocks; i++) {
        if (h[3 * i + 1] < gmin) gmin = h[3 * i + 1];
        if (h[3 * i + 2] > gmax) gmax = h[3 * i + 2];
        g += (double)h[3 * i] / (double)(h[3 * i + 2] - h[3 * i + 1]);
    }
    g /= blocks;
    const double span_cycles = (double)(gmax - gmin) * g;
    // per SM sub-partition: bps CTAs x 8 warps / 4 sub-partitions, each issuing UB_ITERS * 4 * (NW + NI + NL) instructions
    const double instr_per_smsp = (double)UB_ITERS * 4 * (NW + NI + NL) * (8.0 * bps / 4.0);
    *cyc = span_cycles / instr_per_smsp;
    *ghz = g;
    return cudaSuccess;
}

}  // namespace

extern "C" int32_t lcpc_ctx_measure_int_pipes(lcpc_ctx *ctx, double cycles_out[4], double *sm_ghz_out) {
    if (!ctx || !cycles_out) return fail(LCPC_ERR_INVALID_ARG, "null argument");
    ctx = primary(ctx);
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    int n_sm = 0;
    CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device));
    const int blocks = n_sm * 4;
    DevBuf d_out, d_tm;
    CU(d_out.alloc((size_t)blocks * 256 * 4, ctx->stream));
    CU(d_tm.alloc(3 * (size_t)blocks * sizeof(long long), ctx->stream));
    std::vector<long long> h(3 * (size_t)blocks);
    double ghz = 0;
    CU((run_stream<8, 0, 0>(ctx, n_sm, d_out.as<uint32_t>(), d_tm.as<long long>(), h, &cycles_out[0], &ghz)));  // IMAD.WIDE
    CU((run_stream<0, 8, 0>(ctx, n_sm, d_out.as<uint32_t>(), d_tm.as<long long>(), h, &cycles_out[1], &ghz)));  // IMAD
    CU((run_stream<0, 0, 8>(ctx, n_sm, d_out.as<uint32_t>(), d_tm.as<long long>(), h, &cycles_out[2], &ghz)));  // LOP3 (ALU pipe)
    CU((run_stream<0, 8, 8>(ctx, n_sm, d_out.as<uint32_t>(), d_tm.as<long long>(), h, &cycles_out[3], &ghz)));  // IMAD + LOP3 mix
    if (sm_ghz_out) *sm_ghz_out = ghz;
    return LCPC_OK;
}
