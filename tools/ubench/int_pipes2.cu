// Second integer-pipe microbenchmark: pure instruction streams on independent dependent-chains, and fixed
// mixes of two instruction kinds on disjoint chains, to read the issue rate of IMAD / IMAD.WIDE / ALU ops
// and whether they overlap.  Loop bodies are checked with cuobjdump (ptxas rewrites PTX freely).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 8192
__constant__ uint32_t C_Q = 0xb92f8a00u;

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// NW chains of IMAD.WIDE, NI chains of IMAD, NL chains of LOP3, NA chains of 64-bit add (IADD3 + IADD3.X)
template <int NW, int NI, int NL, int NA>
__global__ void __launch_bounds__(256) k(uint32_t *out, long long *tm, uint32_t seed) {
    uint64_t w[NW + 1], s[NA + 1];
    uint32_t a[NI + 1], l[NL + 1];
    const uint32_t q = C_Q;
#pragma unroll
    for (int i = 0; i <= NW; i++) w[i] = seed * 77u + threadIdx.x + i;
#pragma unroll
    for (int i = 0; i <= NI; i++) a[i] = seed * 13u + threadIdx.x * 3 + i;
#pragma unroll
    for (int i = 0; i <= NL; i++) l[i] = seed * 5u + threadIdx.x * 7 + i;
#pragma unroll
    for (int i = 0; i <= NA; i++) s[i] = ((uint64_t)seed << 32) + threadIdx.x * 11 + i;
    const uint64_t inc = ((uint64_t)seed << 33) | 0xfffffff1u;
    __syncthreads();
    long long g0 = gtime(), c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < NW; i++) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32)));
#pragma unroll
            for (int i = 0; i < NI; i++) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(q));
#pragma unroll
            for (int i = 0; i < NL; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[i]) : "r"(q), "r"(seed));
#pragma unroll
            for (int i = 0; i < NA; i++) {
                uint32_t lo = (uint32_t)s[i], hi = (uint32_t)(s[i] >> 32);
                asm volatile("{ add.cc.u32 %0, %0, %1; addc.u32 %1, %1, %0; }" : "+r"(lo), "+r"(hi));
                s[i] = ((uint64_t)hi << 32) | lo;
            }
        }
    }
    long long c1 = clock64(), g1 = gtime();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) acc ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
#pragma unroll
    for (int i = 0; i < NI; i++) acc ^= a[i];
#pragma unroll
    for (int i = 0; i < NL; i++) acc ^= l[i];
#pragma unroll
    for (int i = 0; i < NA; i++) acc ^= (uint32_t)s[i] ^ (uint32_t)(s[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) { tm[3 * blockIdx.x] = c1 - c0; tm[3 * blockIdx.x + 1] = g0; tm[3 * blockIdx.x + 2] = g1; }
}

template <int NW, int NI, int NL, int NA>
static void run(const char *name) {
    int nsm;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int bps = 4, blocks = nsm * bps;
    uint32_t *out; long long *tm;
    cudaMalloc(&out, (size_t)blocks * 256 * 4);
    cudaMalloc(&tm, 3 * blocks * sizeof(long long));
    for (int rep = 0; rep < 2; rep++) { k<NW, NI, NL, NA><<<blocks, 256>>>(out, tm, 12345 + rep); cudaDeviceSynchronize(); }
    long long *h = new long long[3 * blocks];
    cudaMemcpy(h, tm, 3 * blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    long long gmin = h[1], gmax = h[2]; double ghz = 0;
    for (int i = 0; i < blocks; i++) { if (h[3*i+1] < gmin) gmin = h[3*i+1]; if (h[3*i+2] > gmax) gmax = h[3*i+2]; ghz += (double)h[3*i] / (double)(h[3*i+2] - h[3*i+1]); }
    ghz /= blocks;
    const double span = (double)(gmax - gmin) * ghz;              // SM cycles for the whole grid
    const double per_warp_iter = span / ((double)ITERS * 4 * 8 * bps / 4);  // cycles per (one unrolled repetition of one warp) per SMSP
    printf("%-34s W=%d I=%d L=%d A=%d  %6.2f SMSP-cycles per repetition  (%.2f GHz)  %s\n", name, NW, NI, NL, NA, per_warp_iter, ghz,
           cudaGetErrorString(cudaGetLastError()));
    delete[] h; cudaFree(out); cudaFree(tm);
}

int main() {
    run<8, 0, 0, 0>("IMAD.WIDE x8");
    run<0, 8, 0, 0>("IMAD x8");
    run<0, 0, 8, 0>("LOP3 x8");
    run<0, 0, 0, 8>("add64 x8");
    run<4, 0, 4, 0>("IMAD.WIDE x4 + LOP3 x4");
    run<8, 0, 8, 0>("IMAD.WIDE x8 + LOP3 x8");
    run<0, 8, 8, 0>("IMAD x8 + LOP3 x8");
    run<4, 4, 0, 0>("IMAD.WIDE x4 + IMAD x4");
    run<8, 8, 0, 0>("IMAD.WIDE x8 + IMAD x8");
    run<4, 0, 8, 0>("IMAD.WIDE x4 + LOP3 x8");
    run<8, 0, 4, 0>("IMAD.WIDE x8 + LOP3 x4");
    run<6, 0, 0, 6>("IMAD.WIDE x6 + add64 x6");
    run<6, 2, 8, 0>("IMAD.WIDE x6 + IMAD x2 + LOP3 x8");
    return 0;
}
