// Cost of the multi-limb Montgomery building blocks on sm_100a: one row of N = 8 carry-chained 32x32 products
// (lcpc_mont32.cuh row_mad: mad.lo.cc / madc.hi.cc pairs -> IMAD.WIDE.U32.X) against the same eight products as
// independent IMAD.WIDE accumulations, and a whole 8-word Montgomery product.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../lcpc_proof_of_storage_b200/csrc -o widex widex.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "lcpc_field.cuh"

using namespace lcpc;
#define ITERS 4096

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <int T>
__global__ void __launch_bounds__(256) k(uint32_t *out, long long *tm, uint32_t seed) {
    using F = Field<FT255>;
    uint32_t a[8], b[8], X[18], Y[18];
    for (int i = 0; i < 8; i++) { a[i] = seed * 77u + threadIdx.x * 3 + i; b[i] = seed * 13u + threadIdx.x * 7 + i * 5; }
    for (int i = 0; i < 18; i++) X[i] = Y[i] = 0;
    uint64_t acc[8];
    for (int i = 0; i < 8; i++) acc[i] = i;
    typename F::E ea, eb;
    for (int i = 0; i < 4; i++) { ea.v[i] = ((uint64_t)a[2 * i + 1] << 32 | a[2 * i]) >> 2; eb.v[i] = ((uint64_t)b[2 * i + 1] << 32 | b[2 * i]) >> 2; }
    __syncthreads();
    long long g0 = gtime(), c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (T == 0) {  // one carry-chained row: 8 products
            m32::row_mad<8, 0>(X, Y, a[it & 7], [&](int j) { return b[j]; });
        } else if (T == 1) {  // 8 independent IMAD.WIDE accumulations
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] += (uint64_t)a[it & 7] * b[j];
        } else {  // a whole Montgomery product (dependent chain of them)
            ea = F::mul(ea, eb);
        }
    }
    long long c1 = clock64(), g1 = gtime();
    uint32_t r = 0;
    for (int i = 0; i < 18; i++) r ^= X[i] ^ Y[i];
    for (int i = 0; i < 8; i++) r ^= (uint32_t)acc[i] ^ (uint32_t)(acc[i] >> 32);
    for (int i = 0; i < 4; i++) r ^= (uint32_t)ea.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) { tm[3 * blockIdx.x] = c1 - c0; tm[3 * blockIdx.x + 1] = g0; tm[3 * blockIdx.x + 2] = g1; }
}

template <int T>
static void run(const char *name, double units) {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int bps = 2, blocks = nsm * bps;
    uint32_t *out; long long *tm;
    cudaMalloc(&out, (size_t)blocks * 256 * 4); cudaMalloc(&tm, 3 * blocks * sizeof(long long));
    for (int rep = 0; rep < 2; rep++) { k<T><<<blocks, 256>>>(out, tm, 12345 + rep); cudaDeviceSynchronize(); }
    long long *h = new long long[3 * blocks];
    cudaMemcpy(h, tm, 3 * blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    long long gmin = h[1], gmax = h[2]; double ghz = 0;
    for (int i = 0; i < blocks; i++) { if (h[3*i+1] < gmin) gmin = h[3*i+1]; if (h[3*i+2] > gmax) gmax = h[3*i+2]; ghz += (double)h[3*i] / (double)(h[3*i+2] - h[3*i+1]); }
    ghz /= blocks;
    const double span = (double)(gmax - gmin) * ghz;
    // per SM sub-partition: bps blocks * 8 warps / 4 = warps per SMSP
    const double cyc = span / ((double)ITERS * 8 * bps / 4);
    printf("%-46s %7.1f SMSP-cycles per iteration per warp = %.2f per 32x32 product  (%s)\n", name, cyc, cyc / units, cudaGetErrorString(cudaGetLastError()));
    delete[] h; cudaFree(out); cudaFree(tm);
}

int main() {
    run<0>("carry-chained row (8 x IMAD.WIDE.U32.X)", 8);
    run<1>("8 independent IMAD.WIDE accumulations", 8);
    run<2>("8-word Montgomery product (128 products)", 128);
    return 0;
}
