// Integer-pipe throughput microbenchmark for sm_100a: which SASS forms the field arithmetic should be
// built from.  Each test is an unrolled loop of 8 independent chains of one instruction pattern;
// reports warp-instructions per clock per SM (from clock64 deltas) so the numbers are clock-independent.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o int_pipes int_pipes.cu && ./int_pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 16384
#define CHAINS 8

__constant__ uint32_t C_ONE = 1u, C_Q = 0xb92f8a00u;

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <int T>
__global__ void __launch_bounds__(256) k(uint32_t *out, long long *cyc, uint32_t seed) {
    uint32_t a[CHAINS], b[CHAINS];
    uint64_t w[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i * 11 + threadIdx.x; w[i] = ((uint64_t)a[i] << 32) | b[i]; }
    const uint32_t one = C_ONE, q = C_Q;
    __syncthreads();
    long long g0 = gtime();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (T == 0) {  // IADD3
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            } else if (T == 1) {  // IMAD (32-bit)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(q), "r"(b[i]));
            } else if (T == 2) {  // IMAD.WIDE with 64-bit accumulate
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(q));
            } else if (T == 3) {  // alternating IADD3 / IMAD
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(b[i]) : "r"(q));
            } else if (T == 4) {  // alternating IADD3 / IMAD.WIDE
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(q));
            } else if (T == 5) {  // 64-bit add via carry chain (IADD3 + IADD3.X)
                asm volatile("{ add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3; }" : "+r"(a[i]), "+r"(b[i]) : "r"(q), "r"(one));
            } else if (T == 6) {  // ISETP + 2 SEL
                asm volatile("{ .reg .pred p; setp.lt.s32 p, %0, 0; selp.u32 %0, %1, %0, p; selp.u32 %1, %0, %1, p; }" : "+r"(a[i]), "+r"(b[i]));
            } else if (T == 7) {  // ISETP + 2 predicated adds
                asm volatile("{ .reg .pred p; setp.lt.s32 p, %1, 0; @p add.cc.u32 %0, %0, 1; @p addc.u32 %1, %1, 0x46d07600; }" : "+r"(a[i]), "+r"(b[i]));
            } else if (T == 8) {  // LOP3
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(q));
            } else if (T == 9) {  // SHF (funnel shift = rotate)
                asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(a[i]));
            } else if (T == 10) {  // IMAD.WIDE without accumulate (mul.wide)
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(q));
                a[i] ^= (uint32_t)w[i];
            } else if (T == 11) {  // IMAD.HI
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(q), "r"(b[i]));
            } else if (T == 12) {  // 3-input add with two carries: a + b + c over 64 bits
                asm volatile("{ .reg .u32 t0, t1; add.cc.u32 t0, %0, %2; addc.u32 t1, %1, %3; add.cc.u32 %0, t0, 0xffffffff; addc.u32 %1, t1, 0xb92f89ff; }"
                             : "+r"(a[i]), "+r"(b[i]) : "r"(q), "r"(one));
            } else if (T == 13) {  // PRMT
                asm volatile("prmt.b32 %0, %0, %1, 0x1230;" : "+r"(a[i]) : "r"(b[i]));
            } else if (T == 14) {  // 32-bit add issued as IMAD (x*1+y with opaque 1)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(one), "r"(b[i]));
            } else if (T == 16) {  // SEL alone
                asm volatile("{ .reg .pred p; setp.lt.s32 p, %1, 0; selp.u32 %0, %1, %0, p; }" : "+r"(a[i]) : "r"(b[i]));
            } else if (T == 17) {  // 3 ALU : 1 IMAD
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(q));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(a[i]), "r"(q));
            } else if (T == 18) {  // 1 ALU : 1 IMAD.WIDE(RZ)
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(q));
                a[i] += (uint32_t)w[i];
            } else if (T == 19) {  // 3 ALU : 1 IMAD.WIDE(RZ)
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(q));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(q));
                a[i] += (uint32_t)w[i];
            } else if (T == 20) {  // 2 IMAD : 1 ALU
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(q), "r"(b[i]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(q), "r"(a[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(q));
            } else if (T == 15) {  // 2 IADD3 : 1 IMAD.WIDE
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i]));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(q));
            }
        }
    }
    long long t1 = clock64();
    long long g1 = gtime();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) { cyc[3 * blockIdx.x] = t1 - t0; cyc[3 * blockIdx.x + 1] = g0; cyc[3 * blockIdx.x + 2] = g1; }
}

template <int T>
static void run(const char *name, int instr_per_chain_iter, int blocks_per_sm) {
    int nsm;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = nsm * blocks_per_sm;
    uint32_t *out;
    long long *cyc;
    cudaMalloc(&out, (size_t)blocks * 256 * 4);
    cudaMalloc(&cyc, 3 * blocks * sizeof(long long));
    k<T><<<blocks, 256>>>(out, cyc, 12345);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<T><<<blocks, 256>>>(out, cyc, 999);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long *h = new long long[3 * blocks];
    cudaMemcpy(h, cyc, 3 * blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    long long gmin = h[1], gmax = h[2];
    double ghz = 0;
    for (int i = 0; i < blocks; i++) {
        avg += (double)h[3 * i];
        if (h[3 * i + 1] < gmin) gmin = h[3 * i + 1];
        if (h[3 * i + 2] > gmax) gmax = h[3 * i + 2];
        ghz += (double)h[3 * i] / (double)(h[3 * i + 2] - h[3 * i + 1]);
    }
    avg /= blocks;
    ghz /= blocks;
    const double span_cyc = (double)(gmax - gmin) * ghz;  // whole-kernel span in SM cycles
    // warp-instructions issued per SM during one block's lifetime: blocks_per_sm blocks x 8 warps
    const double winstr = (double)ITERS * CHAINS * instr_per_chain_iter * 8 * blocks_per_sm;
    printf("%-40s %d blk/SM  per-block %.3f  whole-span %.3f ptx-instr/clk/SMSP  sm clock %.3f GHz  %.3f ms  %s\n", name, blocks_per_sm,
           winstr / avg / 4, winstr / span_cyc / 4, ghz, ms, cudaGetErrorString(cudaGetLastError()));
    delete[] h;
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int bps = 4; bps <= 4; bps += 2) {
        run<0>("IADD3", 1, bps);
        run<1>("IMAD 32", 1, bps);
        run<14>("IMAD 32 (x*1+y)", 1, bps);
        run<2>("IMAD.WIDE acc64", 1, bps);
        run<10>("IMAD.WIDE (mul.wide)+LOP", 2, bps);
        run<11>("IMAD.HI", 1, bps);
        run<3>("IADD3 + IMAD", 2, bps);
        run<4>("IADD3 + IMAD.WIDE", 2, bps);
        run<15>("2 IADD3 + IMAD.WIDE", 3, bps);
        run<5>("IADD3 + IADD3.X (64-bit add)", 2, bps);
        run<12>("64-bit a+b+c (3-input or 2 chains)", 4, bps);
        run<6>("ISETP + 2 SEL", 3, bps);
        run<7>("ISETP + 2 predicated IADD3", 3, bps);
        run<16>("ISETP + SEL", 2, bps);
        run<17>("3 ALU + 1 IMAD", 4, bps);
        run<18>("IMAD.WIDE(RZ) + IADD3", 2, bps);
        run<19>("IMAD.WIDE(RZ) + 3 ALU", 4, bps);
        run<20>("2 IMAD + 1 LOP3", 3, bps);
        run<8>("LOP3", 1, bps);
        run<9>("SHF", 1, bps);
        run<13>("PRMT", 1, bps);
    }
    return 0;
}
