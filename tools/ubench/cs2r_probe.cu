// Stand-alone reproduction of profiles/r01g_cs2r_hazard.md: the first k_pack_bytes31 (31 file bytes -> four big-endian
// limbs, bytes beyond the end of the file read as zero through a PREDICATED load).  On sm_100a / CUDA 12.9 ptxas zeroes the
// 64-bit default of a term with `CS2R Rn, SRZ`, follows it with the predicated-off writer `@!P IMAD.WIDE.U32 Rn, ...`, and
// schedules the first reader 5 issue cycles after the CS2R.  Built in several variants (tools/ubench/cs2r_probe.sh) and run
// over byte counts 1..64; prints, per variant, the byte counts whose limbs differ from the host computation.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

__global__ void k_pack31(const uint8_t *__restrict__ bytes, size_t n_bytes, uint64_t *__restrict__ elems, size_t n_elems) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint64_t v = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const size_t b = e * 31 + 8 * i + k;
#ifdef PROBE_UNCONDITIONAL  // the shipped form: clamped unconditional load, value masked afterwards
            const uint64_t raw = bytes[b < n_bytes ? b : n_bytes - 1];
            const uint64_t byte = (8 * i + k < 31 && b < n_bytes) ? raw : 0;
#else
            const uint64_t byte = (8 * i + k < 31 && b < n_bytes) ? bytes[b] : 0;
#endif
            v = (v << 8) | byte;
        }
        elems[e * 4 + i] = v;
    }
}

int main() {
    uint8_t h[256];
    for (int i = 0; i < 256; i++) h[i] = (uint8_t)(0x80 | (i * 37 + 11));
    uint8_t *d;
    uint64_t *o;
    cudaMalloc(&d, 256);
    cudaMalloc(&o, 16 * 32);
    cudaMemcpy(d, h, 256, cudaMemcpyHostToDevice);
    int n_bad = 0;
    for (size_t n = 1; n <= 64; n++) {
        const size_t ne = (n + 30) / 31;
        cudaMemset(o, 0xEE, 16 * 32);
        k_pack31<<<1, 32>>>(d, n, o, ne);
        uint64_t got[12];
        cudaMemcpy(got, o, ne * 32, cudaMemcpyDeviceToHost);
        for (size_t e = 0; e < ne; e++)
            for (int i = 0; i < 4; i++) {
                uint64_t v = 0;
                for (int k = 0; k < 8; k++) {
                    const size_t b = e * 31 + 8 * i + k;
                    v = (v << 8) | ((8 * i + k < 31 && b < n) ? h[b] : 0);
                }
                if (v != got[e * 4 + i]) {
                    printf("  n_bytes %zu elem %zu limb %d: got %016llx expected %016llx\n", n, e, i, (unsigned long long)got[e * 4 + i],
                           (unsigned long long)v);
                    n_bad++;
                }
            }
    }
    printf("%s: %d wrong limbs (%s)\n", PROBE_NAME, n_bad, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
