// Dependent-issue latency of the integer ops on the range coder's chain (one thread, one warp, one SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 4096
template <int OP> __global__ void k(uint32_t *out, uint32_t seed, uint32_t c, long long *cyc) {
    uint32_t x = seed, y = seed * 3 + 1;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = x + y;                                   // IADD
        if (OP == 1) x = __clz(x) + y;                            // FLO + IADD
        if (OP == 2) x = (uint32_t)(((uint64_t)x * c + c) >> 16); // IMAD.WIDE + SHF.R.U64
        if (OP == 3) x = (x << (y & 31)) ^ y;                     // SHF + LOP
        if (OP == 4) x = (x >= y) ? x - y : y - 1;                // ISETP + SEL-ish
        if (OP == 5) x = __funnelshift_r(x, y, x & 31);           // SHF dependent on its own shift amount
        if (OP == 6) x = (x & y) | (~x & c);                      // LOP3
        if (OP == 7) x = __clz(x | 1) ;                           // FLO only (result feeds next FLO)
        if (OP == 8) { uint64_t p = (uint64_t)x * c + c; x = (uint32_t)(p >> 32) + (uint32_t)p; }  // IMAD.WIDE + IADD
        if (OP == 9) x = __popc(x) + y;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[OP] = t1 - t0;
}
int main() {
    uint32_t *out; long long *cyc, h[16];
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 16 * 8);
    const char *names[] = {"IADD", "FLO+IADD", "IMAD.WIDE+SHF.R.U64", "SHF+LOP", "ISETP+sel", "SHF(self amt)", "LOP3", "FLO", "IMAD.WIDE+IADD", "POPC+IADD"};
    for (int rep = 0; rep < 2; rep++) {
        k<0><<<1, 32>>>(out, 12345, 40000, cyc); k<1><<<1, 32>>>(out, 12345, 40000, cyc); k<2><<<1, 32>>>(out, 0xF2345678, 65000, cyc);
        k<3><<<1, 32>>>(out, 12345, 40000, cyc); k<4><<<1, 32>>>(out, 12345, 40000, cyc); k<5><<<1, 32>>>(out, 12345, 40000, cyc);
        k<6><<<1, 32>>>(out, 12345, 40000, cyc); k<7><<<1, 32>>>(out, 12345, 40000, cyc); k<8><<<1, 32>>>(out, 12345, 40000, cyc); k<9><<<1, 32>>>(out, 12345, 40000, cyc);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 10; i++) printf("%-22s %.2f cycles/iter\n", names[i], (double)h[i] / N);
    return 0;
}
