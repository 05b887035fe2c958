// Micro-benchmark: issue rate of IMAD vs dp2a / dp4a (IDP) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(int *out, int n, unsigned a0, unsigned b0) {
    unsigned a = a0 + threadIdx.x, b = b0;
    int acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = i;
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) acc[i] = acc[i] + (int)(a + i) * (int)b;
            else if (MODE == 1) asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a + i), "r"(b));
            else if (MODE == 2) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a + i), "r"(b));
            else asm volatile("dp2a.hi.s32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a + i), "r"(b));
        }
        b += 0x01010101u;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(int *d, int n) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, n, 12345, 0x01020304);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, n, 12345, 0x01020304);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    int *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int n = 20000;
    const double ops = 148.0 * 8 * 256 * 8 * n;
    float t0 = run<0>(d, n), t1 = run<1>(d, n), t2 = run<2>(d, n), t3 = run<3>(d, n);
    printf("imad %.3f ms %.1f Gop/s | dp2a.lo.u.s %.3f ms %.1f | dp4a %.3f ms %.1f | dp2a.hi.s.s %.3f ms %.1f\n",
           t0, ops / t0 / 1e6, t1, ops / t1 / 1e6, t2, ops / t2 / 1e6, t3, ops / t3 / 1e6);
    return 0;
}
