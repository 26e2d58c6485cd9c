// Micro-benchmark: DFMA vs FFMA throughput of one SM, and latency of a batch of ld.global.cg loads.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* o, float* of, long long* t, int n) {
    double a = threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    long long t1 = clock64();
    float x = threadIdx.x * 1e-3f, y = 1.0000001f, z = 1e-9f;
    for (int i = 0; i < n; ++i) { x = fmaf(x, y, z); x = fmaf(x, y, z); x = fmaf(x, y, z); x = fmaf(x, y, z); }
    long long t2 = clock64();
    o[threadIdx.x] = a; of[threadIdx.x] = x;
    if (threadIdx.x == 0) { t[0] = t1 - t0; t[1] = t2 - t1; }
}
int main() {
    double* o; float* of; long long* t; cudaMalloc(&o, 8192); cudaMalloc(&of, 4096); cudaMalloc(&t, 16);
    for (int nt : {32, 128, 320, 1024}) {
        k<<<1, nt>>>(o, of, t, 256);
        cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, t, 16, cudaMemcpyDeviceToHost);
        printf("threads %4d: DFMA %.2f cyc per warp-instr-slot (dependent chain of 1024), FFMA %.2f\n", nt, h[0] / 1024.0, h[1] / 1024.0);
    }
    return 0;
}
