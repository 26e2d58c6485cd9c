// Micro-benchmark: cost of chains of tcgen05.mma kind::tf32 (K-major, no swizzle) under different conditions.
#include <cstdio>
#include <cuda_runtime.h>
#include "../ibinn_b200/csrc/tc.cuh"

struct Cfg { int n; int nmma; int two_acc; int row_off; int advance; int lbo_rows; int m; };

__global__ void __launch_bounds__(128) bench(Cfg c, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 1.0f;
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_s, 128);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = tmem_s;
    if (tid == 0) {
        const uint32_t lbo = c.lbo_rows * 16;
        const uint32_t idesc = tc::idesc_tf32(c.m, c.n);
        const uint64_t da0 = tc::smem_desc(tc::smem_u32(sm) + c.row_off * 16, lbo);
        const uint64_t db0 = tc::smem_desc(tc::smem_u32(sm) + 100 * 1024, lbo);
        // warm
        tc::mma_tf32(tm, da0, db0, idesc, false);
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        long long t0 = clock64();
        uint64_t da = da0, db = db0;
        for (int i = 0; i < c.nmma; ++i) {
            tc::mma_tf32(tm + (c.two_acc ? (i & 1) * 64 : 0), da, db, idesc, true);
            if (c.advance) { da += (2 * lbo) >> 4; db += (2 * lbo) >> 4; if ((i & 7) == 7) { da = da0; db = db0; } }
        }
        long long t1 = clock64();
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 1);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 128);
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    Cfg cfgs[] = {
        {64, 64, 0, 0, 0, 128, 128}, {64, 256, 0, 0, 0, 128, 128}, {16, 256, 0, 0, 0, 128, 128}, {256, 256, 0, 0, 0, 128, 128},
        {64, 256, 1, 0, 0, 128, 128}, {64, 256, 0, 3, 0, 128, 128}, {64, 256, 0, 0, 1, 128, 128}, {64, 256, 0, 0, 1, 148, 128},
        {64, 256, 0, 11, 1, 148, 128}, {16, 256, 0, 11, 1, 196, 128}, {64, 256, 0, 0, 0, 128, 64}, {128, 256, 0, 0, 0, 128, 128},
        {32, 256, 0, 0, 1, 164, 128},
    };
    for (auto& c : cfgs) {
        long long h[2];
        bench<<<1, 128, 200 * 1024>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("M=%3d N=%3d nmma=%3d two_acc=%d row_off=%2d advance=%d lbo_rows=%3d : issue %6.1f cyc/mma, total %6.1f cyc/mma  (%s)\n", c.m, c.n, c.nmma, c.two_acc,
               c.row_off, c.advance, c.lbo_rows, (double)h[0] / c.nmma, (double)h[1] / c.nmma, cudaGetErrorString(e));
    }
    return 0;
}
