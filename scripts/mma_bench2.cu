// Micro-benchmark 2: tcgen05.mma kind::tf32 chains while OTHER warps of the CTA keep a second resource busy
// (shared-memory stores / loads, TMEM loads, global loads): which neighbour slows the tensor pipe down?
#include <cstdio>
#include <cuda_runtime.h>
#include "../ibinn_b200/csrc/tc.cuh"

struct Cfg { int n; int nmma; int bg; int bgwarps; int alt; };

__global__ void __launch_bounds__(288) bench(Cfg c, long long* out, const float* gsrc) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 200 * 1024 / 4; i += 288) reinterpret_cast<float*>(sm)[i] = 1.0f;
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); stop = 0; }
    if (warp == 0) tc::tmem_alloc(&tmem_s, 256);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = tmem_s;
    if (warp == 8) {
        const uint32_t lbo = 164 * 16;
        const uint32_t idesc = tc::idesc_tf32(128, c.n), idesc_h = tc::idesc_tf32(128, c.n / 2);
        const uint64_t da0 = tc::smem_desc(tc::smem_u32(sm) + 19 * 16, lbo);
        const uint64_t db0 = tc::smem_desc(tc::smem_u32(sm) + 100 * 1024, lbo);
        if (tc::elect_one()) { tc::mma_tf32(tm, da0, db0, idesc, false); tc::mma_commit(&bar); }
        tc::mbar_wait(&bar, 0);
        long long t0 = clock64();
        uint64_t da = da0, db = db0;
        for (int i = 0; i < c.nmma; ++i) {
            if (tc::elect_one()) tc::mma_tf32(tm, da, db, (c.alt && (i & 1)) ? idesc_h : idesc, true);
            da += (2 * lbo) >> 4; db += (2 * lbo) >> 4;
            if ((i & 3) == 3) { da = da0; db = db0; }
        }
        long long t1 = clock64();
        if (tc::elect_one()) tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 1);
        long long t2 = clock64();
        if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; stop = 1; }
    } else if (warp >= 4 && warp < 4 + c.bgwarps) {
        // background traffic; warps 4.. (own TMEM lane quadrant = warp % 4)
        float acc = 0.f;
        float4* scratch = reinterpret_cast<float4*>(sm + 150 * 1024) + (warp - 4) * 64;
        int it = 0;
        while (!stop) {
            if (c.bg == 1) { scratch[lane] = make_float4(acc, 1.f, 2.f, 3.f); scratch[lane + 32] = make_float4(acc, 1.f, 2.f, 3.f); }
            else if (c.bg == 2) { float4 v = scratch[lane]; float4 w = scratch[lane + 32]; acc += v.x + w.y; }
            else if (c.bg == 3) { float v[8]; tc::tmem_ld8(tm + ((uint32_t)((warp & 3) * 32) << 16) + 128 + (it & 7) * 8, v); tc::tmem_ld_wait(); acc += v[0]; }
            else if (c.bg == 4) { acc += __ldcg(gsrc + ((it * 32 + lane + warp * 4096) & 0xfffff)); }
            else if (c.bg == 5) { acc = fmaf(acc, 1.0001f, 0.5f); acc = fmaf(acc, 1.0001f, 0.5f); acc = fmaf(acc, 1.0001f, 0.5f); acc = fmaf(acc, 1.0001f, 0.5f); }
            ++it;
        }
        if (acc == 123.456f) out[3] = it;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 256);
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    float* g; cudaMalloc(&g, 4 << 20); cudaMemset(g, 0, 4 << 20);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[] = {"none", "STS.128", "LDS.128", "tcgen05.ld", "LDG", "FFMA"};
    for (int n : {32, 64, 128})
        for (int alt = 0; alt < 2; ++alt)
            for (int bg = 0; bg < 6; ++bg)
                for (int bw : {4}) {
                    if (alt && bg > 0) continue;
                    Cfg c{n, 256, bg, bw, alt};
                    long long h[2];
                    bench<<<1, 288, 200 * 1024>>>(c, d, g);
                    cudaError_t e = cudaDeviceSynchronize();
                    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    printf("N=%3d alt=%d bg=%-10s x%d warps : issue %6.1f cyc/mma, total %6.1f cyc/mma  (%s)\n", n, alt, names[bg], bw, (double)h[0] / c.nmma,
                           (double)h[1] / c.nmma, cudaGetErrorString(e));
                }
    return 0;
}
