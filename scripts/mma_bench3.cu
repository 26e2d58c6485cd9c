// Micro-benchmark 3: the exact MMA issue loop of conv_tc_kernel (nine shifted descriptors, hi/lo stacked filters) in isolation.
#include <cstdio>
#include <cuda_runtime.h>
#include "../ibinn_b200/csrc/tc.cuh"

struct G { int cin_p, np, P, pad, kk, halo, npos, lbo_a, lbo_b, a_bytes, tap_bytes, variant, canon, nst; };

template <int CANON>
__global__ void __launch_bounds__(320) bench(G g, long long* out) {
    __shared__ uint32_t s_off[9];
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = CANON ? __shfl_sync(0xffffffffu, tid >> 5, 0) : tid >> 5;
    if (tid < 9) { const int dy = g.pad ? tid / 3 - 1 : 0, dx = g.pad ? tid % 3 - 1 : 0; s_off[tid] = (uint32_t)(g.halo + dy * g.P + dx); }
    for (int i = tid; i < 200 * 1024 / 4; i += 320) reinterpret_cast<float*>(sm)[i] = 1.0f;
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_s, 256);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_s;
    if (warp == 8) {
        unsigned char* wst = sm + 2 * g.a_bytes;
        const uint32_t idesc2 = tc::idesc_tf32(128, 2 * g.np), idesc1 = tc::idesc_tf32(128, g.np);
        const uint64_t db0 = tc::smem_desc(tc::smem_u32(wst), g.lbo_b);
        const int nks = g.cin_p / 8;
        const uint32_t a_step = (uint32_t)(2 * g.lbo_a) >> 4, b_step = (uint32_t)(2 * g.lbo_b) >> 4, st_step = (uint32_t)g.tap_bytes >> 4;
        const uint32_t a_base = tc::smem_u32(sm);
        const uint64_t da_hi = tc::smem_desc(a_base, g.lbo_a), da_lo = tc::smem_desc(a_base + g.a_bytes, g.lbo_a);
        for (int rep = 0; rep < 2; ++rep) {
            long long t0 = clock64();
            uint32_t acc = 0;
            if (g.variant == 0) {
                for (int t = 0; t < g.kk; ++t) {
                    const int dy = g.pad ? t / 3 - 1 : 0, dx = g.pad ? t % 3 - 1 : 0;
                    const uint32_t row_off = (uint32_t)(g.halo + dy * g.P + dx);
                    uint64_t dah = da_hi + row_off, dal = da_lo + row_off;
                    uint64_t db = db0 + (uint64_t)(t * st_step);
#pragma unroll 4
                    for (int ks = 0; ks < nks; ++ks) {
                        if (tc::elect_one()) {
                            tc::mma_tf32(tmem_d, dah, db, idesc2, acc != 0);
                            tc::mma_tf32(tmem_d, dal, db, idesc1, 1);
                        }
                        acc = 1;
                        dah += a_step; dal += a_step; db += b_step;
                    }
                }
            } else if (g.variant == 1) {
                const uint32_t hA = (uint32_t)(da_hi >> 32), hB = (uint32_t)(db0 >> 32);
                const uint32_t lAh = (uint32_t)da_hi, lAl = (uint32_t)da_lo, lB = (uint32_t)db0;
                for (int t = 0; t < g.kk; ++t) {
                    const uint32_t row_off = s_off[t];
                    uint32_t ah = lAh + row_off, al = lAl + row_off, b = lB + t * st_step;
#pragma unroll 4
                    for (int ks = 0; ks < nks; ++ks) {
                        if (tc::elect_one()) {
                            tc::mma_tf32(tmem_d, ((uint64_t)hA << 32) | ah, ((uint64_t)hB << 32) | b, idesc2, acc != 0);
                            tc::mma_tf32(tmem_d, ((uint64_t)hA << 32) | al, ((uint64_t)hB << 32) | b, idesc1, 1);
                        }
                        acc = 1;
                        ah += a_step; al += a_step; b += b_step;
                    }
                }
            } else {
                // one elected lane runs the whole loop
                const uint32_t hA = (uint32_t)(da_hi >> 32), hB = (uint32_t)(db0 >> 32);
                const uint32_t lAh = (uint32_t)da_hi, lAl = (uint32_t)da_lo, lB = (uint32_t)db0;
                if (tc::elect_one()) {
                    for (int t = 0; t < g.kk; ++t) {
                        const uint32_t row_off = s_off[t];
                        uint32_t ah = lAh + row_off, al = lAl + row_off, b = lB + t * st_step;
#pragma unroll 4
                        for (int ks = 0; ks < nks; ++ks) {
                            tc::mma_tf32(tmem_d, ((uint64_t)hA << 32) | ah, ((uint64_t)hB << 32) | b, idesc2, acc != 0);
                            tc::mma_tf32(tmem_d, ((uint64_t)hA << 32) | al, ((uint64_t)hB << 32) | b, idesc1, 1);
                            acc = 1;
                            ah += a_step; al += a_step; b += b_step;
                        }
                    }
                }
                __syncwarp();
            }
            long long t1 = clock64();
            if (tc::elect_one()) tc::mma_commit(&bar);
            tc::mbar_wait(&bar, rep);
            long long t2 = clock64();
            if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, 256);
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct { int cin, cout, W, ks; } cs[] = {{16, 16, 32, 3}, {32, 32, 16, 3}, {48, 48, 8, 3}, {8, 32, 16, 3}, {64, 48, 8, 1}, {32, 16, 16, 1}};
    for (auto& c : cs)
        for (int variant = 0; variant < 6; ++variant) {
            G g;
            g.pad = c.ks / 2; g.kk = c.ks * c.ks; g.P = c.W + g.pad; g.halo = g.pad * (g.P + 1); g.npos = 128 + 2 * g.halo;
            g.cin_p = (c.cin + 7) & ~7; g.np = (c.cout + 15) & ~15; g.lbo_a = g.npos * 16; g.lbo_b = 2 * g.np * 16;
            g.a_bytes = (g.cin_p / 4) * g.lbo_a; g.tap_bytes = (g.cin_p / 4) * g.lbo_b; g.variant = variant % 3; g.canon = variant / 3;
            g.nst = g.kk; if (2 * g.a_bytes + g.kk * g.tap_bytes > 190 * 1024) continue;
            long long h[2];
            if (g.canon) bench<1><<<1, 320, 200 * 1024>>>(g, d); else bench<0><<<1, 320, 200 * 1024>>>(g, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            const int nm = g.kk * (g.cin_p / 8) * 2;
            printf("conv%d %2d->%2d W=%2d variant %d (0-2 tid>>5, 3-5 shfl-uniform warp id): %3d MMAs, issue %6lld cyc, total %6lld cyc = %6.1f cyc/mma (%s)\n", c.ks, c.cin, c.cout, c.W, variant, nm, h[0], h[1],
                   (double)h[1] / nm, cudaGetErrorString(e));
        }
    return 0;
}
