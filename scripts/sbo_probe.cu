// Probe: K-major no-swizzle A operand with a stride between 8-row groups (SBO) other than 128 bytes.  A strip of G
// (rows = co in groups of 8, chunk kappa at 128 (kappa G + g s) bytes) read with LBO = 128 G, SBO = 128 s makes row
// (dy, co) of the descriptor alias strip chunk (k + dy s): three row-shifted copies without copying.
#include <cstdio>
#include <cuda_runtime.h>
#include "../ibinn_b200/csrc/tc.cuh"

__device__ __forceinline__ uint64_t desc_sbo(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(128) probe(float* out, int G, int s) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* A = reinterpret_cast<float*>(sm);
    float* Bm = reinterpret_cast<float*>(sm + 96 * 1024);
    for (int i = tid; i < 128 * 1024 / 4; i += 128) A[i] = 0.f;
    __syncthreads();
    const int coutp = 8 * G, nk = 2 + 2 * s + 2;
    // strip: value(co, kappa, j) = 100 kappa + co + 0.25 j  (j = pixel within the chunk)
    for (int e = tid; e < coutp * nk; e += 128) {
        const int co = e % coutp, kap = e / coutp;
        float* dst = reinterpret_cast<float*>(sm + 128 * (kap * G + (co / 8) * s) + (co % 8) * 16);
        for (int j = 0; j < 4; ++j) dst[j] = 100.f * kap + co + 0.25f * j;
    }
    // B: N = 8 rows, K = 8: B[n][kk] = (kk == n) -> D[r][n] = A[r][kk = n]
    for (int n = tid; n < 8; n += 128)
        for (int kk = 0; kk < 8; ++kk) Bm[((kk / 4) * (9 * 16) + n * 16) / 4 + (kk % 4)] = (kk == n) ? 1.f : 0.f;
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_s, 32);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = tmem_s;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, 8);
        const uint64_t da = desc_sbo(tc::smem_u32(sm), 128 * G, 128 * s), db = tc::smem_desc(tc::smem_u32(Bm), 9 * 16);
        tc::mma_tf32(tm, da, db, idesc, false);
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
    }
    __syncthreads();
    tc::fence_after_sync();
    float v[8];
    tc::tmem_ld8(tm + ((uint32_t)(warp * 32) << 16), v);
    tc::tmem_ld_wait();
    for (int j = 0; j < 8; ++j) out[tid * 8 + j] = v[j];
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 32);
}

int main() {
    float* d; cudaMalloc(&d, 128 * 8 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int cfg[3][2] = {{4, 5}, {2, 9}, {8, 3}};
    for (auto& c : cfg) {
        const int G = c[0], s = c[1], coutp = 8 * G;
        probe<<<1, 128, 160 * 1024>>>(d, G, s);
        cudaError_t e = cudaDeviceSynchronize();
        float h[128 * 8];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < 128 && r < 3 * coutp; ++r) {
            const int dy = r / coutp, co = r % coutp;
            for (int n = 0; n < 8; ++n) {
                const int kap = n / 4 + dy * s, j = n % 4;
                const float want = 100.f * kap + co + 0.25f * j;
                if (h[r * 8 + n] != want) { if (bad < 5) printf("   row %d col %d: got %.2f want %.2f\n", r, n, h[r * 8 + n], want); ++bad; }
            }
        }
        printf("G=%d s=%d (LBO=%d SBO=%d): %s, %d mismatches\n", G, s, 128 * G, 128 * s, cudaGetErrorString(e), bad);
    }
    return 0;
}
