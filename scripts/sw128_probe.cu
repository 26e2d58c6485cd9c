// Probe: K-major A operand in the SWIZZLE_128B layout (one 128-byte row of 32 tf32 per position, 16-byte chunk c of row r stored at
// chunk c ^ (r & 7)) read through descriptors that START AT AN ARBITRARY ROW (the nine shifted filter taps of the convolution
// kernels), against the no-swizzle layout used today (8 rows x 16 bytes = one 128-byte core matrix, misaligned for 8 of 9 taps).
// Prints, per row shift and base_offset choice, whether D = A[row + shift] came out, and the cycles per MMA of both layouts.
#include <cstdio>
#include <cuda_runtime.h>
#include "../ibinn_b200/csrc/tc.cuh"

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t base_off) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
}

constexpr int ROWS = 320;

__global__ void __launch_bounds__(128) probe(float* out, long long* cyc, int shift, int mode, int reps) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* A_sw = sm;                       // [ROWS][128 B] swizzled
    unsigned char* A_ns = sm + 48 * 1024;           // no swizzle: chunk c at c * (ROWS * 16) + r * 16
    unsigned char* Bm = sm + 96 * 1024;             // B: N = 64 rows, K = 32, no swizzle, LBO = 64 * 16
    for (int e = tid; e < ROWS * 8; e += 128) {
        const int r = e / 8, c = e % 8;
        float v[4];
        for (int j = 0; j < 4; ++j) v[j] = (float)((r % 64) * 32 + c * 4 + j);
        *reinterpret_cast<float4*>(A_sw + r * 128 + ((c ^ (r & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(A_ns + c * (ROWS * 16) + r * 16) = make_float4(v[0], v[1], v[2], v[3]);
    }
    for (int e = tid; e < 64 * 8; e += 128) {
        const int n = e / 8, c = e % 8;
        float v[4];
        for (int j = 0; j < 4; ++j) v[j] = (n < 32 && c * 4 + j == n) ? 1.f : 0.f;
        *reinterpret_cast<float4*>(Bm + c * (64 * 16) + n * 16) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_s, 64);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = tmem_s;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, 64);
        const uint64_t db0 = tc::smem_desc(tc::smem_u32(Bm), 64 * 16);
        long long t0 = clock64();
        for (int it = 0; it < reps; ++it) {
            for (int ks = 0; ks < 4; ++ks) {
                uint64_t da;
                if (mode == 0) da = tc::smem_desc(tc::smem_u32(A_ns) + shift * 16 + ks * 2 * (ROWS * 16), ROWS * 16);
                else da = desc_sw128(tc::smem_u32(A_sw) + shift * 128 + ks * 32, mode == 2 ? (uint32_t)(shift & 7) : 0u);
                const uint64_t db = db0 + (uint64_t)((ks * 2 * 64 * 16) >> 4);
                tc::mma_tf32(tm, da, db, idesc, ks > 0);
            }
        }
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        cyc[0] = clock64() - t0;
    }
    __syncthreads();
    tc::fence_after_sync();
    for (int c8 = 0; c8 < 32; c8 += 8) {
        float v[8];
        tc::tmem_ld8(tm + ((uint32_t)(warp * 32) << 16) + c8, v);
        tc::tmem_ld_wait();
        for (int j = 0; j < 8; ++j) out[tid * 32 + c8 + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 64);
}

int main() {
    float* d; long long* c;
    cudaMalloc(&d, 128 * 32 * 4); cudaMalloc(&c, 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    const int shifts[] = {0, 1, 2, 5, 8, 9, 17, 18};
    const char* names[] = {"no-swizzle", "SW128 base_offset=0", "SW128 base_offset=shift&7"};
    for (int mode = 0; mode < 3; ++mode)
        for (int s : shifts) {
            probe<<<1, 128, 128 * 1024>>>(d, c, s, mode, 1);
            cudaError_t e = cudaDeviceSynchronize();
            float h[128 * 32];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < 32; ++n) {
                    const float want = (float)(((r + s) % 64) * 32 + n);
                    if (h[r * 32 + n] != want) { if (bad < 2) printf("      row %d col %d: got %.0f want %.0f\n", r, n, h[r * 32 + n], want); ++bad; }
                }
            probe<<<1, 128, 128 * 1024>>>(d, c, s, mode, 500);
            cudaDeviceSynchronize();
            long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
            printf("%-28s shift %2d: %s, %4d mismatches, %.1f cycles per MMA (M=128 N=64 K=8)\n", names[mode], s, cudaGetErrorString(e), bad, (double)cy / (500 * 4));
        }
    return 0;
}
