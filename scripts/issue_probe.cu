// Probe: the MMA issue loop of block_tc.cu's conv2 (9 taps x 3 tiles x 4 k-steps x 2 MMAs: N = 64 and N = 32, shifted descriptors,
// descriptor arithmetic on 32-bit words) issued by ONE lane, timed with clock64 until the commit arrives: is the phase bound by the
// tensor pipe (216 MMAs x ~45 cycles) or by the issuing thread?  Variants: A as in the kernel, B with the k-loop unrolled by 4,
// C only the N = 64 MMAs (how much the second MMA of the pair costs), D only N = 32.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../ibinn_b200/csrc/tc.cuh"

constexpr int MAXROWS = 420, WIDTH = 32;

__global__ void __launch_bounds__(320) probe(long long* cyc, int variant, int ROWS, int rnd) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar, tapbar[9];
    __shared__ uint32_t tmem_s;
    __shared__ uint32_t s_tapoff[9];
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* A_hi = sm;                                   // 8 chunks x ROWS x 16
    unsigned char* A_lo = sm + 8 * ROWS * 16;
    unsigned char* Bm = sm + 2 * 8 * ROWS * 16;                 // 9 taps x (8 chunks x 64 rows x 16 B)
    for (int i = tid; i < (2 * 8 * ROWS * 16 + 9 * 8 * 64 * 16) / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = rnd ? __uint_as_float(0x3f000000u + ((i * 2654435761u) >> 9)) : 1.f;
    if (tid < 9) s_tapoff[tid] = 18 + (tid / 3 - 1) * 17 + (tid % 3 - 1);
    if (tid == 0) { tc::mbar_init(&bar, 1); for (int i = 0; i < 9; ++i) tc::mbar_init(&tapbar[i], 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_s, 256);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = tmem_s;
    if (warp == 1 && tc::elect_one()) {
        const uint32_t lbo_a = ROWS * 16, lbo_b = 2 * WIDTH * 16;
        const uint32_t id2 = tc::idesc_tf32(128, 2 * WIDTH), id1 = tc::idesc_tf32(128, WIDTH);
        const uint32_t xh = tc::smem_u32(A_hi), xl = tc::smem_u32(A_lo), ba = tc::smem_u32(Bm);
        long long t0 = clock64();
        for (int tap = 0; tap < 9; ++tap)
            for (int t = 0; t < 3; ++t) {
                const uint32_t ro = (uint32_t)(t * 128) + s_tapoff[tap];
                const uint64_t dah = tc::smem_desc(xh + ro * 16, lbo_a), dal = tc::smem_desc(xl + ro * 16, lbo_a), db = tc::smem_desc(ba + tap * 8 * 64 * 16, lbo_b);
                const uint32_t hA = (uint32_t)(dah >> 32), hB = (uint32_t)(db >> 32);
                uint32_t ah = (uint32_t)dah, al = (uint32_t)dal, b = (uint32_t)db;
                const uint32_t a_step = (2 * lbo_a) >> 4, b_step = (2 * lbo_b) >> 4;
                uint32_t acc = tap == 0 ? 0u : 1u;
                const uint32_t acc_addr = tm + t * 64;
                if (variant == 1) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        tc::mma_tf32(acc_addr, ((uint64_t)hA << 32) | (ah + ks * a_step), ((uint64_t)hB << 32) | (b + ks * b_step), id2, ks ? 1u : acc);
                        tc::mma_tf32(acc_addr, ((uint64_t)hA << 32) | (al + ks * a_step), ((uint64_t)hB << 32) | (b + ks * b_step), id1, 1);
                    }
                } else {
                    for (int ks = 0; ks < 4; ++ks) {
                        if (variant != 3) tc::mma_tf32(acc_addr, ((uint64_t)hA << 32) | ah, ((uint64_t)hB << 32) | b, id2, acc);
                        if (variant != 2) tc::mma_tf32(acc_addr, ((uint64_t)hA << 32) | al, ((uint64_t)hB << 32) | b, id1, 1);
                        acc = 1;
                        ah += a_step; al += a_step; b += b_step;
                    }
                }
                if (variant == 4 && t == 2) tc::mma_commit(&tapbar[tap]);      // one commit per filter tap, like the ring release
            }
        long long t1 = clock64();
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        if (blockIdx.x == gridDim.x - 1) { cyc[0] = t1 - t0; cyc[1] = clock64() - t0; }
    }
    // variants 5 / 6: the other warps POLL the completion barrier while the MMAs run (256 threads, or one lane per warp), as the
    // epilogue warps of block_tc.cu do on acc_full
    if (variant == 5 && warp != 1) tc::mbar_wait(&bar, 0);
    if (variant == 6 && warp != 1) { if ((tid & 31) == 0) tc::mbar_wait(&bar, 0); __syncwarp(); }
    __syncthreads();
    tc::fence_after_sync();
    if (warp == 0) tc::tmem_dealloc(tm, 256);
}

int main(int argc, char** argv) {
    const int grid = argc > 1 ? atoi(argv[1]) : 1, rows = argc > 2 ? atoi(argv[2]) : 420, rnd = argc > 3 ? atoi(argv[3]) : 0;
    printf("grid %d rows %d (lbo %d B) random data %d\n", grid, rows, rows * 16, rnd);
    long long* c; cudaMalloc(&c, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[] = {"as in block_tc.cu (N=64 + N=32 per k-step)", "k-loop unrolled, descriptors by constant offsets", "only the N=64 MMAs", "only the N=32 MMAs", "as in block_tc.cu + a commit per tap", "as in block_tc.cu + 9 warps polling the mbarrier", "as in block_tc.cu + one lane per warp polling"};
    for (int v = 0; v < 2; ++v)
        for (int rep = 0; rep < 2; ++rep) {
            probe<<<grid, 320, 200 * 1024>>>(c, v, rows, rnd);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2]; cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
            const int n = (v == 2 || v == 3) ? 108 : 216;
            if (rep) printf("%-52s %s: issue %lld cycles, until complete %lld cycles = %.1f per MMA (%d MMAs)\n", names[v], cudaGetErrorString(e), h[0], h[1], (double)h[1] / n, n);
        }
    return 0;
}
