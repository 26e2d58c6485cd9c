// Probe: the MMA issue loop of block_tc.cu's conv2 (9 taps x 3 tiles x 4 k-steps x 2 MMAs: N = 64 and N = 32, shifted descriptors,
// descriptor arithmetic on 32-bit words) issued by ONE lane, timed with clock64 until the commit arrives: is the phase bound by the
// tensor pipe (216 MMAs x ~45 cycles) or by the issuing thread?  Variants: A as in the kernel, B with the k-loop unrolled by 4,
// C only the N = 64 MMAs (how much the second MMA of the pair costs), D only N = 32.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../ibinn_b200/csrc/tc.cuh"

#ifndef PW
#define PW 32
#endif
#ifndef PNT
#define PNT 3
#endif
constexpr int WIDTH = PW, NT = PNT, NKS = PW / 8, BTAPS = PW == 32 ? 9 : 2;   // (width 64: the 9 taps reuse 2 filter images, 288 KB do not fit)

__global__ void __launch_bounds__(320) probe(long long* cyc, int variant, int ROWS, int rnd, int boff, int aoff, const float* gsrc, int ncopy, int copy_bytes) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar, bar2, cpbar, tapbar[9];
    __shared__ uint32_t tmem_s;
    __shared__ uint32_t s_tapoff[9];
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* A_hi = sm + aoff;                            // 8 chunks x ROWS x 16
    unsigned char* A_lo = sm + aoff + (WIDTH / 4) * ROWS * 16;
    unsigned char* Bm = sm + 2 * (WIDTH / 4) * ROWS * 16 + 1024 + boff;                 // 9 taps x (8 chunks x 64 rows x 16 B)
    for (int i = tid; i < (2 * (WIDTH / 4) * ROWS * 16 + 2048 + BTAPS * (WIDTH / 4) * 2 * WIDTH * 16) / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = rnd ? __uint_as_float(0x3f000000u + ((i * 2654435761u) >> 9)) : 1.f;
    if (tid < 9) s_tapoff[tid] = (PW == 32 ? 18 : 10) + (tid / 3 - 1) * (PW == 32 ? 17 : 9) + (tid % 3 - 1);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_init(&bar2, 2); tc::mbar_init(&cpbar, 1); for (int i = 0; i < 9; ++i) tc::mbar_init(&tapbar[i], 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_s, 256);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = tmem_s;
    if (variant == 7) {
        // two issuing threads: warp 1 issues the x_hi MMAs (N = 64), warp 2 the x_lo MMAs (N = 32) into SEPARATE accumulator columns
        if ((warp == 1 || warp == 2) && tc::elect_one()) {
            const bool lo = warp == 2;
            const uint32_t lbo_a = ROWS * 16, lbo_b = 2 * WIDTH * 16;
            const uint32_t id = tc::idesc_tf32(128, lo ? WIDTH : 2 * WIDTH);
            const uint32_t xa = tc::smem_u32(lo ? A_lo : A_hi), ba = tc::smem_u32(Bm);
            long long t0 = clock64();
            for (int tap = 0; tap < 9; ++tap)
                for (int t = 0; t < 3; ++t) {
                    const uint32_t ro = (uint32_t)(t * 128) + s_tapoff[tap];
                    const uint64_t da = tc::smem_desc(xa + ro * 16, lbo_a), db = tc::smem_desc(ba + tap * 8 * 64 * 16, lbo_b);
                    const uint32_t hA = (uint32_t)(da >> 32), hB = (uint32_t)(db >> 32);
                    const uint32_t a = (uint32_t)da, b = (uint32_t)db, a_step = (2 * lbo_a) >> 4, b_step = (2 * lbo_b) >> 4;
                    const uint32_t acc_addr = tm + (lo ? 192 + t * 32 : t * 64);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        tc::mma_tf32(acc_addr, ((uint64_t)hA << 32) | (a + ks * a_step), ((uint64_t)hB << 32) | (b + ks * b_step), id, (ks || tap) ? 1u : 0u);
                }
            long long t1 = clock64();
            tc::mma_commit(&bar2);
            tc::mbar_wait(&bar2, 0);
            if (blockIdx.x == gridDim.x - 1 && !lo) { cyc[0] = t1 - t0; cyc[1] = clock64() - t0; }
        }
    } else if (warp == 1 && tc::elect_one()) {
        const uint32_t lbo_a = ROWS * 16, lbo_b = 2 * WIDTH * 16;
        const uint32_t id2 = tc::idesc_tf32(128, 2 * WIDTH), id1 = tc::idesc_tf32(128, WIDTH);
        const uint32_t xh = tc::smem_u32(A_hi), xl = tc::smem_u32(A_lo), ba = tc::smem_u32(Bm);
        long long t0 = clock64();
        for (int tap = 0; tap < 9; ++tap)
            for (int t = 0; t < NT; ++t) {
                const uint32_t ro = (uint32_t)(t * 128) + s_tapoff[tap];
                const uint64_t dah = tc::smem_desc(xh + ro * 16, lbo_a), dal = tc::smem_desc(xl + ro * 16, lbo_a), db = tc::smem_desc(ba + (tap % BTAPS) * (WIDTH / 4) * 2 * WIDTH * 16, lbo_b);
                const uint32_t hA = (uint32_t)(dah >> 32), hB = (uint32_t)(db >> 32);
                uint32_t ah = (uint32_t)dah, al = (uint32_t)dal, b = (uint32_t)db;
                const uint32_t a_step = (2 * lbo_a) >> 4, b_step = (2 * lbo_b) >> 4;
                uint32_t acc = tap == 0 ? 0u : 1u;
                const uint32_t acc_addr = tm + t * 2 * WIDTH;
                if (variant == 1) {
#pragma unroll
                    for (int ks = 0; ks < NKS; ++ks) {
                        tc::mma_tf32(acc_addr, ((uint64_t)hA << 32) | (ah + ks * a_step), ((uint64_t)hB << 32) | (b + ks * b_step), id2, ks ? 1u : acc);
                        tc::mma_tf32(acc_addr, ((uint64_t)hA << 32) | (al + ks * a_step), ((uint64_t)hB << 32) | (b + ks * b_step), id1, 1);
                    }
                } else {
                    for (int ks = 0; ks < NKS; ++ks) {
                        if (variant != 3) tc::mma_tf32(acc_addr, ((uint64_t)hA << 32) | ah, ((uint64_t)hB << 32) | b, id2, acc);
                        if (variant != 2) tc::mma_tf32(acc_addr, ((uint64_t)hA << 32) | al, ((uint64_t)hB << 32) | b, id1, 1);
                        acc = 1;
                        ah += a_step; al += a_step; b += b_step;
                    }
                }
                if (variant == 4 && t == NT - 1) tc::mma_commit(&tapbar[tap]);      // one commit per filter tap, like the ring release
            }
        long long t1 = clock64();
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        if (blockIdx.x == gridDim.x - 1) { cyc[0] = t1 - t0; cyc[1] = clock64() - t0; }
    }
    // concurrent TMA traffic: `ncopy` bulk copies of `copy_bytes` from global memory into a scratch region while the MMAs run
    if (ncopy > 0 && warp == 3 && (tid & 31) == 0) {
        unsigned char* scratch = sm + 2 * (WIDTH / 4) * ROWS * 16 + 1024 + BTAPS * (WIDTH / 4) * 2 * WIDTH * 16 + 1024;
        for (int i = 0; i < ncopy; ++i) {
            tc::mbar_arrive_expect_tx(&cpbar, (uint32_t)copy_bytes);
            tc::bulk_g2s(scratch, gsrc + (size_t)(i % 8) * (copy_bytes / 4), (uint32_t)copy_bytes, &cpbar);
            tc::mbar_wait(&cpbar, i & 1);
        }
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
    const int grid = argc > 1 ? atoi(argv[1]) : 1, rows = argc > 2 ? atoi(argv[2]) : 420, rnd = argc > 3 ? atoi(argv[3]) : 0, boff = argc > 4 ? atoi(argv[4]) : 0, aoff = argc > 5 ? atoi(argv[5]) : 0, ncopy = argc > 6 ? atoi(argv[6]) : 0, cbytes = argc > 7 ? atoi(argv[7]) : 8192;
    float* gsrc; cudaMalloc(&gsrc, 8 * 32768); cudaMemset(gsrc, 0, 8 * 32768);
    printf("grid %d rows %d (lbo %d B) random data %d, B operand offset %d B, A operand offset %d B, %d concurrent bulk copies of %d B\n", grid, rows, rows * 16, rnd, boff, aoff, ncopy, cbytes);
    long long* c; cudaMalloc(&c, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[] = {"as in block_tc.cu (N=64 + N=32 per k-step)", "k-loop unrolled, descriptors by constant offsets", "only the N=64 MMAs", "only the N=32 MMAs", "as in block_tc.cu + a commit per tap", "as in block_tc.cu + 9 warps polling the mbarrier", "as in block_tc.cu + one lane per warp polling", "two issuing threads (hi / lo), separate accumulators"};
    for (int v = 0; v < 1; v += 7)
        for (int rep = 0; rep < 2; ++rep) {
            probe<<<grid, 320, 200 * 1024>>>(c, v, rows, rnd, boff, aoff, gsrc, ncopy, cbytes);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2]; cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
            const int n = ((v == 2 || v == 3) ? 1 : 2) * 9 * PNT * (PW / 8);
            if (rep) printf("%-52s %s: issue %lld cycles, until complete %lld cycles = %.1f per MMA (%d MMAs)\n", names[v], cudaGetErrorString(e), h[0], h[1], (double)h[1] / n, n);
        }
    return 0;
}
