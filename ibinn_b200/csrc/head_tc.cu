// GMM head at the scoring batches (BASELINE.json configs[3]; reference model.py:113-126): the distance GEMM z . mu^T on
// tcgen05 as a 3xTF32 (fp32-accurate) contraction, z streamed ONCE from HBM.
//
//   zz_bk = |z_b - m|^2 - 2 (z_b - m).(mu_k - m) + |mu_k - m|^2,    m = mean_k mu_k
// (the reference's expanded form, model.py:122-126, centred on the mean of the class means so that a common offset of the
// class means does not cost digits).  One CTA = 128 samples x one K range (split-K over blockIdx.y):
//   warps 0-3  thread <-> sample row: global z -> registers (32 KB in flight per CTA) -> minus m, |.|^2 accumulated in fp32
//              -> hi / lo split -> K-major operand rows in a 3-stage shared-memory ring; afterwards the epilogue
//   warp 4     MMA issuer: per k-step  z_hi [mu_hi | mu_lo] (N = 2 Cp)  and  z_lo mu_hi (N = Cp)  into TMEM
//   warp 5     streams the prepared class-mean image (hi / lo stacked along N, ibinn_wprep layout) into the same ring by TMA
// The partial dot products / norms of the K ranges go to a workspace; head_gemm_finish_kernel adds them in fixed order and runs
// the same per-sample loss tail as the CUDA-core kernels of head.cu (head_tail.cuh).
#include "common.cuh"
#include "tc.cuh"
#include "head_tail.cuh"

#ifndef IBINN_EMU
namespace {

constexpr int HT_THREADS = 320;      // warps 0-7 loaders (two threads per sample row), warp 8 MMA issuer, warp 9 class-mean producer
constexpr int HT_NST = 3;

struct HG {
    int np, KC, nsplit, kper, nchunk;       // classes padded to 16; d-values per stage; K ranges; d-values per range; stages per range
    int z_bytes, mu_bytes, stage_bytes, tmem_cols;
    size_t smem_bytes;
};

static bool hg_geometry(int B, int C, int D, HG& g) {
    if (C < 1 || C > 112 || D % 64 != 0) return false;
    g.np = (C + 15) & ~15;
    g.KC = g.np <= 32 ? 64 : 32;
    const int ntile = (B + 127) / 128;
    int ns = 1;
    while (ntile * ns * 2 <= 148 && ns < 8 && D % (g.KC * ns * 2) == 0) ns *= 2;      // fill the SMs without a second wave
    g.nsplit = ns;
    g.kper = D / ns;
    if (g.kper % g.KC) return false;
    g.nchunk = g.kper / g.KC;
    g.z_bytes = (g.KC / 4) * 128 * 16;                 // ONE of hi / lo
    g.mu_bytes = (g.KC / 4) * 2 * g.np * 16;
    g.stage_bytes = 2 * g.z_bytes + g.mu_bytes;
    g.tmem_cols = 2 * g.np <= 32 ? 32 : 2 * g.np <= 64 ? 64 : 2 * g.np <= 128 ? 128 : 256;
    g.smem_bytes = (size_t)HT_NST * g.stage_bytes + (size_t)g.kper * 4 + 256;
    return g.smem_bytes <= 225 * 1024;
}

// ---- preparation (per launch; the class means change every training step, and the scoring path calls this once per batch):
// m = mean over the classes, the centred class means as a tensor-core operand image, their squared norms
__global__ void __launch_bounds__(64) head_gemm_prep_kernel(const float* __restrict__ mu, float* __restrict__ m, float* __restrict__ img, int C, int D, int np) {
    const int k4 = blockIdx.x * 64 + threadIdx.x;         // one 16-byte chunk (4 consecutive d) per thread
    if (k4 * 4 >= D) return;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < C; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(mu + (size_t)k * D + k4 * 4);
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    }
    const float inv = 1.0f / (float)C;
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] *= inv;
    *reinterpret_cast<float4*>(m + k4 * 4) = make_float4(s[0], s[1], s[2], s[3]);
    float* dst = img + (size_t)k4 * 2 * np * 4;
    for (int n = 0; n < np; ++n) {
        float hi[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
        if (n < C) {
            const float4 v = *reinterpret_cast<const float4*>(mu + (size_t)n * D + k4 * 4);
            const float c[4] = {v.x - s[0], v.y - s[1], v.z - s[2], v.w - s[3]};
#pragma unroll
            for (int i = 0; i < 4; ++i) tc::split_tf32(c[i], hi[i], lo[i]);
        }
        *reinterpret_cast<float4*>(dst + (size_t)n * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(dst + (size_t)(np + n) * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

__global__ void __launch_bounds__(256) head_gemm_musq_kernel(const float* __restrict__ mu, const float* __restrict__ m, float* __restrict__ musq, int C, int D) {
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;       // one warp per class, fixed order
    if (k >= C) return;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float c = mu[(size_t)k * D + d] - m[d];
        s = fmaf(c, c, s);
    }
    s = warp_sum(s);
    if (lane == 0) musq[k] = s;
}

__device__ __forceinline__ void ht_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(HT_THREADS, 1) head_gemm_kernel(const float* __restrict__ z, const float* __restrict__ m, const float* __restrict__ img,
                                                                  float* __restrict__ ws_dot, float* __restrict__ ws_zsq, int B, int C, int D, const HG g) {
    extern __shared__ __align__(128) unsigned char hs[];
    __shared__ __align__(8) uint64_t full[HT_NST], empty[HT_NST], acc_full;
    __shared__ uint32_t tmem_base_s;
    float* s_m = reinterpret_cast<float*>(hs + (size_t)HT_NST * g.stage_bytes);       // [kper] slice of m
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int tile = blockIdx.x, split = blockIdx.y, k_begin = split * g.kper;

    if (tid == 0) {
        for (int i = 0; i < HT_NST; ++i) { tc::mbar_init(&full[i], 257); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(&acc_full, 1);
        tc::mbar_fence_init();
    }
    if (warp == 8) tc::tmem_alloc(&tmem_base_s, g.tmem_cols);
    for (int i = tid; i < g.kper; i += HT_THREADS) s_m[i] = m[k_begin + i];
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;

    if (warp == 9) {
        if (lane == 0) {
            for (int c = 0; c < g.nchunk; ++c) {
                const int s = c % HT_NST;
                if (c >= HT_NST) tc::mbar_wait(&empty[s], ((c / HT_NST) - 1) & 1);
                tc::mbar_arrive_expect_tx(&full[s], (uint32_t)g.mu_bytes);
                tc::bulk_g2s(hs + (size_t)s * g.stage_bytes + 2 * g.z_bytes, img + ((size_t)(k_begin + c * g.KC) / 4) * 2 * g.np * 4, (uint32_t)g.mu_bytes, &full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 8) {
        if (tc::elect_one()) {
            const uint32_t id2 = tc::idesc_tf32(128, 2 * g.np), id1 = tc::idesc_tf32(128, g.np);
            uint32_t acc = 0;
            for (int c = 0; c < g.nchunk; ++c) {
                const int s = c % HT_NST;
                tc::mbar_wait(&full[s], (c / HT_NST) & 1);
                tc::fence_after_sync();
                const uint32_t a0 = tc::smem_u32(hs + (size_t)s * g.stage_bytes);
                const uint64_t dah = tc::smem_desc(a0, 128 * 16), dal = tc::smem_desc(a0 + g.z_bytes, 128 * 16);
                const uint64_t db = tc::smem_desc(a0 + 2 * g.z_bytes, 2 * g.np * 16);
                const uint32_t hA = (uint32_t)(dah >> 32), hB = (uint32_t)(db >> 32);
                uint32_t ah = (uint32_t)dah, al = (uint32_t)dal, b = (uint32_t)db;
                const uint32_t a_step = (2 * 128 * 16) >> 4, b_step = (2 * 2 * g.np * 16) >> 4;
                for (int ks = 0; ks < g.KC / 8; ++ks) {
                    tc::mma_tf32(tmem_d, ((uint64_t)hA << 32) | ah, ((uint64_t)hB << 32) | b, id2, acc);
                    tc::mma_tf32(tmem_d, ((uint64_t)hA << 32) | al, ((uint64_t)hB << 32) | b, id1, 1);
                    acc = 1;
                    ah += a_step; al += a_step; b += b_step;
                }
                tc::mma_commit(&empty[s]);
            }
            tc::mma_commit(&acc_full);
        }
        __syncwarp();
    } else {
        // ---- loaders: two threads per sample row, each takes half of the stage's 16-byte chunks; the loads of the NEXT stage are
        // in flight while the current one is converted
        const int row = tid & 127, kh = tid >> 7, b = tile * 128 + row;
        const bool live = b < B;
        const int nj = g.KC / 8;                                  // chunks per thread and stage (4 or 8)
        const float* zr = z + (size_t)(live ? b : 0) * D + k_begin + kh * nj * 4;
        float zsq = 0.f;
        float4 cur[8], nxt[8], nx2[8];                            // two stages of loads in flight behind the one being converted
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            cur[j] = (j < nj && live) ? __ldg(reinterpret_cast<const float4*>(zr) + j) : z4;
            nxt[j] = (j < nj && live && 1 < g.nchunk) ? __ldg(reinterpret_cast<const float4*>(zr + g.KC) + j) : z4;
        }
        for (int c = 0; c < g.nchunk; ++c) {
            const int s = c % HT_NST;
#pragma unroll
            for (int j = 0; j < 8; ++j) nx2[j] = (j < nj && live && c + 2 < g.nchunk) ? __ldg(reinterpret_cast<const float4*>(zr + (c + 2) * g.KC) + j) : z4;
            if (c >= HT_NST) tc::mbar_wait(&empty[s], ((c / HT_NST) - 1) & 1);
            unsigned char* dst = hs + (size_t)s * g.stage_bytes + row * 16 + (size_t)(kh * nj) * 128 * 16;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nj) {
                    const float4 mm = *reinterpret_cast<const float4*>(s_m + c * g.KC + (kh * nj + j) * 4);
                    float e[4] = {cur[j].x - mm.x, cur[j].y - mm.y, cur[j].z - mm.z, cur[j].w - mm.w};
                    float hi[4], lo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (!live) e[i] = 0.f;
                        zsq = fmaf(e[i], e[i], zsq);
                        tc::split_tf32(e[i], hi[i], lo[i]);
                    }
                    *reinterpret_cast<float4*>(dst + (size_t)j * 128 * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(dst + g.z_bytes + (size_t)j * 128 * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            tc::fence_async_smem();
            ht_arrive(&full[s]);
#pragma unroll
            for (int j = 0; j < 8; ++j) { cur[j] = nxt[j]; nxt[j] = nx2[j]; }
        }
        if (live) ws_zsq[((size_t)split * 2 + kh) * B + b] = zsq;
        // ---- epilogue: partial dot products of this K range
        tc::mbar_wait(&acc_full, 0);
        tc::fence_after_sync();
        const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
        float* dd = ws_dot + ((size_t)split * B + (live ? b : 0)) * C;
        for (int c8 = kh * 8; c8 < g.np; c8 += 16) {             // the two threads of a row take alternate groups of 8 classes
            float a8[8], b8[8];
            tc::tmem_ld8(taddr + c8, a8);
            tc::tmem_ld8(taddr + g.np + c8, b8);
            tc::tmem_ld_wait();
            if (live)
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (c8 + i < C) dd[c8 + i] = a8[i] + b8[i];
        }
        tc::fence_before_sync();
    }
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem_d, g.tmem_cols);
}

// per sample: distances from the partial results (fixed order over the K ranges), then the loss tail shared with head.cu
__global__ void __launch_bounds__(IBINN_THREADS) head_gemm_finish_kernel(const HeadP p, const float* __restrict__ ws_dot, const float* __restrict__ ws_zsq,
                                                                         const float* __restrict__ musq, int nsplit) {
    extern __shared__ __align__(16) float fs[];
    __shared__ int s_flag;
    const int C = p.C, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = IBINN_THREADS / 32;
    float* s_lw = fs;                       // [C]
    float* s_zz = fs + C + (size_t)wid * C; // [C] per warp
    log_softmax_phi(p.phi, C, s_lw);
    __syncthreads();
    for (int b = blockIdx.x * nw + wid; b < p.B; b += gridDim.x * nw) {
        float zs = 0.f;
        for (int s = 0; s < 2 * nsplit; ++s) zs += ws_zsq[(size_t)s * p.B + b];          // two row halves per K range
        for (int k = lane; k < C; k += 32) {
            float dot = 0.f;
            for (int s = 0; s < nsplit; ++s) dot += ws_dot[((size_t)s * p.B + b) * C + k];
            s_zz[k] = fmaxf(fmaf(-2.f, dot, zs) + musq[k], 0.f);
        }
        __syncwarp();
        head_sample_tail(p, b, s_zz, s_lw, lane);
        __syncwarp();
    }
    if (last_block_ticket(p.counter, gridDim.x, &s_flag)) {
        head_batch_means<IBINN_THREADS>(p);
        if (tid == 0) *p.counter = 0u;
    }
}

}  // namespace

extern "C" int ibinn_gmm_head_gemm_supported(int B, int C, int D) {
    HG g;
    return (B >= 1024 && hg_geometry(B, C, D, g)) ? 1 : 0;
}

extern "C" int64_t ibinn_gmm_head_gemm_workspace_floats(int B, int C, int D) {
    HG g;
    if (!hg_geometry(B, C, D, g)) return IBINN_EINVAL;
    // class-mean image, m, |mu - m|^2, partial dot products, partial norms
    return (int64_t)(D / 4) * 2 * g.np * 4 + D + 128 + (int64_t)g.nsplit * B * C + (int64_t)2 * g.nsplit * B + 64;
}

extern "C" int ibinn_gmm_head_forward_gemm(const float* z, const float* jac, const float* mu, const float* phi, const float* y, float* L_x,
                                           float* logits, float* L_cNLL, float* L_y, float* correct, float* lse, float* means,
                                           float* partials, uint32_t* counter, float* workspace, int mu_prepared, int B, int C, int D,
                                           ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!z || !jac || !mu || !phi || !L_x || !logits || !lse || !means || !partials || !counter || !workspace) return IBINN_ENULL;
    if (y && (!L_cNLL || !L_y || !correct)) return IBINN_ENULL;
    HG g;
    if (!hg_geometry(B, C, D, g)) return IBINN_EINVAL;
    if (((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(mu) | reinterpret_cast<uintptr_t>(workspace)) & 15) != 0) return IBINN_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    float* img = workspace;
    float* m = img + (size_t)(D / 4) * 2 * g.np * 4;
    float* musq = m + D;
    float* ws_dot = musq + 128;
    float* ws_zsq = ws_dot + (size_t)g.nsplit * B * C;
    if (!mu_prepared) {
        IBINN_LAUNCH(head_gemm_prep_kernel, dim3((D / 4 + 63) / 64), dim3(64), 0, st, mu, m, img, C, D, g.np);
        IBINN_LAUNCH(head_gemm_musq_kernel, dim3((C + 7) / 8), dim3(256), 0, st, mu, m, musq, C, D);
    }
    e = ibinn_set_smem(head_gemm_kernel, g.smem_bytes);
    if (e) return e;
    IBINN_LAUNCH(head_gemm_kernel, dim3((B + 127) / 128, g.nsplit), dim3(HT_THREADS), g.smem_bytes, st, z, m, img, ws_dot, ws_zsq, B, C, D, g);
    HeadP p{};
    p.z = z; p.jac = jac; p.mu = mu; p.phi = phi; p.y = y; p.Lx = L_x; p.logits = logits; p.Lc = L_cNLL; p.Ly = L_y;
    p.correct = correct; p.lse = lse; p.means = means; p.partials = partials; p.counter = counter; p.B = B; p.C = C; p.D = D;
    const size_t fsm = ((size_t)C + (size_t)(IBINN_THREADS / 32) * C) * sizeof(float);
    int nb = (B + IBINN_THREADS / 32 - 1) / (IBINN_THREADS / 32);
    if (nb > 148 * 4) nb = 148 * 4;
    IBINN_LAUNCH(head_gemm_finish_kernel, dim3(nb), dim3(IBINN_THREADS), fsm, st, p, ws_dot, ws_zsq, musq, g.nsplit);
    return ibinn_launch_status();
}

#else
extern "C" int ibinn_gmm_head_gemm_supported(int, int, int) { return 0; }
extern "C" int64_t ibinn_gmm_head_gemm_workspace_floats(int, int, int) { return IBINN_EINVAL; }
extern "C" int ibinn_gmm_head_forward_gemm(const float*, const float*, const float*, const float*, const float*, float*, float*, float*, float*, float*,
                                           float*, float*, float*, uint32_t*, float*, int, int, int, int, ibinn_stream_t) { return IBINN_EINVAL; }
#endif
