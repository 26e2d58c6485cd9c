// Per-sample loss tail and batch means of the GMM / IB head (reference model.py:144-163), shared by the CUDA-core kernels of
// head.cu and the tensor-core distance GEMM of head_tc.cu.
#pragma once
#include "common.cuh"

namespace {

struct HeadP {
    const float* z; const float* jac; const float* mu; const float* phi; const float* y;
    float* Lx; float* logits; float* Lc; float* Ly; float* correct; float* lse; float* means;
    float* partials; unsigned* counter;
    int B, C, D;
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// log_softmax(phi) into s_lw (C floats); call with all threads, then __syncthreads().
__device__ __forceinline__ void log_softmax_phi(const float* phi, int C, float* s_lw) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float m = -INFINITY;
        for (int k = lane; k < C; k += 32) m = fmaxf(m, __ldg(phi + k));
        m = warp_max(m);
        float s = 0.f;
        for (int k = lane; k < C; k += 32) s += expf(__ldg(phi + k) - m);
        s = warp_sum(s);
        const float l = m + logf(s);
        for (int k = lane; k < C; k += 32) s_lw[k] = __ldg(phi + k) - l;
    }
}

// Loss terms of one sample from its C squared distances zz (shared memory) by one warp: model.py:144-159.
// The five per-sample quantities that feed the batch means go to p.partials[b] (wsum == nullptr) or are added to wsum[0..5) (lane 0).
__device__ __forceinline__ void head_sample_tail(const HeadP& p, int b, const float* zz, const float* s_lw, int lane, double* wsum = nullptr) {
    const int C = p.C, D = p.D;
    const float jac = __ldg(p.jac + b);
    float m = -INFINITY;
    for (int k = lane; k < C; k += 32) m = fmaxf(m, -0.5f * zz[k] + s_lw[k]);
    m = warp_max(m);
    float se = 0.f;
    for (int k = lane; k < C; k += 32) se += expf(-0.5f * zz[k] + s_lw[k] - m);
    se = warp_sum(se);
    const float ls = logf(se);
    const float lse = m + ls;
    float sl = 0.f;
    for (int k = lane; k < C; k += 32) {
        const float lg = -0.5f * zz[k];
        p.logits[(size_t)b * C + k] = lg;
        sl += lg;
    }
    sl = warp_sum(sl);
    float lc = 0.f, ly = 0.f;
    float by = -INFINITY, bl = -INFINITY; int iy = 0x7fffffff, il = 0x7fffffff;
    if (p.y) {
        for (int k = lane; k < C; k += 32) {
            const float yk = __ldg(p.y + (size_t)b * C + k);
            const float lg = -0.5f * zz[k];
            lc = fmaf(zz[k], rintf(yk), lc);
            // (log_softmax(l)_k - logw_k) evaluated like the reference: shift by the max first
            ly = fmaf(yk, ((lg + s_lw[k] - m) - ls) - s_lw[k], ly);
            if (yk > by) { by = yk; iy = k; }
            if (lg > bl) { bl = lg; il = k; }
        }
        lc = warp_sum(lc);
        ly = warp_sum(ly);
        // arg max with first-index tie break (torch.max semantics)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float oy = __shfl_xor_sync(0xffffffffu, by, o); const int oiy = __shfl_xor_sync(0xffffffffu, iy, o);
            if (oy > by || (oy == by && oiy < iy)) { by = oy; iy = oiy; }
            const float ol = __shfl_xor_sync(0xffffffffu, bl, o); const int oil = __shfl_xor_sync(0xffffffffu, il, o);
            if (ol > bl || (ol == bl && oil < il)) { bl = ol; il = oil; }
        }
    }
    if (lane == 0) {
        const float lx = (-lse - jac) / (float)D;
        p.Lx[b] = lx;
        p.lse[b] = lse;
        float pr[5] = {lx, sl, 0.f, 0.f, 0.f};
        if (p.y) {
            const float c = (0.5f * lc - jac) / (float)D;
            const float ok = (iy == il) ? 1.f : 0.f;
            p.Lc[b] = c; p.Ly[b] = ly; p.correct[b] = ok;
            pr[2] = c; pr[3] = ly; pr[4] = ok;
        }
        if (wsum) {
#pragma unroll
            for (int q = 0; q < 5; ++q) wsum[q] += (double)pr[q];
        } else {
            float* dst = p.partials + (size_t)b * 5;
#pragma unroll
            for (int q = 0; q < 5; ++q) dst[q] = pr[q];
        }
    }
}

// Executed by the last CTA: the five batch means.
template <int NT>
__device__ __forceinline__ void head_batch_means(const HeadP& p) {
    const int B = p.B, C = p.C, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // batch means (fp64, fixed order): thread-strided partial sums of the five per-sample records, a shuffle tree per warp, then
    // warp q adds the eight warp sums of quantity q.  (A single warp per quantity was a 25 us serial tail at B = 8192.)
    {
        __shared__ double s_red[5][NT / 32];
        double t[5] = {0., 0., 0., 0., 0.};
        constexpr int UB = 4;
        for (int b1 = tid; b1 < B; b1 += NT * UB) {
            float v[UB][5];
#pragma unroll
            for (int u = 0; u < UB; ++u)
#pragma unroll
                for (int q = 0; q < 5; ++q) v[u][q] = (b1 + u * NT < B) ? __ldcg(p.partials + (size_t)(b1 + u * NT) * 5 + q) : 0.f;
#pragma unroll
            for (int u = 0; u < UB; ++u)
#pragma unroll
                for (int q = 0; q < 5; ++q) t[q] += (double)v[u][q];
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const double w = warp_sum_d(t[q]);
            if (lane == 0) s_red[q][wid] = w;
        }
        __syncthreads();
        if (tid < 5) {
            double tot = 0.;
            for (int i = 0; i < NT / 32; ++i) tot += s_red[tid][i];
            p.means[tid] = (float)(tot / (tid == 1 ? (double)B * C : (double)B));
        }
    }
}


}  // namespace
