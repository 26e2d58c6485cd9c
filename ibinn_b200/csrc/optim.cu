// Gradient clipping + SGD on flat parameter buffers (reference train.py:109-110
// `clip_grad_norm_(params, 8.)` + `optimizer.step()`, model.py:97-100 torch.optim.SGD):
// two launches per step instead of ~2400.  Hyper-parameters live in device memory so that a
// captured CUDA graph follows the learning-rate schedule without re-capture.
#include "common.cuh"

namespace {

// partial sums of squares in double, fixed grid -> deterministic; the last CTA writes
// out[0] = ||g||_2 and out[1] = min(1, max_norm / (||g|| + 1e-6))  (torch.nn.utils.clip_grad_norm_)
__global__ void __launch_bounds__(IBINN_THREADS) sumsq_kernel(const float* __restrict__ g, long long n, const float* extra_sumsq,
                                                              int n_extra, float max_norm, double* partials, unsigned* counter,
                                                              float* out) {
    __shared__ double s_w[IBINN_THREADS / 32];
    __shared__ int s_flag;
    const int tid = threadIdx.x;
    double acc = 0.;
    const long long n4 = n / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    {
        constexpr int U = 4;                              // 64 bytes in flight per thread
        const long long stride = (long long)gridDim.x * IBINN_THREADS;
        for (long long i = (long long)blockIdx.x * IBINN_THREADS + tid; i < n4; i += stride * U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = (i + u * stride < n4) ? __ldg(g4 + i + u * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < U; ++u) acc += (double)(v[u].x * v[u].x + v[u].y * v[u].y) + (double)(v[u].z * v[u].z + v[u].w * v[u].w);
        }
    }
    if (blockIdx.x == 0)
        for (long long i = n4 * 4 + tid; i < n; i += IBINN_THREADS) acc += (double)g[i] * (double)g[i];
    acc = warp_sum_d(acc);
    if ((tid & 31) == 0) s_w[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.;
        for (int i = 0; i < IBINN_THREADS / 32; ++i) t += s_w[i];
        partials[blockIdx.x] = t;
    }
    if (!last_block_ticket(counter, gridDim.x, &s_flag)) return;
    if (tid < 32) {
        double t = 0.;
        for (unsigned i = tid; i < gridDim.x; i += 32) t += __ldcg(partials + i);       // lane-strided, then a fixed shuffle tree
        t = warp_sum_d(t);
        if (tid != 0) return;
        for (int i = 0; i < n_extra; ++i) t += (double)extra_sumsq[i];
        const double nrm = sqrt(t);
        out[0] = (float)nrm;
        const double c = (double)max_norm / (nrm + 1e-6);
        out[1] = (float)(c < 1. ? c : 1.);
        *counter = 0u;
    }
}

// hyper = {lr, momentum, weight_decay}; coef = clip coefficient (device) or NULL
__global__ void __launch_bounds__(IBINN_THREADS) sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                                            long long n, const float* __restrict__ hyper, const float* __restrict__ coef,
                                                            float grad_scale) {
    const float lr = hyper[0], mom = hyper[1], wd = hyper[2];
    const float c = (coef ? coef[0] : 1.f) * grad_scale;
    for (long long i = (long long)blockIdx.x * IBINN_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * IBINN_THREADS) {
        float d = fmaf(g[i], c, wd * p[i]);
        if (mom != 0.f) {
            d = fmaf(mom, buf[i], d);
            buf[i] = d;
        }
        p[i] -= lr * d;
    }
}

}  // namespace

extern "C" int64_t ibinn_grad_norm_partials_doubles(void) { return 148 * 4; }

extern "C" int ibinn_grad_norm_clip(const float* grad, int64_t n, const float* extra_sumsq, int n_extra, float max_norm,
                                    double* partials, uint32_t* counter, float* out_norm_coef, ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!grad || !partials || !counter || !out_norm_coef) return IBINN_ENULL;
    if (n <= 0) return IBINN_EINVAL;
    if (((uintptr_t)grad & 15) != 0) return IBINN_EALIGN;
    long long blocks = (n / 4 + IBINN_THREADS - 1) / IBINN_THREADS;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 4) blocks = 148 * 4;
    IBINN_LAUNCH(sumsq_kernel, dim3((unsigned)blocks), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, grad, (long long)n, extra_sumsq,
                 n_extra, max_norm, partials, counter, out_norm_coef);
    return ibinn_launch_status();
}

extern "C" int ibinn_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, const float* hyper,
                              const float* clip_coef, float grad_scale, ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!param || !grad || !hyper || !momentum_buf) return IBINN_ENULL;
    if (n <= 0) return IBINN_EINVAL;
    long long blocks = (n + IBINN_THREADS - 1) / IBINN_THREADS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    IBINN_LAUNCH(sgd_kernel, dim3((unsigned)blocks), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, param, grad, momentum_buf,
                 (long long)n, hyper, clip_coef, grad_scale);
    return ibinn_launch_status();
}
