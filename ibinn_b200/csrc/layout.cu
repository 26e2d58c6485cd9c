// Data-movement nodes of the IB-INN graph (HBM-bound, no parameters to learn):
//   * DCTPooling2d        reference dct_transform.py:64-104 (separable NxN transform + channel-fastest flatten)
//   * HaarDownsampling    FrEIA v0.2, used at reference inn_architecture.py:127
//   * PermuteRandom       FrEIA v0.2, used at reference inn_architecture.py:160
#include "common.cuh"

namespace {

// out[b, (kw*N + kh)*C + c] = scale * sum_{h,w} M[kh][h] M[kw][w] x[b,c,h,w]
__global__ void __launch_bounds__(IBINN_THREADS) dct_fwd_kernel(const float* __restrict__ x, const float* __restrict__ M,
                                                                float* __restrict__ out, int C, int N, float scale) {
    IBINN_DYN_SMEM(float, smem);
    // NNp: odd per-channel pitch of the intermediate.  The second pass walks channels fastest (the output is channel-fastest), and
    // with the natural pitch N*N = 64 all 32 lanes of a warp hit one bank (measured: 3.8 % of the HBM peak at batch 8192).
    const int D = C * N * N, NN = N * N, NNp = NN | 1;
    float* s_x = smem;
    float* s_t = s_x + D;
    float* s_M = s_t + C * NNp;
    const int tid = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * D;
    if ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) for (int e = tid * 4; e < D; e += IBINN_THREADS * 4) st4(s_x + e, __ldg(reinterpret_cast<const float4*>(x + base + e)));
    else for (int e = tid; e < D; e += IBINN_THREADS) s_x[e] = __ldg(x + base + e);
    for (int e = tid; e < NN; e += IBINN_THREADS) s_M[e] = __ldg(M + e);
    __syncthreads();
    // t[c][h][kw] = sum_w M[kw][w] x[c][h][w]
    for (int e = tid; e < D; e += IBINN_THREADS) {
        const int kw = e % N, ch = e / N;
        float acc = 0.f;
        for (int w = 0; w < N; ++w) acc = fmaf(s_M[kw * N + w], s_x[ch * N + w], acc);
        const int c = e / NN;
        s_t[c * NNp + (e - c * NN)] = acc;
    }
    __syncthreads();
    for (int e = tid; e < D; e += IBINN_THREADS) {
        const int c = e % C, kk = e / C;
        const int kh = kk % N, kw = kk / N;
        float acc = 0.f;
        for (int h = 0; h < N; ++h) acc = fmaf(s_M[kh * N + h], s_t[c * NNp + h * N + kw], acc);
        out[base + e] = acc * scale;
    }
}

// N = 8 (every CIFAR configuration: the last stage is 8 x 8) with register blocking: a thread transforms a whole 8-vector, so the
// operands cost ~3 shared-memory reads per output instead of 16, the index arithmetic is shifts, the input arrives by 128-bit
// loads and the channel-fastest output is written as full 128-byte lines.  Same summation order as the generic kernel (bit-equal).
__global__ void __launch_bounds__(IBINN_THREADS, 5) dct8_fwd_kernel(const float* __restrict__ x, const float* __restrict__ M,
                                                                 float* __restrict__ out, int C, float scale) {
    IBINN_DYN_SMEM(float, smem);
    const int D = C * 64;
    float* s_x = smem;               // [c][h][w]
    float* s_t = s_x + D;            // [c][h][kw], channel pitch 65
    float* s_M = smem + ((D + C * 65 + 3) & ~3);      // 16-byte aligned for the 128-bit broadcast reads
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t base = (size_t)blockIdx.x * D;
    for (int e = tid * 4; e < D; e += IBINN_THREADS * 4) st4(s_x + e, __ldg(reinterpret_cast<const float4*>(x + base + e)));
    if (tid < 64) s_M[tid] = __ldg(M + tid);
    __syncthreads();
    // t[c][h][kw] = sum_w M[kw][w] x[c][h][w]: one (c,h) row per thread
    for (int r = tid; r < C * 8; r += IBINN_THREADS) {
        const float4 x0 = ld4(s_x + r * 8), x1 = ld4(s_x + r * 8 + 4);
        const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        float* tr_ = s_t + (r >> 3) * 65 + (r & 7) * 8;
#pragma unroll 2
        for (int kw = 0; kw < 8; ++kw) {          // (not fully unrolled: the compiler would keep all 64 coefficients in registers)
            const float4 m0 = ld4(s_M + kw * 8), m1 = ld4(s_M + kw * 8 + 4);
            const float mv[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) acc = fmaf(mv[w], xv[w], acc);
            tr_[kw] = acc;
        }
    }
    __syncthreads();
    // out[(kw*8 + kh)*C + c] = scale * sum_h M[kh][h] t[c][h][kw]: warp <-> kw, lane <-> c
    for (int c = lane; c < C; c += 32) {
        float tv[8];
#pragma unroll
        for (int h = 0; h < 8; ++h) tv[h] = s_t[c * 65 + h * 8 + wid];
#pragma unroll 2
        for (int kh = 0; kh < 8; ++kh) {
            const float4 m0 = ld4(s_M + kh * 8), m1 = ld4(s_M + kh * 8 + 4);
            const float mv[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
            float acc = 0.f;
#pragma unroll
            for (int h = 0; h < 8; ++h) acc = fmaf(mv[h], tv[h], acc);
            out[base + (size_t)(wid * 8 + kh) * C + c] = acc * scale;
        }
    }
}

// x[b,c,h,w] = scale * sum_{kh,kw} A(h,kh) A(w,kw) t[b, (kw*N + kh)*C + c],  A(r,k) = tr ? M[k][r] : M[r][k]
__global__ void __launch_bounds__(IBINN_THREADS) dct_inv_kernel(const float* __restrict__ t, const float* __restrict__ M,
                                                                float* __restrict__ x, int C, int N, float scale, int tr) {
    IBINN_DYN_SMEM(float, smem);
    const int D = C * N * N, NN = N * N, NNp = NN | 1;      // odd channel pitch: the channel-fastest scatter below is conflict-free
    float* s_t = smem;            // [c][kh][kw], channel pitch NNp
    float* s_u = s_t + C * NNp;   // [c][kh][w]
    float* s_A = s_u + D;         // A(r,k) row-major
    const int tid = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * D;
    for (int e = tid; e < D; e += IBINN_THREADS) {
        const int c = e % C, kk = e / C;
        const int kh = kk % N, kw = kk / N;
        s_t[c * NNp + kh * N + kw] = __ldg(t + base + e);
    }
    for (int e = tid; e < NN; e += IBINN_THREADS) {
        const int r = e / N, k = e % N;
        s_A[e] = tr ? __ldg(M + k * N + r) : __ldg(M + e);
    }
    __syncthreads();
    for (int e = tid; e < D; e += IBINN_THREADS) {
        const int w = e % N, ck = e / N, c = ck / N;
        const float* tr_ = s_t + c * NNp + (ck - c * N) * N;
        float acc = 0.f;
        for (int kw = 0; kw < N; ++kw) acc = fmaf(s_A[w * N + kw], tr_[kw], acc);
        s_u[e] = acc;
    }
    __syncthreads();
    for (int e = tid; e < D; e += IBINN_THREADS) {
        const int w = e % N, h = (e / N) % N, c = e / NN;
        float acc = 0.f;
        for (int kh = 0; kh < N; ++kh) acc = fmaf(s_A[h * N + kh], s_u[(c * N + kh) * N + w], acc);
        x[base + e] = acc * scale;
    }
}

// MODE 0: y[b,4c+j,h,w] = fac * sum_{dy,dx} sgn_j(dy,dx) x[b,c,2h+dy,2w+dx]
// MODE 1: x[b,c,2h+dy,2w+dx] = fac * sum_j sgn_j(dy,dx) y[b,4c+j,h,w]      (transpose == inverse up to fac)
template <int MODE>
__global__ void __launch_bounds__(IBINN_THREADS) haar_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                             long long nblk, int C, int H, int W, float fac) {
    const int Ho = H / 2, Wo = W / 2;
    for (long long e = (long long)blockIdx.x * IBINN_THREADS + threadIdx.x; e < nblk; e += (long long)gridDim.x * IBINN_THREADS) {
        const int w = (int)(e % Wo);
        const int h = (int)((e / Wo) % Ho);
        const long long bc = e / ((long long)Wo * Ho);   // b*C + c
        const size_t hi = ((size_t)bc * H + 2 * h) * W + 2 * w;                 // full-res 2x2 block origin
        const size_t lo = ((size_t)bc * 4 * Ho + h) * Wo + w;                   // channel 4c+0 at (h,w)
        const size_t cs = (size_t)Ho * Wo;
        float v0, v1, v2, v3;
        if (MODE == 0) { v0 = __ldg(in + hi); v1 = __ldg(in + hi + 1); v2 = __ldg(in + hi + W); v3 = __ldg(in + hi + W + 1); }
        else { v0 = __ldg(in + lo); v1 = __ldg(in + lo + cs); v2 = __ldg(in + lo + 2 * cs); v3 = __ldg(in + lo + 3 * cs); }
        // the 4x4 sign matrix is symmetric: row j / column (dy,dx): [++++], [+-+-], [++--], [+--+]
        const float r0 = (v0 + v1 + v2 + v3) * fac;
        const float r1 = (v0 - v1 + v2 - v3) * fac;
        const float r2 = (v0 + v1 - v2 - v3) * fac;
        const float r3 = (v0 - v1 - v2 + v3) * fac;
        if (MODE == 0) { out[lo] = r0; out[lo + cs] = r1; out[lo + 2 * cs] = r2; out[lo + 3 * cs] = r3; }
        else { out[hi] = r0; out[hi + 1] = r1; out[hi + W] = r2; out[hi + W + 1] = r3; }
    }
}

// One CTA per row when D is a multiple of 4: the row is read with coalesced 128-bit loads into shared memory, gathered there and
// written with 128-bit stores (the direct gather below reads 32 scattered sectors per warp instruction).
__global__ void __launch_bounds__(IBINN_THREADS) gather_rows_kernel(const float* __restrict__ x, const int* __restrict__ idx,
                                                                    float* __restrict__ out, int B, int D) {
    IBINN_DYN_SMEM(float, s_row);
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const float* xr = x + (size_t)b * D;
        float* orow = out + (size_t)b * D;
        for (int e = threadIdx.x * 4; e < D; e += IBINN_THREADS * 4) st4(s_row + e, __ldg(reinterpret_cast<const float4*>(xr + e)));
        __syncthreads();
        for (int e = threadIdx.x * 4; e < D; e += IBINN_THREADS * 4) {
            const int4 i4 = __ldg(reinterpret_cast<const int4*>(idx + e));
            st4(orow + e, make_float4(s_row[i4.x], s_row[i4.y], s_row[i4.z], s_row[i4.w]));
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(IBINN_THREADS) gather_kernel(const float* __restrict__ x, const int* __restrict__ idx,
                                                               float* __restrict__ out, long long n, int D) {
    for (long long e = (long long)blockIdx.x * IBINN_THREADS + threadIdx.x; e < n; e += (long long)gridDim.x * IBINN_THREADS) {
        const long long b = e / D;
        const int i = (int)(e - b * D);
        out[e] = __ldg(x + b * D + __ldg(idx + i));
    }
}

static inline unsigned grid_for(long long n) {
    long long g = (n + IBINN_THREADS - 1) / IBINN_THREADS;
    const long long cap = 148LL * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int ibinn_dct2d_forward(const float* x, const float* M, float* out, int B, int C, int N, float scale,
                                   ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!x || !M || !out) return IBINN_ENULL;
    if (B <= 0 || C <= 0 || N <= 0 || N > 64) return IBINN_EINVAL;
    if (N == 8 && ((uintptr_t)x & 15) == 0) {
        static_assert(IBINN_THREADS == 256, "dct8_fwd_kernel maps its 8 warps to the 8 column frequencies");
        const size_t sm8 = ((size_t)C * 64 + (size_t)C * 65 + 4 + 64) * sizeof(float);
        e = ibinn_set_smem(dct8_fwd_kernel, sm8);
        if (e) return e;
        IBINN_LAUNCH(dct8_fwd_kernel, dim3(B), dim3(IBINN_THREADS), sm8, (cudaStream_t)stream, x, M, out, C, scale);
        return ibinn_launch_status();
    }
    size_t sm = ((size_t)2 * C * N * N + C + 4 + (size_t)N * N) * sizeof(float);
    e = ibinn_set_smem(dct_fwd_kernel, sm);
    if (e) return e;
    IBINN_LAUNCH(dct_fwd_kernel, dim3(B), dim3(IBINN_THREADS), sm, (cudaStream_t)stream, x, M, out, C, N, scale);
    return ibinn_launch_status();
}

extern "C" int ibinn_dct2d_inverse(const float* t, const float* M, float* x, int B, int C, int N, float scale,
                                   int transpose_M, ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!x || !M || !t) return IBINN_ENULL;
    if (B <= 0 || C <= 0 || N <= 0 || N > 64) return IBINN_EINVAL;
    size_t sm = ((size_t)2 * C * N * N + C + 4 + (size_t)N * N) * sizeof(float);
    e = ibinn_set_smem(dct_inv_kernel, sm);
    if (e) return e;
    IBINN_LAUNCH(dct_inv_kernel, dim3(B), dim3(IBINN_THREADS), sm, (cudaStream_t)stream, t, M, x, C, N, scale, transpose_M);
    return ibinn_launch_status();
}

extern "C" int ibinn_haar_forward(const float* x, float* y, int B, int C, int H, int W, float fac, ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!x || !y) return IBINN_ENULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return IBINN_EINVAL;
    const long long n = (long long)B * C * (H / 2) * (W / 2);
    IBINN_LAUNCH(haar_kernel<0>, dim3(grid_for(n)), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, x, y, n, C, H, W, fac);
    return ibinn_launch_status();
}

extern "C" int ibinn_haar_inverse(const float* y, float* x, int B, int C, int H, int W, float fac, ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!x || !y) return IBINN_ENULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return IBINN_EINVAL;
    const long long n = (long long)B * C * (H / 2) * (W / 2);
    IBINN_LAUNCH(haar_kernel<1>, dim3(grid_for(n)), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, y, x, n, C, H, W, fac);
    return ibinn_launch_status();
}

extern "C" int ibinn_permute_features(const float* x, const int32_t* index, float* out, int B, int D, ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!x || !index || !out) return IBINN_ENULL;
    if (B <= 0 || D <= 0) return IBINN_EINVAL;
    const long long n = (long long)B * D;
    if ((D & 3) == 0 && (size_t)D * sizeof(float) <= 160 * 1024 && (((uintptr_t)x | (uintptr_t)out | (uintptr_t)index) & 15) == 0) {
        const size_t sm = (size_t)D * sizeof(float);
        e = ibinn_set_smem(gather_rows_kernel, sm);
        if (e) return e;
        IBINN_LAUNCH(gather_rows_kernel, dim3(B < 148 * 8 ? B : 148 * 8), dim3(IBINN_THREADS), sm, (cudaStream_t)stream, x, (const int*)index, out, B, D);
        return ibinn_launch_status();
    }
    IBINN_LAUNCH(gather_kernel, dim3(grid_for(n)), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, x, (const int*)index, out, n, D);
    return ibinn_launch_status();
}
