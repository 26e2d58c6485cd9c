// Fully connected subnets of the GLOW coupling block (reference inn_architecture.py:112-120:
// Linear -> ReLU -> Dropout -> Linear) and their backward: exact-fp32 SGEMM on the CUDA cores.
// One strided kernel serves y = x W^T + b, dx = g W and dW = g^T x; ReLU / dropout (forward) and
// their masks (backward) are applied to the A operand while it is staged, so the pre-activation
// is the only tensor that is stored.  Small-M problems are split along K to fill the 148 SMs.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

struct GemmP {
    const float* A; long long sAi, sAk;      // A(i,k) = A[i*sAi + k*sAk]
    const float* A2; const float* Amask;     // pre-op operands, indexed like A
    int a_pre;                               // 0 none, 1 relu(A)*mask, 2 A*[A2>0]*mask
    const float* Bm; long long sBk, sBj;     // B(k,j)
    const float* Bmask; int b_pre;           // 0 none, 1 relu(B)*mask (indexed like B)
    const float* bias;                       // (N) or NULL
    float* C; long long ldc;
    int M, N, K, kchunk, splits, accumulate;
};

__device__ __forceinline__ float gemm_load_a(const GemmP& p, int i, int k) {
    const size_t off = (size_t)i * p.sAi + (size_t)k * p.sAk;
    float v = __ldg(p.A + off);
    if (p.a_pre == 1) {
        v = fmaxf(v, 0.f);
        if (p.Amask) v *= __ldg(p.Amask + off);
    } else if (p.a_pre == 2) {
        v = __ldg(p.A2 + off) > 0.f ? v : 0.f;
        if (p.Amask) v *= __ldg(p.Amask + off);
    }
    return v;
}

__device__ __forceinline__ float gemm_load_b(const GemmP& p, int k, int j) {
    const size_t off = (size_t)k * p.sBk + (size_t)j * p.sBj;
    float v = __ldg(p.Bm + off);
    if (p.b_pre == 1) {
        v = fmaxf(v, 0.f);
        if (p.Bmask) v *= __ldg(p.Bmask + off);
    }
    return v;
}

__global__ void __launch_bounds__(IBINN_THREADS) gemm_kernel(const GemmP p) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * p.kchunk, kend = min(p.K, kbeg + p.kchunk);
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const bool a_kc = (p.sAk == 1), b_jc = (p.sBj == 1);

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        __syncthreads();
        if (a_kc) {
            const int i = tid / 4, kq = (tid % 4) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + kq + j;
                As[kq + j][i] = (i0 + i < p.M && k < kend) ? gemm_load_a(p, i0 + i, k) : 0.f;
            }
        } else {
            const int k = tid / 16, iq = (tid % 16) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + iq + j;
                As[k][iq + j] = (i < p.M && k0 + k < kend) ? gemm_load_a(p, i, k0 + k) : 0.f;
            }
        }
        if (b_jc) {
            const int k = tid / 16, jq = (tid % 16) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int jj = j0 + jq + j;
                Bs[k][jq + j] = (jj < p.N && k0 + k < kend) ? gemm_load_b(p, k0 + k, jj) : 0.f;
            }
        } else {
            const int jl = tid / 4, kq = (tid % 4) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + kq + j;
                Bs[kq + j][jl] = (j0 + jl < p.N && k < kend) ? gemm_load_b(p, k, j0 + jl) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 av = ld4(&As[k][ty * 4]);
            const float4 bv = ld4(&Bs[k][tx * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w};
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(aa[a], bb[b], acc[a][b]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty * 4 + a;
        if (i >= p.M) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = j0 + tx * 4 + b;
            if (j >= p.N) continue;
            float v = acc[a][b];
            if (p.bias && blockIdx.z == 0) v += __ldg(p.bias + j);
            float* dst = p.C + (size_t)i * p.ldc + j;
            if (p.splits > 1) atomicAdd(dst, v);
            else if (p.accumulate) *dst += v;
            else *dst = v;
        }
    }
}

// out[j] (+)= sum_i pre(A)[i][j]   -- bias gradient
__global__ void __launch_bounds__(IBINN_THREADS) colsum_kernel(const float* __restrict__ A, const float* __restrict__ A2,
                                                               const float* __restrict__ mask, int a_pre, float* __restrict__ out,
                                                               int M, int N) {
    const int j = blockIdx.x * IBINN_THREADS + threadIdx.x;
    if (j >= N) return;
    float s = 0.f;
    for (int i = 0; i < M; ++i) {
        const size_t off = (size_t)i * N + j;
        float v = __ldg(A + off);
        if (a_pre == 2) { v = __ldg(A2 + off) > 0.f ? v : 0.f; if (mask) v *= __ldg(mask + off); }
        s += v;
    }
    out[j] += s;
}

static int gemm_run(GemmP p, cudaStream_t st) {
    const int gx = (p.N + BN - 1) / BN, gy = (p.M + BM - 1) / BM;
    int splits = 1;
    {
        const int ctas = gx * gy;
        while (ctas * splits < 148 && splits < 16 && p.K / (splits * 2) >= 4 * BK) splits *= 2;
    }
    int kchunk = (p.K + splits - 1) / splits;
    kchunk = (kchunk + BK - 1) / BK * BK;
    splits = (p.K + kchunk - 1) / kchunk;
    p.kchunk = kchunk;
    p.splits = splits;
    if (splits > 1 && !p.accumulate) {
        // contiguous rows only (ldc == N) for a single memset; otherwise row by row is not needed here
        if (p.ldc != p.N) return IBINN_EINVAL;
        cudaError_t e = cudaMemsetAsync(p.C, 0, (size_t)p.M * p.N * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    IBINN_LAUNCH(gemm_kernel, dim3(gx, gy, splits), dim3(IBINN_THREADS), 0, st, p);
    return ibinn_launch_status();
}

}  // namespace

// y(B,N) = pre(x)(B,K; row stride x_rs) . w(N,K)^T + bias         pre: 0 none, 1 relu(x)*mask
// (a mask is only meaningful for a contiguous x: it is indexed with the same offsets)
extern "C" int ibinn_linear_forward(const float* x, int64_t x_rs, const float* x_mask, int x_pre, const float* w,
                                    const float* bias, float* y, int B, int K, int N, ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!x || !w || !y) return IBINN_ENULL;
    if (B <= 0 || K <= 0 || N <= 0 || (x_pre != 0 && x_pre != 1) || x_rs < K) return IBINN_EINVAL;
    GemmP p{};
    p.A = x; p.sAi = x_rs; p.sAk = 1; p.Amask = x_mask; p.a_pre = x_pre;
    p.Bm = w; p.sBk = 1; p.sBj = K; p.bias = bias; p.C = y; p.ldc = N; p.M = B; p.N = N; p.K = K;
    return gemm_run(p, (cudaStream_t)stream);
}

// dx(B,K; row stride dx_rs) (+)= pre(g)(B,N) . w(N,K)            pre: 0 none, 2 g*[h>0]*mask
extern "C" int ibinn_linear_dgrad(const float* g, const float* h, const float* g_mask, int g_pre, const float* w, float* dx,
                                  int64_t dx_rs, int accumulate, int B, int K, int N, ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!g || !w || !dx || (g_pre == 2 && !h)) return IBINN_ENULL;
    if (B <= 0 || K <= 0 || N <= 0 || (g_pre != 0 && g_pre != 2) || dx_rs < K) return IBINN_EINVAL;
    GemmP p{};
    p.A = g; p.sAi = N; p.sAk = 1; p.A2 = h; p.Amask = g_mask; p.a_pre = g_pre;
    p.Bm = w; p.sBk = K; p.sBj = 1; p.C = dx; p.ldc = dx_rs; p.M = B; p.N = K; p.K = N; p.accumulate = accumulate;
    return gemm_run(p, (cudaStream_t)stream);
}

// dw(N,K) += pre(g)(B,N)^T . pre(x)(B,K; row stride x_rs) ;  db(N) += colsum(pre(g))
//   g_pre: 0 none, 2 g*[h>0]*g_mask ;  x_pre: 0 none, 1 relu(x)*x_mask
extern "C" int ibinn_linear_wgrad(const float* g, const float* h, const float* g_mask, int g_pre, const float* x,
                                  int64_t x_rs, const float* x_mask, int x_pre, float* dw, float* db, int B, int K, int N,
                                  ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!g || !x || !dw || (g_pre == 2 && !h)) return IBINN_ENULL;
    if (B <= 0 || K <= 0 || N <= 0 || (g_pre != 0 && g_pre != 2) || (x_pre != 0 && x_pre != 1) || x_rs < K) return IBINN_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    // C(N,K) = A(N,B) . Bm(B,K) with A(i,k) = g[k][i], Bm(k,j) = x[k][j]
    GemmP p{};
    p.A = g; p.sAi = 1; p.sAk = N; p.A2 = h; p.Amask = g_mask; p.a_pre = g_pre;
    p.Bm = x; p.sBk = x_rs; p.sBj = 1; p.Bmask = x_mask; p.b_pre = x_pre;
    p.C = dw; p.ldc = K; p.M = N; p.N = K; p.K = B; p.accumulate = 1;
    e = gemm_run(p, st);
    if (e) return e;
    if (db) {
        IBINN_LAUNCH(colsum_kernel, dim3((N + IBINN_THREADS - 1) / IBINN_THREADS), dim3(IBINN_THREADS), 0, st, g, h, g_mask,
                     g_pre, db, B, N);
        return ibinn_launch_status();
    }
    return 0;
}
