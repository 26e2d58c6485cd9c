// Fully connected subnets of the GLOW coupling block (reference inn_architecture.py:112-120:
// Linear -> ReLU -> Dropout -> Linear) and their backward as tcgen05 tensor-core GEMMs with
// fp32-accurate 3xTF32 arithmetic (x = hi + lo; D += hi.hi + hi.lo + lo.hi, fp32 accumulate in TMEM).
// One strided kernel serves y = x W^T + b, dx = g W and dW = g^T x; ReLU / dropout (forward) and
// their masks (backward) are applied while an operand is staged, so the pre-activation is the only
// tensor that is stored.  Small-M problems are split along K to fill the 148 SMs.
// (The plain CUDA-core kernel below is what the host-side simulator build runs; the product build
// never launches it.)
#include "common.cuh"
#include "tc.cuh"
#include <stdlib.h>

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
    float* ws; unsigned* counter;            // split-K: partial results [splits][M][N] + one ticket per output tile (NULL: fp32 atomics)
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


#ifndef IBINN_EMU
// ------------------------------------------------------------------------------------ tcgen05 path
// CTA tile: 128 (M) x 64 (N), K staged 32 at a time (4 MMAs of K = 8 per pass), 2 stages.
// All 256 threads stage operands (global -> registers -> pre-op -> hi/lo split -> shared, canonical
// K-major no-swizzle layout); thread 0 issues the MMAs of a stage and commits them to that stage's
// mbarrier, so the staging of stage s^1 overlaps the tensor-core work of stage s.
constexpr int TC_M = 128, TC_N = 64, TC_BK = 32, TC_KC = TC_BK / 4;
constexpr int TC_LBO_A = (TC_M + 1) * 16, TC_LBO_B = (TC_N + 1) * 16;          // +1 row: conflict-free staging stores
constexpr int TC_A_BYTES = TC_KC * TC_LBO_A, TC_B_BYTES = TC_KC * TC_LBO_B;
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;                 // hi + lo of both operands
constexpr int TC_SMEM = 2 * TC_STAGE_BYTES + 64;

// One operand tile (rows [r0, r0+ROWS) x k [k0, k0+32)) travels global -> registers (tc_fetch) -> pre-op, hi/lo split ->
// shared memory (tc_commit).  The two halves are separate so that the fetch of K block i+1 is in flight while block i is
// split, stored and multiplied: the FC GEMMs have 3-12 K blocks per CTA and would otherwise expose one HBM latency each.
template <int ROWS>
struct TcRegs {
    static constexpr int NT = ROWS * TC_KC / IBINN_THREADS;       // chunk tasks per thread (4 for A, 2 for B)
    static_assert(ROWS * TC_KC % IBINN_THREADS == 0, "tile / thread mismatch");
    float v[NT][4], w2[NT][4], mk[NT][4];
};

template <int ROWS, bool IS_A>
__device__ __forceinline__ void tc_fetch(const GemmP& p, TcRegs<ROWS>& R, int r0, int k0, int kend) {
    const float* base = IS_A ? p.A : p.Bm;
    const long long sr = IS_A ? p.sAi : p.sBj, sk = IS_A ? p.sAk : p.sBk;
    const int nrows = IS_A ? p.M : p.N;
    const int pre = IS_A ? p.a_pre : p.b_pre;
    const float* x2 = IS_A ? p.A2 : nullptr;
    const float* mask = IS_A ? p.Amask : p.Bmask;
    const bool kcontig = (sk == 1);
    // 16-byte loads need k contiguous, aligned rows and a full K block
    const bool vec = kcontig && (sr % 4 == 0) && (k0 % 4 == 0) && (k0 + TC_BK <= kend) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0) &&
                     (pre != 2 || (reinterpret_cast<uintptr_t>(x2) & 15) == 0) && (!mask || (reinterpret_cast<uintptr_t>(mask) & 15) == 0);
#pragma unroll
    for (int t = 0; t < TcRegs<ROWS>::NT; ++t) {
        const int c = threadIdx.x + t * IBINN_THREADS;
        int r, k4;
        if (kcontig) { k4 = c % TC_KC; r = c / TC_KC; } else { r = c % ROWS; k4 = c / ROWS; }
        const int row = r0 + r, k = k0 + k4 * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) { R.v[t][i] = 0.f; R.w2[t][i] = 1.f; R.mk[t][i] = 1.f; }
        if (row >= nrows) continue;
        if (vec) {
            const size_t off = (size_t)row * sr + k;
            const float4 a4 = *reinterpret_cast<const float4*>(base + off);
            R.v[t][0] = a4.x; R.v[t][1] = a4.y; R.v[t][2] = a4.z; R.v[t][3] = a4.w;
            if (pre == 2) { const float4 b4 = *reinterpret_cast<const float4*>(x2 + off); R.w2[t][0] = b4.x; R.w2[t][1] = b4.y; R.w2[t][2] = b4.z; R.w2[t][3] = b4.w; }
            if (pre != 0 && mask) { const float4 m4 = *reinterpret_cast<const float4*>(mask + off); R.mk[t][0] = m4.x; R.mk[t][1] = m4.y; R.mk[t][2] = m4.z; R.mk[t][3] = m4.w; }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (k + i < kend) {
                    const size_t off = (size_t)row * sr + (size_t)(k + i) * sk;
                    R.v[t][i] = __ldg(base + off);
                    if (pre == 2) R.w2[t][i] = __ldg(x2 + off);
                    if (pre != 0 && mask) R.mk[t][i] = __ldg(mask + off);
                }
            }
        }
    }
}

template <int ROWS, int LBO, bool IS_A>
__device__ __forceinline__ void tc_commit(const GemmP& p, const TcRegs<ROWS>& R, unsigned char* hi, unsigned char* lo) {
    const long long sk = IS_A ? p.sAk : p.sBk;
    const int pre = IS_A ? p.a_pre : p.b_pre;
    const bool kcontig = (sk == 1);
#pragma unroll
    for (int t = 0; t < TcRegs<ROWS>::NT; ++t) {
        const int c = threadIdx.x + t * IBINN_THREADS;
        int r, k4;
        if (kcontig) { k4 = c % TC_KC; r = c / TC_KC; } else { r = c % ROWS; k4 = c / ROWS; }
        float h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float tv = R.v[t][i];
            if (pre == 1) tv = fmaxf(tv, 0.f) * R.mk[t][i];
            else if (pre == 2) tv = (R.w2[t][i] > 0.f ? tv : 0.f) * R.mk[t][i];
            tc::split_tf32(tv, h[i], l[i]);
        }
        *reinterpret_cast<float4*>(hi + k4 * LBO + r * 16) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(lo + k4 * LBO + r * 16) = make_float4(l[0], l[1], l[2], l[3]);
    }
}

__global__ void __launch_bounds__(IBINN_THREADS) gemm_tc_kernel(const GemmP p) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t bar_stage[2];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * TC_M, n0 = blockIdx.x * TC_N;
    const int kbeg = blockIdx.z * p.kchunk, kend = min(p.K, kbeg + p.kchunk);

    TcRegs<TC_M> ra;
    TcRegs<TC_N> rb;
    tc_fetch<TC_M, true>(p, ra, m0, kbeg, kend);          // first K block: in flight during the set-up below
    tc_fetch<TC_N, false>(p, rb, n0, kbeg, kend);

    if (tid == 0) {
        tc::mbar_init(&bar_stage[0], 1);
        tc::mbar_init(&bar_stage[1], 1);
        tc::mbar_init(&bar_done, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, TC_N);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t idesc = tc::idesc_tf32(TC_M, TC_N);

    int it = 0;
    for (int k0 = kbeg; k0 < kend; k0 += TC_BK, ++it) {
        const int s = it & 1;
        unsigned char* st = tc_smem + s * TC_STAGE_BYTES;
        unsigned char *a_hi = st, *a_lo = st + TC_A_BYTES, *b_hi = st + 2 * TC_A_BYTES, *b_lo = b_hi + TC_B_BYTES;
        if (it >= 2) tc::mbar_wait(&bar_stage[s], ((it >> 1) - 1) & 1);      // MMAs that read this stage are done
        tc_commit<TC_M, TC_LBO_A, true>(p, ra, a_hi, a_lo);
        tc_commit<TC_N, TC_LBO_B, false>(p, rb, b_hi, b_lo);
        if (k0 + TC_BK < kend) {                                             // next K block: loads overlap this block's MMAs
            tc_fetch<TC_M, true>(p, ra, m0, k0 + TC_BK, kend);
            tc_fetch<TC_N, false>(p, rb, n0, k0 + TC_BK, kend);
        }
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            const uint32_t ah = tc::smem_u32(a_hi), al = tc::smem_u32(a_lo), bh = tc::smem_u32(b_hi), bl = tc::smem_u32(b_lo);
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t aa = pass == 2 ? al : ah, bb = pass == 1 ? bl : bh;
#pragma unroll
                for (int ks = 0; ks < TC_BK / 8; ++ks) {
                    const uint64_t da = tc::smem_desc(aa + ks * 2 * TC_LBO_A, TC_LBO_A);
                    const uint64_t db = tc::smem_desc(bb + ks * 2 * TC_LBO_B, TC_LBO_B);
                    tc::mma_tf32(tmem_d, da, db, idesc, (it | pass | ks) != 0);
                }
            }
            tc::mma_commit(&bar_stage[s]);
        }
    }
    if (tid == 0) tc::mma_commit(&bar_done);
    tc::mbar_wait(&bar_done, 0);
    tc::fence_after_sync();

    if (warp < 4) {
        // TMEM lane = output row, so a thread owns a row: storing from here would scatter every instruction over 32 rows.
        // Each warp transposes its 32 x 64 block through the (idle) stage buffers and writes rows as 128-byte segments.
        float* s_o = reinterpret_cast<float*>(tc_smem) + warp * 32 * (TC_N + 1);
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c8 = 0; c8 < TC_N / 8; ++c8) {
            float v[8];
            tc::tmem_ld8(taddr + c8 * 8, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) s_o[lane * (TC_N + 1) + c8 * 8 + i] = v[i];
        }
        __syncwarp();
        float bj[TC_N / 32];
#pragma unroll
        for (int h = 0; h < TC_N / 32; ++h) {
            const int j = n0 + h * 32 + lane;
            bj[h] = (p.bias && blockIdx.z == 0 && j < p.N) ? __ldg(p.bias + j) : 0.f;
        }
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
            const int row = m0 + warp * 32 + r;
            if (row >= p.M) break;
#pragma unroll
            for (int h = 0; h < TC_N / 32; ++h) {
                const int j = n0 + h * 32 + lane;
                if (j >= p.N) continue;
                const float o = s_o[r * (TC_N + 1) + h * 32 + lane] + bj[h];
                float* dst = p.C + (size_t)row * p.ldc + j;
                if (p.splits > 1) {
                    if (p.ws) p.ws[((size_t)blockIdx.z * p.M + row) * p.N + j] = o;
                    else atomicAdd(dst, o);
                } else if (p.accumulate) *dst += o;
                else *dst = o;
            }
        }
    }
    if (p.splits > 1 && p.ws) {
        // Deterministic split-K: the last CTA of an output tile to arrive adds the partial tiles in split order (fixed order:
        // the step is bit-reproducible; no fp32 atomics, no memset of C).
        const unsigned tile = blockIdx.y * gridDim.x + blockIdx.x;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(p.counter + tile, 1u) == (unsigned)p.splits - 1u) ? 1 : 0;
        __syncthreads();
        if (s_last) {
            __threadfence();
            // all 256 threads, four consecutive columns each, up to 16 partial values in flight per thread (the sum itself runs in
            // split order): the tail costs two or three L2 round trips instead of one per split and row
            const int rows = min(TC_M, p.M - m0);
            const bool v4 = ((p.N & 3) == 0) && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (n0 + TC_N <= p.N);
            if (v4) {
                constexpr int UB = 8;
                for (int e = tid; e < rows * (TC_N / 4); e += 2 * IBINN_THREADS) {
                    const int e2 = e + IBINN_THREADS;
                    const bool has2 = e2 < rows * (TC_N / 4);
                    const int r1 = e / (TC_N / 4), c1 = (e % (TC_N / 4)) * 4, r2 = e2 / (TC_N / 4), c2 = (e2 % (TC_N / 4)) * 4;
                    const float* s1 = p.ws + (size_t)(m0 + r1) * p.N + n0 + c1;
                    const float* s2 = p.ws + (size_t)(m0 + r2) * p.N + n0 + c2;
                    const size_t zs = (size_t)p.M * p.N;
                    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
                    for (int z0 = 0; z0 < p.splits; z0 += UB) {
                        float4 v1[UB], v2[UB];
#pragma unroll
                        for (int u = 0; u < UB; ++u) {
                            const bool in = z0 + u < p.splits;
                            v1[u] = in ? __ldcg(reinterpret_cast<const float4*>(s1 + (size_t)(z0 + u) * zs)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            v2[u] = (in && has2) ? __ldcg(reinterpret_cast<const float4*>(s2 + (size_t)(z0 + u) * zs)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int u = 0; u < UB; ++u) {
                            a1.x += v1[u].x; a1.y += v1[u].y; a1.z += v1[u].z; a1.w += v1[u].w;
                            a2.x += v2[u].x; a2.y += v2[u].y; a2.z += v2[u].z; a2.w += v2[u].w;
                        }
                    }
                    float4* d1 = reinterpret_cast<float4*>(p.C + (size_t)(m0 + r1) * p.ldc + n0 + c1);
                    if (p.accumulate) { const float4 o = *d1; a1.x += o.x; a1.y += o.y; a1.z += o.z; a1.w += o.w; }
                    *d1 = a1;
                    if (has2) {
                        float4* d2 = reinterpret_cast<float4*>(p.C + (size_t)(m0 + r2) * p.ldc + n0 + c2);
                        if (p.accumulate) { const float4 o = *d2; a2.x += o.x; a2.y += o.y; a2.z += o.z; a2.w += o.w; }
                        *d2 = a2;
                    }
                }
            } else {
                for (int e = tid; e < rows * TC_N; e += IBINN_THREADS) {
                    const int r = e / TC_N, j = n0 + e % TC_N;
                    if (j >= p.N) continue;
                    float acc = 0.f;
                    for (int z = 0; z < p.splits; ++z) acc += __ldcg(p.ws + ((size_t)z * p.M + m0 + r) * p.N + j);
                    float* dst = p.C + (size_t)(m0 + r) * p.ldc + j;
                    if (p.accumulate) *dst += acc;
                    else *dst = acc;
                }
            }
            if (tid == 0) p.counter[tile] = 0u;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, TC_N);
}
#endif  // !IBINN_EMU

// out[j] (+)= sum_i pre(A)[i][j]   -- bias gradient.  Block = 32 columns x 8 row slices; every slice keeps 8 independent
// loads in flight and the slices are combined in fixed order (deterministic).
__global__ void __launch_bounds__(IBINN_THREADS) colsum_kernel(const float* __restrict__ A, const float* __restrict__ A2,
                                                               const float* __restrict__ mask, int a_pre, float* __restrict__ out,
                                                               int M, int N) {
    __shared__ float s_part[8][33];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (j < N) {
        for (int i0 = slice; i0 < M; i0 += 8 * 8) {
            float v[8], h[8], mk[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + 8 * u;
                v[u] = 0.f; h[u] = 1.f; mk[u] = 1.f;
                if (i < M) {
                    const size_t off = (size_t)i * N + j;
                    v[u] = __ldg(A + off);
                    if (a_pre == 2) { h[u] = __ldg(A2 + off); if (mask) mk[u] = __ldg(mask + off); }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) s += (a_pre == 2) ? (h[u] > 0.f ? v[u] : 0.f) * mk[u] : v[u];
        }
    }
    s_part[slice][lane] = s;
    __syncthreads();
    if (slice == 0 && j < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += s_part[k][lane];
        out[j] += t;
    }
}

// grid and K split of a GEMM with an (M, N) result and contraction length K
static void gemm_plan(int M, int N, int K, int& gx, int& gy, int& splits, int& kchunk) {
#ifndef IBINN_EMU
    const int tile_m = TC_M, tile_n = TC_N, tile_k = TC_BK;
#else
    const int tile_m = BM, tile_n = BN, tile_k = BK;
#endif
    gx = (N + tile_n - 1) / tile_n;
    gy = (M + tile_m - 1) / tile_m;
    splits = 1;
    {
        const int ctas = gx * gy;
        // CTAs to aim for.  96 (4 splits for the 128 x 1536 x 1536 GEMMs of the default model) beats 148 (8 splits) once the partial
        // tiles are summed in order by the last CTA of a tile: 9.00 vs 9.04 ms per step (2 splits: 9.16, no split: 9.35).
        static const int cap = [] { const char* e = getenv("IBINN_GEMM_SPLIT_TARGET"); return e ? atoi(e) : 96; }();
        while (ctas * splits < cap && splits < 16 && K / (splits * 2) >= 4 * tile_k) splits *= 2;
    }
    kchunk = (K + splits - 1) / splits;
    kchunk = (kchunk + tile_k - 1) / tile_k * tile_k;
    splits = (K + kchunk - 1) / kchunk;
}

static int gemm_run(GemmP p, cudaStream_t st) {
    int gx, gy, splits, kchunk;
    gemm_plan(p.M, p.N, p.K, gx, gy, splits, kchunk);
    p.kchunk = kchunk;
    p.splits = splits;
#ifdef IBINN_EMU
    p.ws = nullptr;           // (the simulator's CUDA-core kernel keeps the atomics: its threads run one after the other anyway)
#endif
    if (p.ws && !p.counter) return IBINN_ENULL;
    if (splits > 1 && !p.accumulate && !p.ws) {
        // contiguous rows only (ldc == N) for a single memset; otherwise row by row is not needed here
        if (p.ldc != p.N) return IBINN_EINVAL;
        cudaError_t e = cudaMemsetAsync(p.C, 0, (size_t)p.M * p.N * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
#ifndef IBINN_EMU
    int e = ibinn_set_smem(gemm_tc_kernel, TC_SMEM);
    if (e) return e;
    IBINN_LAUNCH(gemm_tc_kernel, dim3(gx, gy, splits), dim3(IBINN_THREADS), TC_SMEM, st, p);
#else
    IBINN_LAUNCH(gemm_kernel, dim3(gx, gy, splits), dim3(IBINN_THREADS), 0, st, p);
#endif
    return ibinn_launch_status();
}

}  // namespace

// y(B,N) = pre(x)(B,K; row stride x_rs) . w(N,K)^T + bias         pre: 0 none, 1 relu(x)*mask
// (a mask is only meaningful for a contiguous x: it is indexed with the same offsets)
// workspace / tickets of the three GEMMs below (result (M, N), contraction length Kc): forward (B, N, K), dgrad (B, K, N),
// wgrad (N, K, B).  0 floats = no split along K for this shape.
extern "C" int64_t ibinn_linear_workspace_floats(int M, int N, int Kc) {
    if (M <= 0 || N <= 0 || Kc <= 0) return IBINN_EINVAL;
    int gx, gy, splits, kchunk;
    gemm_plan(M, N, Kc, gx, gy, splits, kchunk);
    return splits > 1 ? (int64_t)splits * M * N : 0;
}
extern "C" int ibinn_linear_tickets(int M, int N, int Kc) {
    if (M <= 0 || N <= 0 || Kc <= 0) return IBINN_EINVAL;
    int gx, gy, splits, kchunk;
    gemm_plan(M, N, Kc, gx, gy, splits, kchunk);
    return gx * gy;
}

extern "C" int ibinn_linear_forward(const float* x, int64_t x_rs, const float* x_mask, int x_pre, const float* w,
                                    const float* bias, float* y, int B, int K, int N, float* workspace, uint32_t* tickets,
                                    ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!x || !w || !y) return IBINN_ENULL;
    if (B <= 0 || K <= 0 || N <= 0 || (x_pre != 0 && x_pre != 1) || x_rs < K) return IBINN_EINVAL;
    GemmP p{};
    p.A = x; p.sAi = x_rs; p.sAk = 1; p.Amask = x_mask; p.a_pre = x_pre;
    p.Bm = w; p.sBk = 1; p.sBj = K; p.bias = bias; p.C = y; p.ldc = N; p.M = B; p.N = N; p.K = K;
    p.ws = workspace; p.counter = tickets;
    return gemm_run(p, (cudaStream_t)stream);
}

// dx(B,K; row stride dx_rs) (+)= pre(g)(B,N) . w(N,K)            pre: 0 none, 2 g*[h>0]*mask
extern "C" int ibinn_linear_dgrad(const float* g, const float* h, const float* g_mask, int g_pre, const float* w, float* dx,
                                  int64_t dx_rs, int accumulate, int B, int K, int N, float* workspace, uint32_t* tickets,
                                  ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!g || !w || !dx || (g_pre == 2 && !h)) return IBINN_ENULL;
    if (B <= 0 || K <= 0 || N <= 0 || (g_pre != 0 && g_pre != 2) || dx_rs < K) return IBINN_EINVAL;
    GemmP p{};
    p.A = g; p.sAi = N; p.sAk = 1; p.A2 = h; p.Amask = g_mask; p.a_pre = g_pre;
    p.Bm = w; p.sBk = K; p.sBj = 1; p.C = dx; p.ldc = dx_rs; p.M = B; p.N = K; p.K = N; p.accumulate = accumulate;
    p.ws = workspace; p.counter = tickets;
    return gemm_run(p, (cudaStream_t)stream);
}

// dw(N,K) += pre(g)(B,N)^T . pre(x)(B,K; row stride x_rs) ;  db(N) += colsum(pre(g))
//   g_pre: 0 none, 2 g*[h>0]*g_mask ;  x_pre: 0 none, 1 relu(x)*x_mask
extern "C" int ibinn_linear_wgrad(const float* g, const float* h, const float* g_mask, int g_pre, const float* x,
                                  int64_t x_rs, const float* x_mask, int x_pre, float* dw, float* db, int B, int K, int N,
                                  float* workspace, uint32_t* tickets, ibinn_stream_t stream) {
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
    p.ws = workspace; p.counter = tickets;
    e = gemm_run(p, st);
    if (e) return e;
    if (db) {
        IBINN_LAUNCH(colsum_kernel, dim3((N + 31) / 32), dim3(IBINN_THREADS), 0, st, g, h, g_mask,
                     g_pre, db, B, N);
        return ibinn_launch_status();
    }
    return 0;
}
