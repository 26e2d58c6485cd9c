// Class-conditional Gaussian-mixture head and the Information-Bottleneck loss terms
// (reference model.py:113-126 `cluster_distances`, model.py:144-163 loss dict) in one pass over z,
// plus the hand-derived backward.
//
// Distances are evaluated as sum_d (z_d - mu_kd)^2 (not the expanded -2 z.mu + |z|^2 + |mu|^2 the
// reference uses): same value in exact arithmetic, no cancellation in fp32.
#include "common.cuh"
#include <stdlib.h>

#include "head_tail.cuh"

namespace {

// SB samples per CTA: every CTA streams all of mu (C x D floats, L2-resident) once, so at the scoring batch sizes one sample per
// CTA is bound by B x C x D x 4 bytes of L2 traffic (1 GB at B = 8192, C = 10) instead of by reading z once from HBM.  The
// arithmetic per (sample, class) is the same lane-strided sum followed by a warp reduction for every SB, so the results do not
// depend on the batch size.
template <int SB>
__global__ void __launch_bounds__(IBINN_THREADS) head_fwd_kernel(const HeadP p) {
    IBINN_DYN_SMEM(float, smem);
    __shared__ int s_flag;
    const int B = p.B, C = p.C, D = p.D;
    float* s_z = smem;               // [SB][D]
    float* s_zz = s_z + SB * D;      // [SB][C]
    float* s_lw = s_zz + SB * C;     // [C]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, b0 = blockIdx.x * SB;
    const int nb = min(SB, B - b0);
    {
        const float* zb = p.z + (size_t)b0 * D;
        const int n = nb * D;
        if ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(p.z) & 15) == 0) {
            constexpr int ZB = 6;                         // 96 bytes in flight per thread before the first shared-memory store
            const int n4 = n >> 2;
            const float4* zb4 = reinterpret_cast<const float4*>(zb);
            for (int i0 = tid; i0 < n4; i0 += IBINN_THREADS * ZB) {
                float4 v[ZB];
#pragma unroll
                for (int u = 0; u < ZB; ++u) if (i0 + u * IBINN_THREADS < n4) v[u] = __ldg(zb4 + i0 + u * IBINN_THREADS);
#pragma unroll
                for (int u = 0; u < ZB; ++u) if (i0 + u * IBINN_THREADS < n4) st4(s_z + (size_t)(i0 + u * IBINN_THREADS) * 4, v[u]);
            }
        } else
            for (int e = tid; e < n; e += IBINN_THREADS) s_z[e] = __ldg(zb + e);
        for (int e = n + tid; e < SB * D; e += IBINN_THREADS) s_z[e] = 0.f;
    }
    log_softmax_phi(p.phi, C, s_lw);
    __syncthreads();
    // Work item = (class k, part of the feature axis): with few classes (C = 10 on 8 warps) whole-class items leave most warps idle
    // in the last round.  Every item is a lane-strided sum (128-bit loads: a lane owns 4 consecutive features per step) followed by
    // a warp reduction; a class's parts are added in order below.  The split depends on (C, D) only, never on the batch size.
    const bool vec = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(p.mu) & 15) == 0;
    const int nsp = (vec && C < 16 && D % 16 == 0) ? 4 : (vec && C < 32 && D % 8 == 0) ? 2 : 1;
    const int Dp = D / nsp;
    float* s_part = s_lw + C;        // [SB][C][nsp]
    for (int item = wid; item < C * nsp; item += IBINN_THREADS / 32) {
        const int k = item / nsp, part = item - k * nsp;
        const float* mk = p.mu + (size_t)k * D + part * Dp;
        const float* zk = s_z + part * Dp;
        float acc[SB];
#pragma unroll
        for (int s = 0; s < SB; ++s) acc[s] = 0.f;
        if (vec) {
            constexpr int MU = 6;                         // class-mean loads (L2) in flight per lane
            for (int d0 = lane * 4; d0 < Dp; d0 += 128 * MU) {
                float4 mv[MU];
#pragma unroll
                for (int u = 0; u < MU; ++u) if (d0 + 128 * u < Dp) mv[u] = __ldg(reinterpret_cast<const float4*>(mk + d0 + 128 * u));
#pragma unroll
                for (int u = 0; u < MU; ++u) {
                    const int d = d0 + 128 * u;
                    if (d >= Dp) break;
                    const float4 m = mv[u];
#pragma unroll
                    for (int s = 0; s < SB; ++s) {
                        const float4 zv = ld4(zk + s * D + d);
                        float t = zv.x - m.x; acc[s] = fmaf(t, t, acc[s]);
                        t = zv.y - m.y; acc[s] = fmaf(t, t, acc[s]);
                        t = zv.z - m.z; acc[s] = fmaf(t, t, acc[s]);
                        t = zv.w - m.w; acc[s] = fmaf(t, t, acc[s]);
                    }
                }
            }
        } else {
            for (int d = lane; d < Dp; d += 32) {
                const float m = __ldg(mk + d);
#pragma unroll
                for (int s = 0; s < SB; ++s) {
                    const float t = zk[s * D + d] - m;
                    acc[s] = fmaf(t, t, acc[s]);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < SB; ++s) {
            const float v = warp_sum(acc[s]);
            if (lane == 0) s_part[(s * C + k) * nsp + part] = v;
        }
    }
    __syncthreads();
    for (int e = tid; e < SB * C; e += IBINN_THREADS) {
        float v = 0.f;
        for (int q = 0; q < nsp; ++q) v += s_part[e * nsp + q];
        s_zz[e] = v;
    }
    __syncthreads();
    for (int s = wid; s < nb; s += IBINN_THREADS / 32) head_sample_tail(p, b0 + s, s_zz + s * C, s_lw, lane);
    if (!last_block_ticket(p.counter, gridDim.x, &s_flag)) return;
    head_batch_means<IBINN_THREADS>(p);
    if (tid == 0) *p.counter = 0u;
}

// N per-lane partial sums -> N warp totals with 31-ish shuffles per 32 values instead of 5 N: at every butterfly step a lane
// keeps one half of its values and sends the other half to its partner.  Every total is formed by exactly the tree of
// `warp_sum` ((l, l^16), then ^8, ...; a + b == b + a), so the value does not depend on the slot it is computed in.  Afterwards
// the lane holds the totals `base + r` for r < lim (lim <= 2 for N <= 64) in v[r].
template <int N, int O>
struct WarpMultiSum {
    template <int NV>
    static __device__ __forceinline__ void run(float (&v)[NV], int lane, int& base, int& lim) {
        constexpr int H = (N + 1) / 2;
        const bool up = (lane & O) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const float lo = v[i];
            const float hi = (i + H < N) ? v[i + H] : 0.f;
            const float got = __shfl_xor_sync(0xffffffffu, up ? lo : hi, O);
            v[i] = (up ? hi : lo) + got;
        }
        base += up ? H : 0;
        lim = up ? lim - H : min(lim, H);
        WarpMultiSum<H, O / 2>::run(v, lane, base, lim);
    }
};
template <int N>
struct WarpMultiSum<N, 0> {
    template <int NV>
    static __device__ __forceinline__ void run(float (&)[NV], int, int&, int&) {}
};
template <int N>
__device__ __forceinline__ void warp_multi_sum(float (&v)[N], int lane, int& base, int& lim) {
    base = 0;
    lim = N;
    WarpMultiSum<N, 16>::run(v, lane, base, lim);
}

// Ordered loads for the software pipeline of head_fwd_stream_kernel: ptxas gives a load whose first use is in the NEXT loop
// iteration the lowest priority and sinks it to the end of the loop body (seen in SASS: all refills after the last FFMA2), which
// exposes the full HBM latency once per body.  Volatile accesses keep their program order among themselves, so issuing both the
// z refill (global) and the class-mean reads (shared) as volatile pins each refill between two columns' arithmetic.
#ifdef IBINN_EMU
__device__ __forceinline__ float4 ld4_ordered_global(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4_ordered_shared(const float* p) { return *reinterpret_cast<const float4*>(p); }
#else
__device__ __forceinline__ float4 ld4_ordered_global(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld4_ordered_shared(const float* p) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}
#endif

// Scoring-batch forward for few classes (C = CC, C x D floats of class means fit into shared memory): persistent CTAs of NT / 32 warps keep
// mu resident in shared memory and stream z from HBM exactly once; a warp owns SW samples at a time, keeps their SW x CC partial
// distances in registers (one shared-memory read of mu serves SW samples) and never stages z.  The summation order per
// (sample, class) is that of head_fwd_kernel with nsp = 4 (four feature parts, lane-strided 128-bit steps, warp tree, parts added in
// order), so the per-sample results are bit-identical to the small-batch kernel.  The class means sit in shared memory interleaved by
// class PAIRS ([pair][half][D/4][(d, class), 4 floats]) so that one 128-bit read yields two packed (m_2p[d], m_2p+1[d]) operands and the
// distance update runs as FADD2 / FFMA2 (z broadcast to both halves): half the fp32 issue slots of the scalar form (23 M -> 17 M
// warp instructions at B = 8192).  Measured limit (profiles/r01_head_stream.md): 2 B C D fp32 operations is what bounds this kernel,
// not HBM: the packed instructions run on the FMA-heavy pipe only, which is busy 46 % of the active cycles.
template <int SW, int CC, int NT, int UU>
__global__ void __launch_bounds__(NT, 1) head_fwd_stream_kernel(const HeadP p) {
    IBINN_DYN_SMEM(float, smem);
    __shared__ int s_flag;
    constexpr int NW = NT / 32, NV = SW * CC;
    const int B = p.B, D = p.D, Dp = D >> 2, ncol = Dp >> 7;     // host: D % (512 * UU) == 0
    float* s_mu = smem;                       // [CC / 2][2][D / 4][4]
    float* s_zz = s_mu + (size_t)CC * D;      // [NW][NV]
    float* s_lw = s_zz + NW * NV;             // [CC]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    static_assert(CC % 2 == 0, "class pairs");
    const int Q = D >> 2;                     // 128-bit columns per class
    for (int i = tid; i < (CC / 2) * Q; i += NT) {
        const int kp = i / Q, q = i - kp * Q;
        const float4 a = __ldg(reinterpret_cast<const float4*>(p.mu + (size_t)(2 * kp) * D) + q);
        const float4 c = __ldg(reinterpret_cast<const float4*>(p.mu + (size_t)(2 * kp + 1) * D) + q);
        st4(s_mu + ((size_t)(2 * kp) * Q + q) * 4, make_float4(a.x, c.x, a.y, c.y));
        st4(s_mu + ((size_t)(2 * kp + 1) * Q + q) * 4, make_float4(a.z, c.z, a.w, c.w));
    }
    __shared__ double s_wred[NW][5];
    if (tid < NW * 5) (&s_wred[0][0])[tid] = 0.;
    log_softmax_phi(p.phi, CC, s_lw);
    __syncthreads();
    const int ngroups = (B + SW - 1) / SW;
    float* my_zz = s_zz + wid * NV;
    double* wsum = s_wred[wid];               // lane 0 only
    for (int g = blockIdx.x + gridDim.x * wid; g < ngroups; g += gridDim.x * NW) {
        const int b0 = g * SW, nb = min(SW, B - b0);
        const float* zs[SW];
#pragma unroll
        for (int s = 0; s < SW; ++s) zs[s] = p.z + (size_t)(b0 + min(s, nb - 1)) * D;
        float tot0 = 0.f, tot1 = 0.f;
        int base = 0, lim = 0;
        // Rolling prefetch: a column's registers are refilled with the column UU steps ahead as soon as it has been consumed, so
        // UU x SW 128-bit loads per lane stay in flight across the whole sample (and across the part boundaries).
        float4 zv[UU][SW];
#pragma unroll
        for (int u = 0; u < UU; ++u)
#pragma unroll
            for (int s = 0; s < SW; ++s) zv[u][s] = __ldcs(reinterpret_cast<const float4*>(zs[s] + lane * 4 + 128 * u));
        int pn = 0, cn = UU;                  // (part, column) of the next load
        if (cn == ncol) { cn = 0; pn = 1; }
        int noff = pn * Dp + lane * 4 + 128 * cn;
#pragma unroll 1
        for (int part = 0; part < 4; ++part) {
            f32x2 acc2[SW][CC / 2];
#pragma unroll
            for (int s = 0; s < SW; ++s)
#pragma unroll
                for (int kp = 0; kp < CC / 2; ++kp) acc2[s][kp] = f2_make(0.f, 0.f);
#pragma unroll 1
            for (int c0 = 0; c0 < ncol; c0 += UU) {
#pragma unroll
                for (int u = 0; u < UU; ++u) {
                    const float* mq = s_mu + (size_t)(part * Dp + lane * 4 + 128 * (c0 + u));
#pragma unroll
                    for (int kp = 0; kp < CC / 2; ++kp) {
                        const float4 m01 = ld4_ordered_shared(mq + (size_t)(2 * kp) * D), m23 = ld4_ordered_shared(mq + (size_t)(2 * kp + 1) * D);
                        const f32x2 m0 = f2_make(m01.x, m01.y), m1 = f2_make(m01.z, m01.w);
                        const f32x2 m2 = f2_make(m23.x, m23.y), m3 = f2_make(m23.z, m23.w);
#pragma unroll
                        for (int s = 0; s < SW; ++s) {
                            f32x2 a = acc2[s][kp];
                            f32x2 t = f2_sub(f2_make(zv[u][s].x, zv[u][s].x), m0); a = f2_fma(t, t, a);
                            t = f2_sub(f2_make(zv[u][s].y, zv[u][s].y), m1); a = f2_fma(t, t, a);
                            t = f2_sub(f2_make(zv[u][s].z, zv[u][s].z), m2); a = f2_fma(t, t, a);
                            t = f2_sub(f2_make(zv[u][s].w, zv[u][s].w), m3); a = f2_fma(t, t, a);
                            acc2[s][kp] = a;
                        }
                    }
                    // (unconditional: past the end the last column is simply read again; a branch here makes the compiler sink
                    // all the refills below the arithmetic of the whole UU-column body)
#pragma unroll
                    for (int s = 0; s < SW; ++s) zv[u][s] = ld4_ordered_global(zs[s] + noff);
                    ++cn;
                    const bool wrap = cn == ncol;
                    cn = wrap ? 0 : cn;
                    pn = wrap ? pn + 1 : pn;
                    noff = min(pn, 3) * Dp + lane * 4 + 128 * (pn > 3 ? ncol - 1 : cn);
                }
            }
            float acc[NV];
#pragma unroll
            for (int s = 0; s < SW; ++s)
#pragma unroll
                for (int kp = 0; kp < CC / 2; ++kp) {
                    acc[s * CC + 2 * kp] = f2_lo(acc2[s][kp]);
                    acc[s * CC + 2 * kp + 1] = f2_hi(acc2[s][kp]);
                }
            warp_multi_sum<NV>(acc, lane, base, lim);
            tot0 += acc[0];
            tot1 += acc[1];
        }
        if (lim > 0) my_zz[base] = tot0;
        if (lim > 1) my_zz[base + 1] = tot1;
        __syncwarp();
        for (int s = 0; s < nb; ++s) head_sample_tail(p, b0 + s, my_zz + s * CC, s_lw, lane, wsum);
        __syncwarp();
    }
    // batch means (fp64, fixed order): warp sums -> one record per CTA (front of the partials workspace) -> the last CTA adds the
    // gridDim.x records.  (Reading back one record per SAMPLE in the last CTA was a ~4 us serial tail at B = 8192.)
    __syncthreads();
    double* recs = reinterpret_cast<double*>(p.partials);
    if (tid < 5) {
        double t = 0.;
        for (int w = 0; w < NW; ++w) t += s_wred[w][tid];
        recs[(size_t)blockIdx.x * 5 + tid] = t;
    }
    if (!last_block_ticket(p.counter, gridDim.x, &s_flag)) return;
    if (tid < 5) {
        double t = 0.;
        for (unsigned i = 0; i < gridDim.x; ++i) t += __ldcg(recs + (size_t)i * 5 + tid);
        p.means[tid] = (float)(t / (tid == 1 ? (double)B * CC : (double)B));
    }
    if (tid == 0) *p.counter = 0u;
}

struct HeadBwdP {
    const float* z; const float* mu; const float* phi; const float* y; const float* logits; const float* lse;
    const float* gLx; const float* gLc; const float* gLy; const float* gLogits;   // may be NULL
    int sLx, sLc, sLy, sLogits;                                                   // element strides (0 = broadcast scalar)
    float scale_vec, scale_logits;                                                // applied to the incoming grads
    float* G; float* Glw; float* g_jac;                                           // (B,C), (B,C), (B)
    float* g_z; float* g_mu; float* g_phi; int acc_mu, acc_phi;
    int B, C, D, dchunk;
};

// per sample: G = dLoss/dzz, Glw = dLoss/dlogw contributions, g_jac.  The class posterior is
// re-normalised here (p = exp(l - max) / sum) so that sum_k p_k = 1 to rounding: with logits of
// magnitude 1e3 a saved fp32 log-sum-exp would leave a 1e-4 common factor that the cancelling
// (-y + p) terms amplify.
__global__ void __launch_bounds__(32) head_bwd_g_kernel(const HeadBwdP p) {
    const int C = p.C, b = blockIdx.x, lane = threadIdx.x;
    // log_softmax(phi)
    float pm = -INFINITY;
    for (int k = lane; k < C; k += 32) pm = fmaxf(pm, __ldg(p.phi + k));
    pm = warp_max(pm);
    float ps = 0.f;
    for (int k = lane; k < C; k += 32) ps += expf(__ldg(p.phi + k) - pm);
    ps = warp_sum(ps);
    const float plse = pm + logf(ps);
    const float invD = 1.f / (float)p.D;
    const float glx = p.gLx ? __ldg(p.gLx + (size_t)b * p.sLx) * p.scale_vec : 0.f;
    const float glc = p.gLc ? __ldg(p.gLc + (size_t)b * p.sLc) * p.scale_vec : 0.f;
    const float gly = p.gLy ? __ldg(p.gLy + (size_t)b * p.sLy) * p.scale_vec : 0.f;
    float m = -INFINITY, sy = 0.f;
    for (int k = lane; k < C; k += 32) {
        m = fmaxf(m, __ldg(p.logits + (size_t)b * C + k) + (__ldg(p.phi + k) - plse));
        if (p.y) sy += __ldg(p.y + (size_t)b * C + k);
    }
    m = warp_max(m);
    sy = warp_sum(sy);
    float se = 0.f;
    for (int k = lane; k < C; k += 32) se += expf(__ldg(p.logits + (size_t)b * C + k) + (__ldg(p.phi + k) - plse) - m);
    se = warp_sum(se);
    const float inv_se = 1.f / se;
    for (int k = lane; k < C; k += 32) {
        const float l = __ldg(p.logits + (size_t)b * C + k) + (__ldg(p.phi + k) - plse);
        const float pk = expf(l - m) * inv_se;
        float g = glx * pk * 0.5f * invD;
        if (p.y) {
            const float yk = __ldg(p.y + (size_t)b * C + k);
            g += gly * 0.5f * (pk * sy - yk) + glc * rintf(yk) * 0.5f * invD;
        }
        if (p.gLogits) g -= 0.5f * __ldg(p.gLogits + ((size_t)b * C + k) * p.sLogits) * p.scale_logits;
        p.G[(size_t)b * C + k] = g;
        p.Glw[(size_t)b * C + k] = -glx * pk * invD;
    }
    if (lane == 0) p.g_jac[b] = -(glx + glc) * invD;
}

// per chunk of latent dims: g_z = 2 (z rowsum(G) - G mu), g_mu = -2 (G^T z - mu colsum(G)); block 0 also g_phi
__global__ void __launch_bounds__(IBINN_THREADS) head_bwd_zmu_kernel(const HeadBwdP p) {
    IBINN_DYN_SMEM(float, smem);
    const int B = p.B, C = p.C, D = p.D, DC = p.dchunk;
    const int d0 = blockIdx.x * DC, nd = min(DC, D - d0);
    float* s_G = smem;                 // [B][C]
    float* s_z = s_G + (size_t)B * C;  // [B][DC]
    float* s_mu = s_z + (size_t)B * DC;  // [C][DC]
    float* s_rs = s_mu + (size_t)C * DC; // [B] row sums
    float* s_cs = s_rs + B;              // [C] col sums
    const int tid = threadIdx.x;
    for (int e = tid; e < B * C; e += IBINN_THREADS) s_G[e] = __ldg(p.G + e);
    for (int e = tid; e < B * DC; e += IBINN_THREADS) {
        const int bb = e / DC, d = e - bb * DC;
        s_z[e] = d < nd ? __ldg(p.z + (size_t)bb * D + d0 + d) : 0.f;
    }
    for (int e = tid; e < C * DC; e += IBINN_THREADS) {
        const int k = e / DC, d = e - k * DC;
        s_mu[e] = d < nd ? __ldg(p.mu + (size_t)k * D + d0 + d) : 0.f;
    }
    __syncthreads();
    for (int bb = tid; bb < B; bb += IBINN_THREADS) {
        float s = 0.f;
        for (int k = 0; k < C; ++k) s += s_G[bb * C + k];
        s_rs[bb] = s;
    }
    for (int k = tid; k < C; k += IBINN_THREADS) {
        float s = 0.f;
        for (int bb = 0; bb < B; ++bb) s += s_G[bb * C + k];
        s_cs[k] = s;
    }
    __syncthreads();
    for (int e = tid; e < B * DC; e += IBINN_THREADS) {
        const int bb = e / DC, d = e - bb * DC;
        if (d >= nd) continue;
        float acc = 0.f;
        for (int k = 0; k < C; ++k) acc = fmaf(s_G[bb * C + k], s_mu[k * DC + d], acc);
        p.g_z[(size_t)bb * D + d0 + d] = 2.f * (s_z[e] * s_rs[bb] - acc);
    }
    if (p.g_mu) {
        for (int e = tid; e < C * DC; e += IBINN_THREADS) {
            const int k = e / DC, d = e - k * DC;
            if (d >= nd) continue;
            float acc = 0.f;
            for (int bb = 0; bb < B; ++bb) acc = fmaf(s_G[bb * C + k], s_z[bb * DC + d], acc);
            const float v = -2.f * (acc - s_mu[e] * s_cs[k]);
            float* dst = p.g_mu + (size_t)k * D + d0 + d;
            if (p.acc_mu) *dst += v; else *dst = v;
        }
    }
    if (blockIdx.x == 0 && p.g_phi) {
        // g_logw_k = sum_b Glw[b][k];  g_phi = g_logw - softmax(phi) * sum_k g_logw
        __shared__ float s_lw[1024];
        __shared__ float s_gl[1024];
        __shared__ float s_scr[40];
        log_softmax_phi(p.phi, C, s_lw);
        float part = 0.f;
        for (int k = tid; k < C; k += IBINN_THREADS) {
            float s = 0.f;
            for (int bb = 0; bb < B; ++bb) s += __ldg(p.Glw + (size_t)bb * C + k);
            s_gl[k] = s;
            part += s;
        }
        const float tot = block_sum(part, s_scr);
        for (int k = tid; k < C; k += IBINN_THREADS) {
            const float v = s_gl[k] - expf(s_lw[k]) * tot;
            if (p.acc_phi) p.g_phi[k] += v; else p.g_phi[k] = v;
        }
    }
}

}  // namespace

extern "C" int64_t ibinn_gmm_head_partials_floats(int B) { return (int64_t)B * 5; }

extern "C" int ibinn_gmm_head_forward(const float* z, const float* jac, const float* mu, const float* phi, const float* y,
                                      float* L_x, float* logits, float* L_cNLL, float* L_y, float* correct, float* lse,
                                      float* means, float* partials, uint32_t* counter, int B, int C, int D,
                                      ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!z || !jac || !mu || !phi || !L_x || !logits || !lse || !means || !partials || !counter) return IBINN_ENULL;
    if (y && (!L_cNLL || !L_y || !correct)) return IBINN_ENULL;
    if (B <= 0 || C <= 0 || D <= 0 || C > 1024) return IBINN_EINVAL;
    HeadP p{};
    p.z = z; p.jac = jac; p.mu = mu; p.phi = phi; p.y = y; p.Lx = L_x; p.logits = logits; p.Lc = L_cNLL; p.Ly = L_y;
    p.correct = correct; p.lse = lse; p.means = means; p.partials = partials; p.counter = counter; p.B = B; p.C = C; p.D = D;
    // Scoring batches of the 10-class configs: class means resident in shared memory, z streamed once (head_fwd_stream_kernel), from
    // the batch at which most of its 148 x 16 warps get a pair of samples.  IBINN_HEAD_STREAM = 0 / 2 / 1024 overrides the choice
    // (A/B measurements only; the per-sample results are bit-identical either way).
    {
        static const int force = [] { const char* v = getenv("IBINN_HEAD_STREAM"); return v ? atoi(v) : -1; }();
        const bool can = C == 10 && D % 512 == 0 && (int64_t)B * 5 * 4 >= 148 * 5 * 8 &&
                         ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(mu)) & 15) == 0 && (reinterpret_cast<uintptr_t>(partials) & 7) == 0;
        int sw = !can ? 0 : B >= 3 * 148 * 16 ? 1024 : B >= 2048 ? 2 : 0;
        if (can && force >= 0) sw = force;
#define IBINN_HEAD_STREAM(SW_, NT_, UU_)                                                                                                  \
    do {                                                                                                                                  \
        const size_t sms = ((size_t)C * D + (size_t)(NT_ / 32) * SW_ * C + C + 8) * sizeof(float);                                        \
        if (sms > 227 * 1024 || D % (512 * UU_) != 0) break;                                                                              \
        const int ngroups = (B + SW_ - 1) / SW_;                                                                                          \
        const dim3 grid(min(148, ngroups));     /* groups go round-robin over (CTA, warp): every SM streams */                            \
        e = ibinn_set_smem(head_fwd_stream_kernel<SW_, 10, NT_, UU_>, sms);                                                               \
        if (e) return e;                                                                                                                  \
        IBINN_LAUNCH((head_fwd_stream_kernel<SW_, 10, NT_, UU_>), grid, dim3(NT_), sms, (cudaStream_t)stream, p);                         \
        return ibinn_launch_status();                                                                                                     \
    } while (0)
        if (sw == 2) IBINN_HEAD_STREAM(2, 512, 3);               // 128 registers: three columns of z in flight per lane
        else if (sw == 1024) IBINN_HEAD_STREAM(2, 1024, 2);      // 64 registers, 32 warps per SM: best at B = 8192 (38.5 vs 40.4 us)
#undef IBINN_HEAD_STREAM
    }
    // samples per CTA: as many as still leave two CTAs for each of the 148 SMs; 8 only where the class means dominate the traffic
    // (4 samples = 48 KB of z per CTA keeps four CTAs per SM resident, which hides the load latencies better at C = 10)
    const int sb = (B >= 148 * 2 * 8 && C >= 32) ? 8 : B >= 148 * 2 * 4 ? 4 : B >= 148 * 2 * 2 ? 2 : 1;
    const size_t sm = ((size_t)sb * D + (size_t)(sb + 1) * C + (size_t)sb * C * 4 + 8) * sizeof(float);
    const dim3 grid((B + sb - 1) / sb);
#define IBINN_HEAD_FWD(SB_)                                                                           \
    do {                                                                                              \
        e = ibinn_set_smem(head_fwd_kernel<SB_>, sm);                                                 \
        if (e) return e;                                                                              \
        IBINN_LAUNCH(head_fwd_kernel<SB_>, grid, dim3(IBINN_THREADS), sm, (cudaStream_t)stream, p);   \
    } while (0)
    if (sb == 8) IBINN_HEAD_FWD(8);
    else if (sb == 4) IBINN_HEAD_FWD(4);
    else if (sb == 2) IBINN_HEAD_FWD(2);
    else IBINN_HEAD_FWD(1);
#undef IBINN_HEAD_FWD
    return ibinn_launch_status();
}

extern "C" int ibinn_gmm_head_backward(const float* z, const float* mu, const float* phi, const float* y, const float* logits,
                                       const float* lse, const float* g_Lx, int s_Lx, const float* g_LcNLL, int s_Lc,
                                       const float* g_Ly, int s_Ly, const float* g_logits, int s_logits, float scale_vec,
                                       float scale_logits, float* G_ws, float* Glw_ws, float* g_z, float* g_jac, float* g_mu,
                                       int accumulate_mu, float* g_phi, int accumulate_phi, int B, int C, int D,
                                       ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!z || !mu || !phi || !logits || !lse || !G_ws || !Glw_ws || !g_z || !g_jac) return IBINN_ENULL;
    if (B <= 0 || C <= 0 || D <= 0 || C > 1024) return IBINN_EINVAL;
    HeadBwdP p{};
    p.z = z; p.mu = mu; p.phi = phi; p.y = y; p.logits = logits; p.lse = lse;
    p.gLx = g_Lx; p.gLc = g_LcNLL; p.gLy = g_Ly; p.gLogits = g_logits;
    p.sLx = s_Lx; p.sLc = s_Lc; p.sLy = s_Ly; p.sLogits = s_logits; p.scale_vec = scale_vec; p.scale_logits = scale_logits;
    p.G = G_ws; p.Glw = Glw_ws; p.g_jac = g_jac; p.g_z = g_z; p.g_mu = g_mu; p.g_phi = g_phi;
    p.acc_mu = accumulate_mu; p.acc_phi = accumulate_phi; p.B = B; p.C = C; p.D = D;
    cudaStream_t st = (cudaStream_t)stream;
    IBINN_LAUNCH(head_bwd_g_kernel, dim3(B), dim3(32), 0, st, p);
    e = ibinn_launch_status();
    if (e) return e;
    int dc = 32;
    size_t sm;
    for (;;) {
        sm = ((size_t)B * C + (size_t)B * dc + (size_t)C * dc + B + C + 8) * sizeof(float);
        if (sm <= 200 * 1024 || dc == 4) break;
        dc /= 2;
    }
    if (sm > 200 * 1024) return IBINN_EINVAL;
    p.dchunk = dc;
    e = ibinn_set_smem(head_bwd_zmu_kernel, sm);
    if (e) return e;
    IBINN_LAUNCH(head_bwd_zmu_kernel, dim3((D + dc - 1) / dc), dim3(IBINN_THREADS), sm, st, p);
    return ibinn_launch_status();
}
