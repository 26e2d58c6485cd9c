// Affine coupling kernels of the IB-INN nodes (HBM-bound elementwise + per-sample log-det):
//   * AIO_Block: tanh-clamped affine coupling + ActNorm + fixed orthogonal 1x1 "soft permutation"
//     fused in one kernel per direction (reference all_in_one_block.py:61-114);
//   * DownsampleCouplingBlock: space-to-depth fused with the affine coupling
//     (reference downsampling_coupling_block.py:49-97);
//   * GLOWCouplingBlock (FrEIA v0.2, used at inn_architecture.py:161): atan-clamped affine halves.
// One CTA owns one sample, so the per-sample log-det is a fixed-order in-CTA reduction
// (warp shuffles + one shared-memory pass) and is written, never atomically accumulated.
#include "common.cuh"

namespace {

__device__ __forceinline__ int round4(int v) { return (v + 3) & ~3; }

// ------------------------------------------------------------------------------------ AIO
struct AioP {
    const float* x; const float* a; const float* w; const float* an; const float* off;
    float* out; float* jac;
    int B, C, HW; float clamp, alpha;
    // backward only
    const float* g_out; const float* g_jac; const float* out_saved;
    float* g_x; float* g_a; float* g_an; float* g_off; float* partials; unsigned* counter;
    int nsplit, defer;
    int hw_shift;        // log2(HW) when HW is a power of two >= 4 and x / a / out are 16-byte aligned (vector path), else 0
};

// MODE 0: forward   out = (w . [x1 | x2*exp(sig)+t]) * exp(an) + off ; jac = sum sig + HW*sum an
// MODE 1: unpermute u = w_inv . ((y - off) * exp(-an))                (first half of the inverse)
template <int MODE>
__global__ void __launch_bounds__(IBINN_THREADS, 4) aio_mix_kernel(const AioP p) {
    IBINN_DYN_SMEM(float, smem);
    __shared__ float s_scr[40];
    __shared__ float s_san;
    const int C = p.C, HW = p.HW, HWp = round4(HW), Cp = round4(C);
    const int c1 = C - C / 2, c2 = C / 2;
    float* s_u = smem;                 // [C][HWp]
    float* s_w = s_u + (size_t)C * HWp;  // [C(in)][Cp(out)]
    float* s_sc = s_w + (size_t)C * Cp;  // [C] exp(+-an)
    float* s_of = s_sc + Cp;             // [C]
    const int tid = threadIdx.x, b = blockIdx.x;
    const float* xb = p.x + (size_t)b * C * HW;

    // (warp <-> input channel, lane <-> output channel: no integer divisions in the prologue)
    for (int i = tid >> 5; i < C; i += IBINN_THREADS / 32)
        for (int o = tid & 31; o < Cp; o += 32) s_w[i * Cp + o] = (o < C) ? __ldg(p.w + (size_t)o * C + i) : 0.f;
    for (int c = tid; c < C; c += IBINN_THREADS) {
        s_sc[c] = expf(MODE == 0 ? p.an[c] : -p.an[c]);
        s_of[c] = p.off[c];
    }
    if (MODE == 0 && (tid >> 5) == IBINN_THREADS / 32 - 1) {      // sum of the ActNorm log-scales for the log-det, off the tail
        float v = 0.f;
        for (int c = tid & 31; c < C; c += 32) v += p.an[c];
        v = warp_sum(v);
        if ((tid & 31) == 0) s_san = v;
    }
    if (MODE == 1) __syncthreads();

    float jsum = 0.f;
    if (p.hw_shift) {
        // vector path: 128-bit loads, channel / pixel from shifts (the scalar path below spends ~40 instructions per element on the
        // two runtime divisions, which made the kernel issue-bound at the scoring batch: 100 M warp instructions at B = 8192)
        constexpr int KB4 = 3;
        const int sh = p.hw_shift, n4 = (C * HW) >> 2;
        const float* ab = p.a + (size_t)b * 2 * c2 * HW;
        for (int i0 = 0; i0 < n4; i0 += IBINN_THREADS * KB4) {
            float4 xv[KB4], sv[KB4], tv[KB4];
#pragma unroll
            for (int k = 0; k < KB4; ++k) {
                const int e = (i0 + tid + IBINN_THREADS * k) * 4;
                xv[k] = make_float4(0.f, 0.f, 0.f, 0.f); sv[k] = xv[k]; tv[k] = xv[k];
                if (e < n4 * 4) {
                    xv[k] = __ldg(reinterpret_cast<const float4*>(xb + e));
                    if (MODE == 0) {
                        const int c = e >> sh, q = e & (HW - 1);
                        if (c >= c1) {
                            sv[k] = __ldg(reinterpret_cast<const float4*>(ab + (size_t)(c - c1) * HW + q));
                            tv[k] = __ldg(reinterpret_cast<const float4*>(ab + (size_t)(c2 + c - c1) * HW + q));
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < KB4; ++k) {
                const int e = (i0 + tid + IBINN_THREADS * k) * 4;
                if (e >= n4 * 4) continue;
                const int c = e >> sh, q = e & (HW - 1);
                float u[4] = {xv[k].x, xv[k].y, xv[k].z, xv[k].w};
                if (MODE == 0) {
                    if (c >= c1) {
                        const float ss[4] = {sv[k].x, sv[k].y, sv[k].z, sv[k].w}, tt[4] = {tv[k].x, tv[k].y, tv[k].z, tv[k].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float sig = p.clamp * tanhf(p.alpha * ss[j]);
                            u[j] = fmaf(u[j], expf(sig), tt[j]);
                            jsum += sig;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) u[j] = (u[j] - s_of[c]) * s_sc[c];
                }
                st4(s_u + (size_t)c * HWp + q, make_float4(u[0], u[1], u[2], u[3]));
            }
        }
    } else {
        // all loads of a batch of elements are issued before the first use (the kernel is latency-bound: 12 elements per thread)
        constexpr int KB = 12;
        const int n = C * HW;
        const float* ab = p.a + (size_t)b * 2 * c2 * HW;
        for (int e0 = 0; e0 < n; e0 += IBINN_THREADS * KB) {
            float xv[KB], sv[KB], tv[KB];
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int e = e0 + tid + IBINN_THREADS * k;
                xv[k] = 0.f; sv[k] = 0.f; tv[k] = 0.f;
                if (e < n) {
                    xv[k] = __ldg(xb + e);
                    if (MODE == 0) {
                        const int c = e / HW, q = e - c * HW;
                        if (c >= c1) {
                            sv[k] = __ldg(ab + (size_t)(c - c1) * HW + q);
                            tv[k] = __ldg(ab + (size_t)(c2 + c - c1) * HW + q);
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int e = e0 + tid + IBINN_THREADS * k;
                if (e >= n) continue;
                const int c = e / HW, q = e - c * HW;
                float u = xv[k];
                if (MODE == 0) {
                    if (c >= c1) {
                        const float sig = p.clamp * tanhf(p.alpha * sv[k]);
                        u = fmaf(u, expf(sig), tv[k]);
                        jsum += sig;
                    }
                } else {
                    u = (u - s_of[c]) * s_sc[c];
                }
                s_u[(size_t)c * HWp + q] = u;
            }
        }
    }
    // zero the pixel padding so that the 4-wide reads below stay finite
    if (HWp != HW)
        for (int c = tid; c < C; c += IBINN_THREADS)
            for (int q = HW; q < HWp; ++q) s_u[(size_t)c * HWp + q] = 0.f;
    __syncthreads();

    const int npg = HWp / 4, nog = Cp / 4;
    float* ob = p.out + (size_t)b * C * HW;
    const bool vec = (HW % 4) == 0;
    for (int item = tid; item < npg * nog; item += IBINN_THREADS) {
        const int pg = item % npg, og = item / npg;
        // 4 output channels x 4 pixels per thread; the pixels are held as two packed pairs so that the update is 8 FFMA2 (weight
        // broadcast to both halves) instead of 16 FFMA per input channel: at C = 48 the mix is FMA-issue-bound, not HBM-bound.
        // Every component rounds like the scalar fmaf, so the results are unchanged.
        f32x2 acc2[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc2[i][0] = f2_make(0.f, 0.f); acc2[i][1] = f2_make(0.f, 0.f); }
        for (int i = 0; i < C; ++i) {
            const float4 uv = ld4(s_u + (size_t)i * HWp + pg * 4);
            const float4 wv = ld4(s_w + (size_t)i * Cp + og * 4);
            const f32x2 u01 = f2_make(uv.x, uv.y), u23 = f2_make(uv.z, uv.w);
            const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const f32x2 w2 = f2_make(ww[o], ww[o]);
                acc2[o][0] = f2_fma(w2, u01, acc2[o][0]);
                acc2[o][1] = f2_fma(w2, u23, acc2[o][1]);
            }
        }
        float acc[4][4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            acc[o][0] = f2_lo(acc2[o][0]); acc[o][1] = f2_hi(acc2[o][0]);
            acc[o][2] = f2_lo(acc2[o][1]); acc[o][3] = f2_hi(acc2[o][1]);
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int oc = og * 4 + o;
            if (oc >= C) continue;
            float r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = (MODE == 0) ? fmaf(acc[o][j], s_sc[oc], s_of[oc]) : acc[o][j];
            float* dst = ob + (size_t)oc * HW + pg * 4;
            if (vec) st4(dst, make_float4(r[0], r[1], r[2], r[3]));
            else
#pragma unroll
                for (int j = 0; j < 4; ++j) if (pg * 4 + j < HW) dst[j] = r[j];
        }
    }
    if (MODE == 0) {
        const float tot = block_sum(jsum, s_scr);
        if (tid == 0) p.jac[b] = tot + (float)HW * s_san;
    }
}

// second half of the inverse: x2 = (u2 - t) * exp(-sig), in place on u; jac = -sum sig - HW*sum an
__global__ void __launch_bounds__(IBINN_THREADS) aio_affine_inverse_kernel(const AioP p) {
    __shared__ float s_scr[40];
    const int C = p.C, HW = p.HW, c1 = C - C / 2, c2 = C / 2;
    const int tid = threadIdx.x, b = blockIdx.x;
    float* ub = p.out + (size_t)b * C * HW + (size_t)c1 * HW;
    const float* ab = p.a + (size_t)b * 2 * c2 * HW;
    float jsum = 0.f;
    for (int e = tid; e < c2 * HW; e += IBINN_THREADS) {
        const float sig = p.clamp * tanhf(p.alpha * __ldg(ab + e));
        ub[e] = (ub[e] - __ldg(ab + (size_t)c2 * HW + e)) * expf(-sig);
        jsum += sig;
    }
    const float tot = block_sum(jsum, s_scr);
    if (tid == 0) {
        float san = 0.f;
        for (int c = 0; c < C; ++c) san += p.an[c];
        p.jac[b] = -tot - (float)HW * san;
    }
}

// backward of MODE 0.  Needs x, a, out (the block output = next block's input) and g_out.
#ifndef IBINN_AIO_BWD_MINB
// Occupancy is what this kernel lives on (its phases are load -> sync -> compute -> store per sample): unbounded it takes 136 registers =
// ONE CTA per SM; 3 -> 79 registers, no spills; 4 -> 64 registers, 124 bytes of spill.  B = 8192, C = 3 / 12 / 48: 559 / 706 / 873 us (1),
// 208 / 288 / 399 us (3), 186 / 241 / 335 us (4); training step at B = 128: 9.51 ms (1), 8.97 ms (3 and 4) - the smaller footprint also
// lets the weight-gradient stream's CTAs share the SMs.
#define IBINN_AIO_BWD_MINB 4
#endif
__global__ void __launch_bounds__(IBINN_THREADS, IBINN_AIO_BWD_MINB) aio_bwd_kernel(const AioP p) {
    IBINN_DYN_SMEM(float, smem);
    __shared__ int s_flag;
    ibinn_pdl_wait();           // (common.cuh: launched with IBINN_LAUNCH_PDL)
    ibinn_pdl_trigger();
    // A sample may be split over `nsplit` CTAs along the pixel axis (HWl pixels each, starting at qo): nothing in the backward
    // couples pixels except the ActNorm sums, which are per-CTA partial records anyway.
    const int C = p.C, HW = p.HW, HWl = HW / p.nsplit, HWp = round4(HWl), Cp = round4(C);
    const int c1 = C - C / 2, c2 = C / 2;
    float* s_u = smem;                   // g_v [C][HWp]
    float* s_w = s_u + (size_t)C * HWp;  // [C(out)][Cp(in)]
    const int tid = threadIdx.x, b = blockIdx.x / p.nsplit, qo = (blockIdx.x % p.nsplit) * HWl, lane = tid & 31, wid = tid >> 5;
    const size_t ib = (size_t)b * C * HW + qo;
    const float gj = p.g_jac ? __ldg(p.g_jac + b) : 0.f;

    for (int o = wid; o < C; o += IBINN_THREADS / 32)          // warp <-> output channel, lane <-> input channel: no divisions
        for (int i = lane; i < Cp; i += 32) s_w[o * Cp + i] = (i < C) ? __ldg(p.w + (size_t)o * C + i) : 0.f;
    // phase A: g_v = g_out * exp(an) into shared memory, per-channel sums of g and g (y - off) for the ActNorm gradients
    if (HWl % 32 == 0) {
        // fast path: every load of the sample is in flight at once; a warp's 32 consecutive elements share a channel, so
        // a warp reduction per element group and a fixed-order sum of the HW/32 group partials per channel finish the job
        float* s_sc = s_w + (size_t)C * Cp;
        float* s_of = s_sc + Cp;
        float* s_part = s_of + Cp + 16;                      // [C * HW / 32][2]
        for (int c = tid; c < C; c += IBINN_THREADS) { s_sc[c] = expf(p.an[c]); s_of[c] = p.off[c]; }
        __syncthreads();
        constexpr int KB = 12;
        const int n = C * HWl;
        const int shl = p.hw_shift;        // log2(HWl) when it is a power of two: channel / pixel by shifts instead of two divisions per element
        for (int e0 = 0; e0 < n; e0 += IBINN_THREADS * KB) {
            float g[KB], y[KB];
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int e = e0 + tid + IBINN_THREADS * k;
                g[k] = 0.f; y[k] = 0.f;
                if (e < n) {
                    const int o = shl ? (e >> shl) : e / HWl, q = e - o * HWl;
                    g[k] = __ldg(p.g_out + ib + (size_t)o * HW + q); y[k] = __ldg(p.out_saved + ib + (size_t)o * HW + q);
                }
            }
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int e = e0 + tid + IBINN_THREADS * k;
                if (e - lane >= n) continue;                 // whole warp in or out (n is a multiple of 32)
                const int o = shl ? (e >> shl) : e / HWl, q = e - o * HWl;
                s_u[(size_t)o * HWp + q] = g[k] * s_sc[o];
                const float so = warp_sum(g[k]), sn = warp_sum(g[k] * (y[k] - s_of[o]));
                if (lane == 0) { s_part[(e >> 5) * 2] = so; s_part[(e >> 5) * 2 + 1] = sn; }
            }
        }
        __syncthreads();
        if (tid < C) {
            const int per = HWl / 32;
            float so = 0.f, sn = 0.f;
            for (int i = tid * per; i < (tid + 1) * per; ++i) { so += s_part[2 * i]; sn += s_part[2 * i + 1]; }
            p.partials[((size_t)blockIdx.x * 2 + 0) * C + tid] = so;
            p.partials[((size_t)blockIdx.x * 2 + 1) * C + tid] = sn + (float)HWl * gj;
        }
    } else {
    for (int o = wid; o < C; o += IBINN_THREADS / 32) {
        const float sc = expf(p.an[o]), of = p.off[o];
        float so = 0.f, sn = 0.f;
        for (int q = lane; q < HWp; q += 32) {
            float gv = 0.f;
            if (q < HWl) {
                const float g = __ldg(p.g_out + ib + (size_t)o * HW + q);
                const float y = __ldg(p.out_saved + ib + (size_t)o * HW + q);
                so += g;
                sn = fmaf(g, y - of, sn);
                gv = g * sc;
            }
            s_u[(size_t)o * HWp + q] = gv;
        }
        so = warp_sum(so);
        sn = warp_sum(sn);
        if (lane == 0) {
            p.partials[((size_t)blockIdx.x * 2 + 0) * C + o] = so;
            p.partials[((size_t)blockIdx.x * 2 + 1) * C + o] = sn + (float)HWl * gj;
        }
    }
    }
    __syncthreads();

    // phase B: g_u = w^T g_v, then the affine-coupling backward per element
    const int npg = HWp / 4, nig = Cp / 4;
    const bool vec = (HW % 4) == 0 && (HWl % 4) == 0;
    const float* ab = p.a + (size_t)b * 2 * c2 * HW + qo;
    float* gab = p.g_a + (size_t)b * 2 * c2 * HW + qo;
    const float ak = p.alpha * p.clamp;
    for (int item = tid; item < npg * nig; item += IBINN_THREADS) {
        const int pg = item % npg, ig = item / npg;
        // the coupling operands (s and x2 of the passive-half channels) do not depend on the channel mix: fetch them first so
        // that their latency hides behind the C-deep FMA loop
        float sv[4][4], xv[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ic = ig * 4 + i;
#pragma unroll
            for (int j = 0; j < 4; ++j) { sv[i][j] = 0.f; xv[i][j] = 0.f; }
            if (ic >= C || ic < c1) continue;
            const float* sp = ab + (size_t)(ic - c1) * HW + pg * 4;
            const float* xp = p.x + ib + (size_t)ic * HW + pg * 4;
            if (vec) {
                const float4 s4 = ld4(sp), x4 = ld4(xp);
                sv[i][0] = s4.x; sv[i][1] = s4.y; sv[i][2] = s4.z; sv[i][3] = s4.w;
                xv[i][0] = x4.x; xv[i][1] = x4.y; xv[i][2] = x4.z; xv[i][3] = x4.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (pg * 4 + j < HWl) { sv[i][j] = __ldg(sp + j); xv[i][j] = __ldg(xp + j); }
            }
        }
        f32x2 acc2[4][2];                     // packed pixel pairs: 8 FFMA2 per output channel (see aio_mix_kernel)
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc2[i][0] = f2_make(0.f, 0.f); acc2[i][1] = f2_make(0.f, 0.f); }
        for (int o = 0; o < C; ++o) {
            const float4 uv = ld4(s_u + (size_t)o * HWp + pg * 4);
            const float4 wv = ld4(s_w + (size_t)o * Cp + ig * 4);
            const f32x2 u01 = f2_make(uv.x, uv.y), u23 = f2_make(uv.z, uv.w);
            const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const f32x2 w2 = f2_make(ww[i], ww[i]);
                acc2[i][0] = f2_fma(w2, u01, acc2[i][0]);
                acc2[i][1] = f2_fma(w2, u23, acc2[i][1]);
            }
        }
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[i][0] = f2_lo(acc2[i][0]); acc[i][1] = f2_hi(acc2[i][0]);
            acc[i][2] = f2_lo(acc2[i][1]); acc[i][3] = f2_hi(acc2[i][1]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ic = ig * 4 + i;
            if (ic >= C) continue;
            float gx[4];
            if (ic < c1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) gx[j] = acc[i][j];
            } else {
                const int k = ic - c1;
                float gs[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int q = pg * 4 + j;
                    gx[j] = 0.f; gs[j] = 0.f;
                    if (q < HWl) {
                        const float th = tanhf(p.alpha * sv[i][j]);
                        const float es = expf(p.clamp * th);
                        const float gu = acc[i][j];
                        gx[j] = gu * es;
                        gs[j] = (gu * xv[i][j] * es + gj) * ak * (1.f - th * th);
                    }
                }
                float* ds = gab + (size_t)k * HW + pg * 4;
                float* dt = gab + (size_t)(c2 + k) * HW + pg * 4;
                if (vec) {
                    st4(ds, make_float4(gs[0], gs[1], gs[2], gs[3]));
                    st4(dt, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (pg * 4 + j < HWl) { ds[j] = gs[j]; dt[j] = acc[i][j]; }
                }
            }
            float* dx = p.g_x + ib + (size_t)ic * HW + pg * 4;
            if (vec) st4(dx, make_float4(gx[0], gx[1], gx[2], gx[3]));
            else
#pragma unroll
                for (int j = 0; j < 4; ++j) if (pg * 4 + j < HWl) dx[j] = gx[j];
        }
    }

    // ActNorm parameter gradients: the last CTA sums the per-sample partials (fixed order: sample groups, then groups)
    if (p.defer) return;                 // the caller sums the records of many blocks with one launch (ibinn_aio_param_reduce_many)
    if (!last_block_ticket(p.counter, gridDim.x, &s_flag)) return;
    {
        const int nrec = p.B * p.nsplit;
        const int T = 2 * C, ngrp = IBINN_THREADS / T > 0 ? IBINN_THREADS / T : 1;
        const bool par = (ngrp * T <= C * HWp) && (T <= IBINN_THREADS);
        if (par) {
            const int e = tid % T, grp = tid / T;
            float t = 0.f;
            if (grp < ngrp) {
                constexpr int UB = 8;
                for (int b0 = grp; b0 < nrec; b0 += ngrp * UB) {
                    float v[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        const int bb = b0 + u * ngrp;
                        v[u] = bb < nrec ? __ldcg(p.partials + (size_t)bb * T + e) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < UB; ++u) t += v[u];
                }
                s_u[grp * T + e] = t;
            }
            __syncthreads();
            if (tid < T) {
                float tot = 0.f;
                for (int gi = 0; gi < ngrp; ++gi) tot += s_u[gi * T + tid];
                if (tid < C) { if (p.g_off) p.g_off[tid] += tot; }
                else if (p.g_an) p.g_an[tid - C] += tot;
            }
        } else {
            for (int e = tid; e < T; e += IBINN_THREADS) {
                float t = 0.f;
                for (int bb = 0; bb < nrec; ++bb) t += __ldcg(p.partials + (size_t)bb * T + e);
                if (e < C) { if (p.g_off) p.g_off[e] += t; }
                else if (p.g_an) p.g_an[e - C] += t;
            }
        }
    }
    if (tid == 0) *p.counter = 0u;
}

// log2(HW) if the vector path of aio_mix_kernel applies (HW a power of two >= 4, 16-byte aligned operands), else 0
static int aio_hw_shift(int HW, const void* x, const void* a, const void* out) {
    if (HW < 4 || (HW & (HW - 1)) != 0) return 0;
    if (((uintptr_t)x | (uintptr_t)a | (uintptr_t)out) & 15) return 0;
    int sh = 0;
    while ((1 << sh) < HW) ++sh;
    return sh;
}

static size_t aio_smem(int C, int HW) {
    const int HWp = (HW + 3) & ~3, Cp = (C + 3) & ~3;
    return ((size_t)C * HWp + (size_t)C * Cp + 2 * Cp + 32 + 2 * ((size_t)C * HWp / 32 + 1)) * sizeof(float);
}

// ------------------------------------------------------------------------------------ S2D
struct S2dP {
    const float* x; long long x_bs; const float* a; float* out; long long out_bs; float* jac; int jac_acc;
    const float* g_out; long long g_bs; const float* g_jac; float* g_x; long long gx_bs; float* g_a;
    int B, cp, H, W; float clamp, alpha;
    int wo_shift, hwo_shift;     // log2 of the output width / plane when both are powers of two (index arithmetic by shifts), else 0
};

// MODE 0 forward, 1 inverse (x <- out), 2 backward
template <int MODE>
__global__ void __launch_bounds__(IBINN_THREADS) s2d_kernel(const S2dP p) {
    __shared__ float s_scr[40];
    const int Ho = p.H / 2, Wo = p.W / 2, HWo = Ho * Wo, n = 4 * p.cp * HWo;
    const int tid = threadIdx.x, b = blockIdx.x;
    const float* ab = p.a + (size_t)b * 8 * p.cp * HWo;
    const float gj = (MODE == 2 && p.g_jac) ? __ldg(p.g_jac + b) : 0.f;
    const float ak = p.alpha * p.clamp;
    float jsum = 0.f;
    if (MODE == 0 && p.hwo_shift > 0) {
        // forward at power-of-two extents (every CIFAR stage): no divisions, and the 12 loads of 4 elements are issued before the
        // first use (the scalar loop below is issue-bound on its two runtime divisions and exposes one load latency per element)
        constexpr int U = 4;
        const int ws = p.wo_shift, hs = p.hwo_shift;
        const float* xb = p.x + (size_t)b * p.x_bs;
        float* ob = p.out + (size_t)b * p.out_bs;
        for (int e0 = tid; e0 < n; e0 += IBINN_THREADS * U) {
            float sv[U], tv[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * IBINN_THREADS;
                sv[u] = 0.f; tv[u] = 0.f; xv[u] = 0.f;
                if (e < n) {
                    const int oc = e >> hs, q = e & (HWo - 1), h = q >> ws, w = q & (Wo - 1);
                    sv[u] = __ldg(ab + e);
                    tv[u] = __ldg(ab + (size_t)n + e);
                    xv[u] = __ldg(xb + ((size_t)(oc >> 2) * p.H + 2 * h + ((oc >> 1) & 1)) * p.W + 2 * w + (oc & 1));
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * IBINN_THREADS;
                if (e >= n) continue;
                const float sig = p.clamp * tanhf(p.alpha * sv[u]);
                ob[e] = fmaf(xv[u], expf(sig), tv[u]);
                jsum += sig;
            }
        }
    } else
    for (int e = tid; e < n; e += IBINN_THREADS) {
        const int oc = e / HWo, q = e - oc * HWo;
        const int h = q / Wo, w = q - h * Wo;
        const int c = oc >> 2, dy = (oc >> 1) & 1, dx = oc & 1;
        const size_t xi = ((size_t)c * p.H + 2 * h + dy) * p.W + 2 * w + dx;
        const float s = __ldg(ab + e);
        const float th = tanhf(p.alpha * s);
        const float sig = p.clamp * th;
        if (MODE == 0) {
            const float t = __ldg(ab + (size_t)n + e);
            p.out[(size_t)b * p.out_bs + e] = fmaf(__ldg(p.x + (size_t)b * p.x_bs + xi), expf(sig), t);
            jsum += sig;
        } else if (MODE == 1) {
            const float t = __ldg(ab + (size_t)n + e);
            // here `x` is the output tensor (full resolution) and `out` the block-side tensor
            const_cast<float*>(p.x)[(size_t)b * p.x_bs + xi] = (p.out[(size_t)b * p.out_bs + e] - t) * expf(-sig);
            jsum -= sig;
        } else {
            const float es = expf(sig);
            const float g = __ldg(p.g_out + (size_t)b * p.g_bs + e);
            const float xv = __ldg(p.x + (size_t)b * p.x_bs + xi);
            p.g_x[(size_t)b * p.gx_bs + xi] = g * es;
            float* gab = p.g_a + (size_t)b * 8 * p.cp * HWo;
            gab[e] = (g * xv * es + gj) * ak * (1.f - th * th);
            gab[(size_t)n + e] = g;
        }
    }
    if (MODE != 2) {
        const float tot = block_sum(jsum, s_scr);
        if (tid == 0) p.jac[b] = (p.jac_acc ? p.jac[b] : 0.f) + tot;
    }
}

// ------------------------------------------------------------------------------------ GLOW
struct GlowP {
    const float* x; long long x_rs; const float* r; float* y; long long y_rs; float* jac; int jac_acc;
    const float* g_y; long long gy_rs; const float* g_jac; float* g_x; long long gx_rs; float* g_r;
    int B, L; float clamp;
};

__device__ __forceinline__ float glow_le(float s, float clamp) { return clamp * 0.636f * atanf(s / clamp); }

template <int MODE>
__global__ void __launch_bounds__(IBINN_THREADS) glow_kernel(const GlowP p) {
    __shared__ float s_scr[40];
    const int tid = threadIdx.x, b = blockIdx.x, L = p.L;
    const float* rb = p.r + (size_t)b * 2 * L;
    const float gj = (MODE == 2 && p.g_jac) ? __ldg(p.g_jac + b) : 0.f;
    float jsum = 0.f;
    if (MODE != 2 && (L & 3) == 0 && ((p.x_rs | p.y_rs) & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.y) | reinterpret_cast<uintptr_t>(p.r)) & 15) == 0) {
        // 128-bit loads, two groups (24 floats) in flight per thread: the scalar loop below exposes one load latency per element
        constexpr int U = 2;
        const float* xb = p.x + (size_t)b * p.x_rs;
        float* yb = p.y + (size_t)b * p.y_rs;
        for (int i0 = tid * 4; i0 < L; i0 += IBINN_THREADS * 4 * U) {
            float4 sv[U], tv[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * IBINN_THREADS * 4;
                sv[u] = make_float4(0.f, 0.f, 0.f, 0.f); tv[u] = sv[u]; xv[u] = sv[u];
                if (i < L) {
                    sv[u] = __ldg(reinterpret_cast<const float4*>(rb + i));
                    tv[u] = __ldg(reinterpret_cast<const float4*>(rb + L + i));
                    xv[u] = __ldg(reinterpret_cast<const float4*>(xb + i));
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * IBINN_THREADS * 4;
                if (i >= L) continue;
                const float ss[4] = {sv[u].x, sv[u].y, sv[u].z, sv[u].w}, tt[4] = {tv[u].x, tv[u].y, tv[u].z, tv[u].w};
                const float xx[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
                float r[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float le = glow_le(ss[j], p.clamp);
                    if (MODE == 0) { r[j] = fmaf(expf(le), xx[j], tt[j]); jsum += le; }
                    else { r[j] = (xx[j] - tt[j]) * expf(-le); jsum -= le; }
                }
                st4(yb + i, make_float4(r[0], r[1], r[2], r[3]));
            }
        }
    } else
    for (int i = tid; i < L; i += IBINN_THREADS) {
        const float s = __ldg(rb + i);
        const float le = glow_le(s, p.clamp);
        if (MODE == 0) {
            p.y[(size_t)b * p.y_rs + i] = fmaf(expf(le), __ldg(p.x + (size_t)b * p.x_rs + i), __ldg(rb + L + i));
            jsum += le;
        } else if (MODE == 1) {
            p.y[(size_t)b * p.y_rs + i] = (__ldg(p.x + (size_t)b * p.x_rs + i) - __ldg(rb + L + i)) * expf(-le);
            jsum -= le;
        } else {
            const float e = expf(le);
            const float g = __ldg(p.g_y + (size_t)b * p.gy_rs + i);
            const float xv = __ldg(p.x + (size_t)b * p.x_rs + i);
            const float q = s / p.clamp;
            p.g_x[(size_t)b * p.gx_rs + i] = g * e;
            p.g_r[(size_t)b * 2 * L + i] = (g * xv * e + gj) * (0.636f / (1.f + q * q));
            p.g_r[(size_t)b * 2 * L + L + i] = g;
        }
    }
    if (MODE != 2) {
        const float tot = block_sum(jsum, s_scr);
        if (tid == 0) p.jac[b] = (p.jac_acc ? p.jac[b] : 0.f) + tot;
    }
}

}  // namespace

#define IBINN_ENTRY_GUARD()                   \
    do {                                      \
        int e_ = ibinn_check_device();        \
        if (e_) return e_;                    \
    } while (0)

extern "C" int ibinn_aio_coupling_forward(const float* x, const float* a, const float* w, const float* act_norm,
                                          const float* act_offset, float* out, float* jac, int B, int C, int H, int W,
                                          float clamp, float alpha, ibinn_stream_t stream) {
    IBINN_ENTRY_GUARD();
    if (!x || !w || !act_norm || !act_offset || !out || !jac || (C > 1 && !a)) return IBINN_ENULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return IBINN_EINVAL;
    AioP p{};
    p.x = x; p.a = a; p.w = w; p.an = act_norm; p.off = act_offset; p.out = out; p.jac = jac;
    p.B = B; p.C = C; p.HW = H * W; p.clamp = clamp; p.alpha = alpha;
    p.hw_shift = aio_hw_shift(H * W, x, a, out);
    size_t sm = aio_smem(C, H * W);
    int e = ibinn_set_smem(aio_mix_kernel<0>, sm);
    if (e) return e;
    IBINN_LAUNCH(aio_mix_kernel<0>, dim3(B), dim3(IBINN_THREADS), sm, (cudaStream_t)stream, p);
    return ibinn_launch_status();
}

extern "C" int ibinn_aio_unpermute(const float* y, const float* w_inv, const float* act_norm, const float* act_offset,
                                   float* u, int B, int C, int H, int W, ibinn_stream_t stream) {
    IBINN_ENTRY_GUARD();
    if (!y || !w_inv || !act_norm || !act_offset || !u) return IBINN_ENULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return IBINN_EINVAL;
    AioP p{};
    p.x = y; p.w = w_inv; p.an = act_norm; p.off = act_offset; p.out = u;
    p.B = B; p.C = C; p.HW = H * W;
    p.hw_shift = aio_hw_shift(H * W, y, nullptr, u);
    size_t sm = aio_smem(C, H * W);
    int e = ibinn_set_smem(aio_mix_kernel<1>, sm);
    if (e) return e;
    IBINN_LAUNCH(aio_mix_kernel<1>, dim3(B), dim3(IBINN_THREADS), sm, (cudaStream_t)stream, p);
    return ibinn_launch_status();
}

extern "C" int ibinn_aio_affine_inverse(float* u, const float* a, const float* act_norm, float* jac, int B, int C, int H,
                                        int W, float clamp, float alpha, ibinn_stream_t stream) {
    IBINN_ENTRY_GUARD();
    if (!u || !act_norm || !jac || (C > 1 && !a)) return IBINN_ENULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return IBINN_EINVAL;
    AioP p{};
    p.out = u; p.a = a; p.an = act_norm; p.jac = jac; p.B = B; p.C = C; p.HW = H * W; p.clamp = clamp; p.alpha = alpha;
    IBINN_LAUNCH(aio_affine_inverse_kernel, dim3(B), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, p);
    return ibinn_launch_status();
}

static int aio_bwd_nsplit(int HW) { return (HW % 64 == 0) ? 2 : 1; }      // two CTAs per sample when each half keeps whole warps

extern "C" int ibinn_aio_backward_records(int B, int C, int H, int W) {
    (void)C;
    if (B <= 0 || H <= 0 || W <= 0) return IBINN_EINVAL;
    return B * aio_bwd_nsplit(H * W);
}

extern "C" int ibinn_aio_coupling_backward(const float* g_out, const float* g_jac, const float* x, const float* a,
                                           const float* out, const float* w, const float* act_norm,
                                           const float* act_offset, float* g_x, float* g_a, float* g_act_norm,
                                           float* g_act_offset, float* partials, uint32_t* counter, int B, int C, int H,
                                           int W, float clamp, float alpha, int defer_reduce, ibinn_stream_t stream) {
    IBINN_ENTRY_GUARD();
    if (!g_out || !x || !out || !w || !act_norm || !act_offset || !g_x || !partials || (!defer_reduce && !counter)) return IBINN_ENULL;
    if (C > 1 && (!a || !g_a)) return IBINN_ENULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return IBINN_EINVAL;
    AioP p{};
    p.g_out = g_out; p.g_jac = g_jac; p.x = x; p.a = a; p.out_saved = out; p.w = w; p.an = act_norm; p.off = act_offset;
    p.g_x = g_x; p.g_a = g_a; p.g_an = g_act_norm; p.g_off = g_act_offset; p.partials = partials; p.counter = counter;
    p.B = B; p.C = C; p.HW = H * W; p.clamp = clamp; p.alpha = alpha;
    p.nsplit = aio_bwd_nsplit(H * W);
    {
        const int HWl = H * W / p.nsplit;          // pixels per CTA
        if (HWl >= 2 && (HWl & (HWl - 1)) == 0)
            while ((1 << p.hw_shift) < HWl) ++p.hw_shift;
    }
    p.defer = defer_reduce ? 1 : 0;
    size_t sm = aio_smem(C, H * W / p.nsplit);
    int e = ibinn_set_smem(aio_bwd_kernel, sm);
    if (e) return e;
    IBINN_LAUNCH_PDL(aio_bwd_kernel, dim3(B * p.nsplit), dim3(IBINN_THREADS), sm, (cudaStream_t)stream, p);
    return ibinn_launch_status();
}

namespace {
// one block per table entry: thread <-> (record group, element); groups combined in fixed order
__global__ void __launch_bounds__(IBINN_THREADS) aio_reduce_many_kernel(const ibinn_aio_reduce_desc* __restrict__ table) {
    __shared__ float s_red[IBINN_THREADS];
    const ibinn_aio_reduce_desc d = table[blockIdx.x];
    const int T = 2 * d.C, tid = threadIdx.x;
    for (int e0 = 0; e0 < T; e0 += IBINN_THREADS) {              // T <= 256 in every shipped config: one pass
        const int Tl = min(T - e0, IBINN_THREADS);
        const int ngrp = IBINN_THREADS / Tl;
        const int e = tid % Tl, grp = tid / Tl;
        float t = 0.f;
        if (grp < ngrp) {
            constexpr int UB = 8;
            for (int r0 = grp; r0 < d.nrec; r0 += ngrp * UB) {
                float v[UB];
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    const int r = r0 + u * ngrp;
                    v[u] = r < d.nrec ? __ldg(d.partials + (size_t)r * T + e0 + e) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < UB; ++u) t += v[u];
            }
        }
        __syncthreads();
        s_red[tid] = t;
        __syncthreads();
        if (tid < Tl) {
            float tot = 0.f;
            for (int gi = 0; gi < ngrp; ++gi) tot += s_red[gi * Tl + tid];
            const int c = e0 + tid;
            if (c < d.C) { if (d.g_act_offset) d.g_act_offset[c] += tot; }
            else if (d.g_act_norm) d.g_act_norm[c - d.C] += tot;
        }
    }
}
}  // namespace

extern "C" int ibinn_aio_param_reduce_many(const ibinn_aio_reduce_desc* table, int n, ibinn_stream_t stream) {
    IBINN_ENTRY_GUARD();
    if (!table) return IBINN_ENULL;
    if (n <= 0 || n > 65535) return IBINN_EINVAL;
    IBINN_LAUNCH(aio_reduce_many_kernel, dim3(n), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, table);
    return ibinn_launch_status();
}

static int s2d_check(int B, int cp, int H, int W) {
    if (B <= 0 || cp <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return IBINN_EINVAL;
    return 0;
}

extern "C" int ibinn_s2d_coupling_forward(const float* x, int64_t x_bs, const float* a, float* out, int64_t out_bs,
                                          float* jac, int jac_accumulate, int B, int cp, int H, int W, float clamp,
                                          float alpha, ibinn_stream_t stream) {
    IBINN_ENTRY_GUARD();
    if (!x || !a || !out || !jac) return IBINN_ENULL;
    int e = s2d_check(B, cp, H, W);
    if (e) return e;
    S2dP p{};
    p.x = x; p.x_bs = x_bs; p.a = a; p.out = out; p.out_bs = out_bs; p.jac = jac; p.jac_acc = jac_accumulate;
    p.B = B; p.cp = cp; p.H = H; p.W = W; p.clamp = clamp; p.alpha = alpha;
    {
        const int Wo = W / 2, HWo = (H / 2) * Wo;
        if ((Wo & (Wo - 1)) == 0 && (HWo & (HWo - 1)) == 0 && HWo > 1) {
            while ((1 << p.wo_shift) < Wo) ++p.wo_shift;
            while ((1 << p.hwo_shift) < HWo) ++p.hwo_shift;
        }
    }
    IBINN_LAUNCH(s2d_kernel<0>, dim3(B), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, p);
    return ibinn_launch_status();
}

extern "C" int ibinn_s2d_coupling_inverse(const float* y, int64_t y_bs, const float* a, float* x, int64_t x_bs,
                                          float* jac, int jac_accumulate, int B, int cp, int H, int W, float clamp,
                                          float alpha, ibinn_stream_t stream) {
    IBINN_ENTRY_GUARD();
    if (!x || !a || !y || !jac) return IBINN_ENULL;
    int e = s2d_check(B, cp, H, W);
    if (e) return e;
    S2dP p{};
    p.x = x; p.x_bs = x_bs; p.a = a; p.out = const_cast<float*>(y); p.out_bs = y_bs; p.jac = jac; p.jac_acc = jac_accumulate;
    p.B = B; p.cp = cp; p.H = H; p.W = W; p.clamp = clamp; p.alpha = alpha;
    IBINN_LAUNCH(s2d_kernel<1>, dim3(B), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, p);
    return ibinn_launch_status();
}

extern "C" int ibinn_s2d_coupling_backward(const float* g_out, int64_t g_bs, const float* g_jac, const float* x,
                                           int64_t x_bs, const float* a, float* g_x, int64_t gx_bs, float* g_a, int B,
                                           int cp, int H, int W, float clamp, float alpha, ibinn_stream_t stream) {
    IBINN_ENTRY_GUARD();
    if (!g_out || !x || !a || !g_x || !g_a) return IBINN_ENULL;
    int e = s2d_check(B, cp, H, W);
    if (e) return e;
    S2dP p{};
    p.g_out = g_out; p.g_bs = g_bs; p.g_jac = g_jac; p.x = x; p.x_bs = x_bs; p.a = a; p.g_x = g_x; p.gx_bs = gx_bs; p.g_a = g_a;
    p.B = B; p.cp = cp; p.H = H; p.W = W; p.clamp = clamp; p.alpha = alpha;
    IBINN_LAUNCH(s2d_kernel<2>, dim3(B), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, p);
    return ibinn_launch_status();
}

extern "C" int ibinn_glow_affine_forward(const float* x, int64_t x_rs, const float* r, float* y, int64_t y_rs, float* jac,
                                         int jac_accumulate, int B, int L, float clamp, ibinn_stream_t stream) {
    IBINN_ENTRY_GUARD();
    if (!x || !r || !y || !jac) return IBINN_ENULL;
    if (B <= 0 || L <= 0) return IBINN_EINVAL;
    GlowP p{};
    p.x = x; p.x_rs = x_rs; p.r = r; p.y = y; p.y_rs = y_rs; p.jac = jac; p.jac_acc = jac_accumulate; p.B = B; p.L = L; p.clamp = clamp;
    IBINN_LAUNCH(glow_kernel<0>, dim3(B), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, p);
    return ibinn_launch_status();
}

extern "C" int ibinn_glow_affine_inverse(const float* y, int64_t y_rs, const float* r, float* x, int64_t x_rs, float* jac,
                                         int jac_accumulate, int B, int L, float clamp, ibinn_stream_t stream) {
    IBINN_ENTRY_GUARD();
    if (!x || !r || !y || !jac) return IBINN_ENULL;
    if (B <= 0 || L <= 0) return IBINN_EINVAL;
    GlowP p{};
    p.x = y; p.x_rs = y_rs; p.r = r; p.y = x; p.y_rs = x_rs; p.jac = jac; p.jac_acc = jac_accumulate; p.B = B; p.L = L; p.clamp = clamp;
    IBINN_LAUNCH(glow_kernel<1>, dim3(B), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, p);
    return ibinn_launch_status();
}

extern "C" int ibinn_glow_affine_backward(const float* g_y, int64_t gy_rs, const float* g_jac, const float* x, int64_t x_rs,
                                          const float* r, float* g_x, int64_t gx_rs, float* g_r, int B, int L, float clamp,
                                          ibinn_stream_t stream) {
    IBINN_ENTRY_GUARD();
    if (!g_y || !x || !r || !g_x || !g_r) return IBINN_ENULL;
    if (B <= 0 || L <= 0) return IBINN_EINVAL;
    GlowP p{};
    p.g_y = g_y; p.gy_rs = gy_rs; p.g_jac = g_jac; p.x = x; p.x_rs = x_rs; p.r = r; p.g_x = g_x; p.gx_rs = gx_rs; p.g_r = g_r;
    p.B = B; p.L = L; p.clamp = clamp;
    IBINN_LAUNCH(glow_kernel<2>, dim3(B), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, p);
    return ibinn_launch_status();
}
