// Subnet convolutions as tcgen05 implicit GEMMs (stride 1, 3x3 or 1x1; forward and dgrad), with
// fp32-accurate 3xTF32 arithmetic.  Same contract as the CUDA-core kernel in conv.cu
// (pre-op while staging, BatchNorm-statistics / mask epilogues); conv.cu dispatches here.
//
// GEMM view: M = pixels, N = Cout, K = Cin per filter tap, 9 (or 1) taps accumulated into one TMEM
// tile.  The whole batch is laid out as ONE padded 1-D pixel sequence: pitch P = W + pad, one shared
// zero row between images (pad = ksize/2).  A CTA owns 128 consecutive positions of that sequence;
// it stages positions [q0 - pad(P+1), q0 + 128 + pad(P+1)) ONCE (NCHW global -> pre-op -> hi/lo split
// -> K-major no-swizzle shared memory, rows 16 bytes apart) and expresses every tap (dy,dx) as the
// same tile viewed from row offset pad(P+1) + dy*P + dx: nine shifted descriptors, no im2col copy.
// Weights are pre-arranged per tap in the same canonical layout by a small prep kernel and streamed
// tap by tap with 1-D TMA bulk copies through a 3-deep mbarrier ring, so 64x64x9 filters (288 KB as
// hi+lo) never have to fit in shared memory at once.
// Warp roles: all 8 warps stage the activation tile; then warp 1 lane 0 = weight producer,
// warp 0 lane 0 = MMA issuer, all 8 warps = epilogue (warps w and w+4 split the columns of TMEM
// lanes 32(w%4)..+31).
#include "common.cuh"
#include "conv_internal.cuh"
#include "tc.cuh"

#ifndef IBINN_EMU
namespace {

constexpr int TM = 128;
constexpr int NST = 3;          // weight ring depth
constexpr int MAXC_TC = 256;

struct TcGeom {
    int pad, P, S, Qout, ntiles, npos, halo, cin_p, np, kk, tap_bytes, lbo_a, lbo_b, a_bytes, nst, tmem_cols;
    size_t smem_bytes;
};

static bool tc_geometry(const ibinn_conv_args& a, TcGeom& g) {
    if (a.stride != 1 || a.Cout > 64 || a.Cin > MAXC_TC) return false;
    if (a.Hout != a.Hin || a.Wout != a.Win) return false;
    g.pad = a.ksize / 2;
    g.kk = a.ksize * a.ksize;
    g.P = a.Win + g.pad;
    g.S = (a.Hin + g.pad) * g.P;
    g.Qout = a.B * g.S;
    g.ntiles = (g.Qout + TM - 1) / TM;
    g.halo = g.pad * (g.P + 1);
    g.npos = TM + 2 * g.halo;
    g.cin_p = (a.Cin + 7) & ~7;
    g.np = (a.Cout + 15) & ~15;
    g.lbo_a = g.npos * 16;
    g.lbo_b = g.np * 16;
    g.a_bytes = (g.cin_p / 4) * g.lbo_a;              // one of hi / lo
    g.tap_bytes = 2 * (g.cin_p / 4) * g.lbo_b;        // hi + lo of one tap
    g.nst = g.kk < NST ? g.kk : NST;
    g.tmem_cols = g.np <= 32 ? 32 : 64;
    size_t b = 2 * (size_t)g.a_bytes + (size_t)g.nst * g.tap_bytes + (size_t)g.npos * 4 /*pos table*/ +
               (3 * MAXC_TC + 4 * 64) * 4 /*coef tables*/ + (2 * 4 * 64 + 16) * 4 /*reduction*/ + 256;
    g.smem_bytes = b;
    if (g.lbo_a >= (1 << 18) || b > 220 * 1024) return false;
    return true;
}

// ---- weight preparation: raw filters -> per tap [hi|lo][k4][n][4] (canonical K-major rows of 16 bytes)
__global__ void __launch_bounds__(IBINN_THREADS) wprep_kernel(const float* __restrict__ w, float* __restrict__ out, int Cin, int Cout,
                                                             int cin_p, int np, int kk, int transposed) {
    const int per_tap = (cin_p / 4) * np * 4;
    const int total = kk * per_tap;
    for (int e = blockIdx.x * IBINN_THREADS + threadIdx.x; e < total; e += gridDim.x * IBINN_THREADS) {
        const int i = e & 3, n = (e >> 2) % np, k4 = (e >> 2) / np % (cin_p / 4), tap = e / per_tap;
        const int ci = k4 * 4 + i;
        float v = 0.f;
        if (n < Cout && ci < Cin)
            v = transposed ? __ldg(w + ((size_t)ci * Cout + n) * kk + (kk - 1 - tap)) : __ldg(w + ((size_t)n * Cin + ci) * kk + tap);
        float hi, lo;
        tc::split_tf32(v, hi, lo);
        const size_t o = (size_t)tap * 2 * per_tap + (size_t)(k4 * np + n) * 4 + i;
        out[o] = hi;
        out[o + per_tap] = lo;
    }
}

__device__ __forceinline__ float tc_invstd(const ibinn_bn_ref& r, int c) {
    const float s = r.scale[c];
    return r.scale_is_variance ? rsqrtf(s + r.eps) : s;
}

__global__ void __launch_bounds__(IBINN_THREADS, 1) conv_tc_kernel(const ibinn_conv_args p, const TcGeom g) {
    extern __shared__ __align__(128) unsigned char tcs[];
    __shared__ __align__(8) uint64_t bar_full[NST];
    __shared__ __align__(8) uint64_t bar_empty[NST];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_flag;
    __shared__ float s_cnt[4];

    unsigned char* a_hi = tcs;
    unsigned char* a_lo = a_hi + g.a_bytes;
    unsigned char* wst = a_lo + g.a_bytes;                                  // nst stages of tap_bytes
    int* s_pos = reinterpret_cast<int*>(wst + (size_t)g.nst * g.tap_bytes);   // element offset of (b, c=0, y, x) or -1
    float* s_pa = reinterpret_cast<float*>(s_pos + g.npos);
    float* s_pb = s_pa + MAXC_TC;
    float* s_pc = s_pb + MAXC_TC;
    float* s_ea = s_pc + MAXC_TC;
    float* s_eb = s_ea + 64;
    float* s_em = s_eb + 64;
    float* s_ei = s_em + 64;
    float* s_r1 = s_ei + 64;        // [4][64]
    float* s_r2 = s_r1 + 4 * 64;    // [4][64]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * TM;
    const int HsWs = p.Hs * p.Ws;

    if (tid == 0) {
        for (int i = 0; i < g.nst; ++i) { tc::mbar_init(&bar_full[i], 1); tc::mbar_init(&bar_empty[i], 1); }
        tc::mbar_init(&bar_done, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, g.tmem_cols);

    // ---- coefficient tables (same algebra as conv.cu)
    if (p.pre == IBINN_PRE_BN_RELU) {
        for (int c = tid; c < p.Cin; c += IBINN_THREADS) {
            const float is = tc_invstd(p.pre_bn, c), A = p.pre_bn.gamma[c] * is;
            s_pa[c] = A;
            s_pb[c] = p.pre_bn.beta[c] - p.pre_bn.mean[c] * A;
        }
    } else if (p.pre == IBINN_PRE_BNBWD) {
        for (int c = tid; c < p.Cin; c += IBINN_THREADS) {
            const float is = tc_invstd(p.pre_bn, c), k = p.pre_bn.gamma[c] * is;
            const float m1 = p.pre_sums[c] * p.pre_inv_count, m2 = p.pre_sums[p.Cin + c] * p.pre_inv_count;
            s_pa[c] = k;
            s_pb[c] = -k * m2 * is;
            s_pc[c] = -k * m1 + k * m2 * is * p.pre_bn.mean[c];
        }
    }
    if (p.epi == IBINN_EPI_MASKSTATS) {
        for (int c = tid; c < 64; c += IBINN_THREADS) {
            float A = 0.f, Bc = 0.f, mean = 0.f, is = 0.f;
            if (c < p.Cout) {
                is = tc_invstd(p.m_bn, c);
                mean = p.m_bn.mean[c];
                A = p.m_bn.gamma[c] * is;
                Bc = p.m_bn.beta[c] - mean * A;
            }
            s_ea[c] = A; s_eb[c] = Bc; s_em[c] = mean; s_ei[c] = is;
        }
    }
    // ---- position table of the staged window
    for (int i = tid; i < g.npos; i += IBINN_THREADS) {
        const int q = q0 - g.halo + i;
        int off = -1;
        if (q >= 0 && q < g.Qout) {
            const int b = q / g.S, rem = q - b * g.S;
            const int yy = rem / g.P - g.pad, xx = rem % g.P;
            if (yy >= 0 && yy < p.Hin && xx < p.Win) {
                if (p.upsample2) {
                    if (!((yy | xx) & 1) && (yy >> 1) < p.Hs && (xx >> 1) < p.Ws) off = b * (int)p.x_bs + (yy >> 1) * p.Ws + (xx >> 1);
                } else off = b * (int)p.x_bs + yy * p.Ws + xx;
            }
        }
        s_pos[i] = off;
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;

    // ---- weight producer starts streaming right away (overlaps the activation staging)
    if (warp == 1 && lane == 0) {
        for (int t = 0; t < g.nst; ++t) {
            tc::mbar_arrive_expect_tx(&bar_full[t], g.tap_bytes);
            tc::bulk_g2s(wst + (size_t)t * g.tap_bytes, reinterpret_cast<const unsigned char*>(p.wprep) + (size_t)t * g.tap_bytes, g.tap_bytes,
                         &bar_full[t]);
        }
    }

    // ---- stage the activation window: 16-byte chunk = 4 channels of one position
    {
        const int nk4 = g.cin_p / 4;
        const int b_x2 = (int)p.x2_bs, b_x = (int)p.x_bs;
        for (int c = tid; c < g.npos * nk4; c += IBINN_THREADS) {
            const int i = c % g.npos, k4 = c / g.npos;
            const int off = s_pos[i];
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (off >= 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ch = k4 * 4 + j;
                    if (ch < p.Cin) {
                        const size_t o = (size_t)off + (size_t)ch * HsWs;
                        float t = __ldg(p.x + o);
                        if (p.pre == IBINN_PRE_RELU) t = fmaxf(t, 0.f);
                        else if (p.pre == IBINN_PRE_BN_RELU) {
                            t = fmaxf(fmaf(t, s_pa[ch], s_pb[ch]), 0.f);
                            if (p.pre_mask) t *= __ldg(p.pre_mask + (size_t)(off / b_x) * p.Cin + ch);
                        } else if (p.pre == IBINN_PRE_BNBWD) {
                            // same (b, y, x) in x2, whose batch stride may differ
                            const int bb = off / b_x, in_img = off - bb * b_x;
                            t = fmaf(s_pa[ch], t, fmaf(s_pb[ch], __ldg(p.x2 + (size_t)bb * b_x2 + in_img + (size_t)ch * HsWs), s_pc[ch]));
                        }
                        v[j] = t;
                    }
                }
            }
            float h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) tc::split_tf32(v[j], h[j], l[j]);
            *reinterpret_cast<float4*>(a_hi + (size_t)k4 * g.lbo_a + i * 16) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(a_lo + (size_t)k4 * g.lbo_a + i * 16) = make_float4(l[0], l[1], l[2], l[3]);
        }
    }
    tc::fence_async_smem();
    __syncthreads();

    // ---- main loop over the filter taps
    if (warp == 1 && lane == 0) {
        for (int t = g.nst; t < g.kk; ++t) {
            const int s = t % g.nst;
            tc::mbar_wait(&bar_empty[s], ((t / g.nst) - 1) & 1);
            tc::mbar_arrive_expect_tx(&bar_full[s], g.tap_bytes);
            tc::bulk_g2s(wst + (size_t)s * g.tap_bytes, reinterpret_cast<const unsigned char*>(p.wprep) + (size_t)t * g.tap_bytes, g.tap_bytes,
                         &bar_full[s]);
        }
    } else if (warp == 0 && lane == 0) {
        tc::fence_after_sync();
        const uint32_t idesc = tc::idesc_tf32(TM, g.np);
        const uint32_t ah = tc::smem_u32(a_hi), al = tc::smem_u32(a_lo);
        const int nks = g.cin_p / 8;
        const uint32_t b_lo_off = (uint32_t)(g.cin_p / 4) * g.lbo_b;
        for (int t = 0; t < g.kk; ++t) {
            const int s = t % g.nst;
            tc::mbar_wait(&bar_full[s], (t / g.nst) & 1);
            tc::fence_after_sync();
            const int dy = g.pad ? t / 3 - 1 : 0, dx = g.pad ? t % 3 - 1 : 0;
            const uint32_t row_off = (uint32_t)(g.halo + dy * g.P + dx) * 16;
            const uint32_t bh = tc::smem_u32(wst + (size_t)s * g.tap_bytes), bl = bh + b_lo_off;
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t aa = (pass == 2 ? al : ah) + row_off, bb = pass == 1 ? bl : bh;
                for (int ks = 0; ks < nks; ++ks) {
                    const uint64_t da = tc::smem_desc(aa + ks * 2 * g.lbo_a, g.lbo_a);
                    const uint64_t db = tc::smem_desc(bb + ks * 2 * g.lbo_b, g.lbo_b);
                    tc::mma_tf32(tmem_d, da, db, idesc, (t | pass | ks) != 0);
                }
            }
            tc::mma_commit(&bar_empty[s]);
        }
        tc::mma_commit(&bar_done);
    }
    __syncwarp();
    tc::mbar_wait(&bar_done, 0);
    tc::fence_after_sync();

    // ---- epilogue: thread <-> output position q0 + 32*(warp%4) + lane, columns split between warp and warp+4
    const int wq = warp & 3, half = warp >> 2;
    const int ncol = g.np / 2;                       // 8, 16, 24 or 32 columns per thread
    const int col0 = half * ncol;
    const int q = q0 + wq * 32 + lane;
    bool valid = false;
    size_t obase = 0, mbase = 0, abase = 0;
    int bimg = 0;
    if (q < g.Qout) {
        const int b = q / g.S, rem = q - b * g.S;
        const int yy = rem / g.P - g.pad, xx = rem % g.P;
        if (yy >= 0 && yy < p.Hout && xx < p.Wout) {
            valid = true;
            bimg = b;
            const size_t in_img = (size_t)yy * p.Wout + xx;
            obase = (size_t)b * p.y_bs + in_img;
            mbase = (size_t)b * p.m_bs + in_img;
            abase = (size_t)b * p.addend_bs + in_img;
        }
    }
    const size_t HWo = (size_t)p.Hout * p.Wout;
    const uint32_t taddr = tmem_d + ((uint32_t)(wq * 32) << 16) + col0;
    const bool stats = (p.epi == IBINN_EPI_BNSTATS || p.epi == IBINN_EPI_MASKSTATS);

    float vals[32];
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
        if (c8 * 8 < ncol) {
            float v[8];
            tc::tmem_ld8(taddr + c8 * 8, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) vals[c8 * 8 + i] = v[i];
        }
    }
    // element-wise part + store
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j >= ncol) continue;
        const int c = col0 + j;
        float v = vals[j], xh = 0.f;
        if (valid && c < p.Cout) {
            if (p.bias) v += __ldg(p.bias + c);
            if (p.epi == IBINN_EPI_MASKSTATS) {
                const float h = __ldg(p.m_h + mbase + (size_t)c * HWo);
                v = fmaf(h, s_ea[c], s_eb[c]) > 0.f ? v : 0.f;
                if (p.m_drop) v *= __ldg(p.m_drop + (size_t)bimg * p.Cout + c);
                xh = (h - s_em[c]) * s_ei[c];
            } else if (p.epi == IBINN_EPI_RELUMASK) {
                v = __ldg(p.m_h + mbase + (size_t)c * HWo) > 0.f ? v : 0.f;
            }
            float o = v;
            if (p.addend) o += p.addend[abase + (size_t)c * HWo];
            p.y[obase + (size_t)c * HWo] = o;
        } else v = 0.f;
        vals[j] = v;
        if (p.epi == IBINN_EPI_MASKSTATS) {
            // per-channel sums of v and v*xhat over the tile
            const float a1 = warp_sum(v), a2 = warp_sum(v * xh);
            if (lane == 0) { s_r1[wq * 64 + c] = a1; s_r2[wq * 64 + c] = a2; }
        } else if (p.epi == IBINN_EPI_BNSTATS) {
            const float a1 = warp_sum(v);
            if (lane == 0) s_r1[wq * 64 + c] = a1;
        }
    }
    if (stats) {
        const float cw = warp_sum(valid ? 1.f : 0.f);
        if (half == 0 && lane == 0) s_cnt[wq] = cw;
        __syncthreads();
        const float n_tile = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
        float* rec = p.partials + (size_t)blockIdx.x * (2 * p.Cout + 4);
        if (p.epi == IBINN_EPI_BNSTATS) {
            // second pass: centred sum of squares around the tile mean
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j >= ncol) continue;
                const int c = col0 + j;
                const float mean = n_tile > 0.f ? (s_r1[c] + s_r1[64 + c] + s_r1[128 + c] + s_r1[192 + c]) / n_tile : 0.f;
                const float d = (valid && c < p.Cout) ? vals[j] - mean : 0.f;
                const float a2 = warp_sum(d * d);
                if (lane == 0) s_r2[wq * 64 + c] = a2;
            }
            __syncthreads();
            if (tid < p.Cout) {
                const float s = s_r1[tid] + s_r1[64 + tid] + s_r1[128 + tid] + s_r1[192 + tid];
                rec[tid] = n_tile > 0.f ? s / n_tile : 0.f;
                rec[p.Cout + tid] = s_r2[tid] + s_r2[64 + tid] + s_r2[128 + tid] + s_r2[192 + tid];
            }
        } else {
            if (tid < p.Cout) {
                rec[tid] = s_r1[tid] + s_r1[64 + tid] + s_r1[128 + tid] + s_r1[192 + tid];
                rec[p.Cout + tid] = s_r2[tid] + s_r2[64 + tid] + s_r2[128 + tid] + s_r2[192 + tid];
            }
        }
        if (tid == 0) rec[2 * p.Cout] = n_tile;
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, g.tmem_cols);
    if (!stats) return;

    // ---- last CTA: combine the per-tile records in tile order (deterministic), like conv.cu
    const unsigned nblk = gridDim.x;
    if (!last_block_ticket(p.counter, nblk, &s_flag)) return;
    int cpad = 8;
    while (cpad < p.Cout) cpad <<= 1;
    const int tpc = IBINN_THREADS / cpad > 32 ? 32 : IBINN_THREADS / cpad;
    const int c = tid / tpc, j = tid % tpc;
    const bool cval = (c < p.Cout) && (tid < cpad * tpc);
    const size_t rs = 2 * p.Cout + 4;
    if (p.epi == IBINN_EPI_BNSTATS) {
        double n = 0., mean = 0., m2 = 0.;
        if (cval) {
            for (unsigned bb = j; bb < nblk; bb += tpc) {
                const double nb = p.partials[bb * rs + 2 * p.Cout];
                if (nb <= 0.) continue;
                const double mb = p.partials[bb * rs + c], qb = p.partials[bb * rs + p.Cout + c];
                const double nn = n + nb, d = mb - mean;
                mean += d * nb / nn;
                m2 += qb + d * d * n * nb / nn;
                n = nn;
            }
        }
        for (int o = tpc >> 1; o > 0; o >>= 1) {
            const double n2 = __shfl_xor_sync(0xffffffffu, n, o);
            const double mean2 = __shfl_xor_sync(0xffffffffu, mean, o);
            const double q2 = __shfl_xor_sync(0xffffffffu, m2, o);
            const double nn = n + n2;
            if (nn > 0.) {
                const double d = mean2 - mean;
                const double nm = (mean * n + mean2 * n2) / nn;
                m2 = m2 + q2 + d * d * n * n2 / nn;
                mean = nm;
                n = nn;
            }
        }
        long long nbt = 0;
        if (p.num_batches_tracked) nbt = *p.num_batches_tracked;
        __syncthreads();
        if (cval && j == 0) {
            const double var_b = m2 / n;
            p.save_mean[c] = (float)mean;
            p.save_invstd[c] = (float)(1.0 / sqrt(var_b + (double)p.bn_eps));
            if (p.running_mean) {
                const double f = p.momentum >= 0.f ? (double)p.momentum : 1.0 / (double)(nbt + 1);
                const double var_u = n > 1. ? m2 / (n - 1.) : var_b;
                p.running_mean[c] = (float)((1.0 - f) * p.running_mean[c] + f * mean);
                p.running_var[c] = (float)((1.0 - f) * p.running_var[c] + f * var_u);
            }
        }
        if (tid == 0 && p.num_batches_tracked) *p.num_batches_tracked = nbt + 1;
    } else {
        double a1 = 0., a2 = 0.;
        if (cval) {
            for (unsigned bb = j; bb < nblk; bb += tpc) {
                a1 += (double)p.partials[bb * rs + c];
                a2 += (double)p.partials[bb * rs + p.Cout + c];
            }
        }
        for (int o = tpc >> 1; o > 0; o >>= 1) {
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (cval && j == 0) {
            p.sums_out[c] = (float)a1;
            p.sums_out[p.Cout + c] = (float)a2;
            if (p.m_dbeta) p.m_dbeta[c] += (float)a1;
            if (p.m_dgamma) p.m_dgamma[c] += (float)a2;
        }
    }
    if (tid == 0) *p.counter = 0u;
}

}  // namespace

bool conv_tc_supported(const ibinn_conv_args& a) {
    TcGeom g;
    if (!tc_geometry(a, g)) return false;
    // 32-bit position offsets
    if ((long long)a.B * a.x_bs >= (1ll << 31)) return false;
    return true;
}

int64_t conv_tc_partials_floats(const ibinn_conv_args& a) {
    TcGeom g;
    if (!tc_geometry(a, g)) return IBINN_EINVAL;
    return (int64_t)g.ntiles * (2 * a.Cout + 4);
}

int64_t conv_tc_wprep_floats(const ibinn_conv_args& a) {
    TcGeom g;
    if (!tc_geometry(a, g)) return IBINN_EINVAL;
    return (int64_t)g.kk * g.tap_bytes / 4;
}

int conv_tc_launch(const ibinn_conv_args& a, cudaStream_t st) {
    TcGeom g;
    if (!tc_geometry(a, g)) return IBINN_EINVAL;
    if (!a.wprep) return IBINN_ENULL;
    if (((uintptr_t)a.wprep & 15) != 0) return IBINN_EALIGN;
    const int total = g.kk * (g.cin_p / 4) * g.np * 4;
    int nb = (total + IBINN_THREADS - 1) / IBINN_THREADS;
    if (nb > 148) nb = 148;
    IBINN_LAUNCH(wprep_kernel, dim3(nb), dim3(IBINN_THREADS), 0, st, a.w, a.wprep, a.Cin, a.Cout, g.cin_p, g.np, g.kk, a.w_transposed);
    int e = ibinn_launch_status();
    if (e) return e;
    e = ibinn_set_smem(conv_tc_kernel, g.smem_bytes);
    if (e) return e;
    IBINN_LAUNCH(conv_tc_kernel, dim3(g.ntiles), dim3(IBINN_THREADS), g.smem_bytes, st, a, g);
    return ibinn_launch_status();
}

#else  // IBINN_EMU: the simulator build has no tensor cores; conv.cu keeps everything on its CUDA-core kernel
bool conv_tc_supported(const ibinn_conv_args&) { return false; }
int64_t conv_tc_partials_floats(const ibinn_conv_args&) { return 0; }
int64_t conv_tc_wprep_floats(const ibinn_conv_args&) { return 0; }
int conv_tc_launch(const ibinn_conv_args&, cudaStream_t) { return IBINN_EINVAL; }
#endif
