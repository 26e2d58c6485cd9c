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
#define DBG(i)

namespace {

constexpr int TM = 128;
constexpr int NST = 9;          // weight ring depth (as many taps in flight as shared memory allows)
constexpr int MAXC_TC = 256;

struct TcGeom {
    int pad, P, S, Qout, ntiles, npos, halo, cin_p, np, kk, tap_bytes, lbo_a, lbo_b, a_bytes, nst, tmem_cols, op_bytes;
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
    {   // ring depth: every tap resident if it fits next to the activation tile, else what ~200 KB allows (>= 2)
        const long long room = 200 * 1024 - 2LL * g.a_bytes - 8 * 1024;
        long long n = room / g.tap_bytes;
        if (n > g.kk) n = g.kk;
        if (n > NST) n = NST;
        if (n < 2 && g.kk > 1) n = 2;
        if (n < 1) n = 1;
        g.nst = (int)n;
    }
    g.tmem_cols = g.np <= 32 ? 32 : 64;
    size_t operands = 2 * (size_t)g.a_bytes + (size_t)g.nst * g.tap_bytes;
    const size_t stat_tiles = 2 * 64 * (TM + 4) * 4;             // reuse of the operand area by the statistics epilogue
    if (operands < stat_tiles) {                                 // grow the weight ring region so the tiles fit
        operands = stat_tiles;
    }
    g.op_bytes = (int)operands;
    size_t b = operands + (size_t)g.npos * 4 /*pos table*/ + (3 * MAXC_TC + 4 * 64) * 4 /*coef tables*/ + TM * 4 + 256;
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

template <int PRE, int EPI>
__global__ void __launch_bounds__(IBINN_THREADS, 1) conv_tc_kernel(const ibinn_conv_args p, const TcGeom g) {
    extern __shared__ __align__(128) unsigned char tcs[];
    __shared__ __align__(8) uint64_t bar_full[NST];
    __shared__ __align__(8) uint64_t bar_empty[NST];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_flag;

    unsigned char* a_hi = tcs;
    unsigned char* a_lo = a_hi + g.a_bytes;
    unsigned char* wst = a_lo + g.a_bytes;                                  // nst stages of tap_bytes
    int* s_pos = reinterpret_cast<int*>(tcs + g.op_bytes);                   // element offset of (b, c=0, y, x) or -1
    float* s_pa = reinterpret_cast<float*>(s_pos + g.npos);
    float* s_pb = s_pa + MAXC_TC;
    float* s_pc = s_pb + MAXC_TC;
    float* s_ea = s_pc + MAXC_TC;
    float* s_eb = s_ea + 64;
    float* s_em = s_eb + 64;
    float* s_ei = s_em + 64;
    float* s_valid = s_ei + 64;     // [128] 1/0 per output row of the tile
    // after the MMAs the operand area is free: two [64][TP] transposition tiles for the statistics
    constexpr int TP = TM + 4;
    float* s_t1 = reinterpret_cast<float*>(tcs);
    float* s_t2 = s_t1 + 64 * TP;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * TM;
    const int HsWs = p.Hs * p.Ws;
    DBG(0);

    if (tid == 0) {
        for (int i = 0; i < g.nst; ++i) { tc::mbar_init(&bar_full[i], 1); tc::mbar_init(&bar_empty[i], 1); }
        tc::mbar_init(&bar_done, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, g.tmem_cols);

    // ---- coefficient tables (same algebra as conv.cu)
    if (PRE == IBINN_PRE_BN_RELU) {
        for (int c = tid; c < p.Cin; c += IBINN_THREADS) {
            const float is = tc_invstd(p.pre_bn, c), A = p.pre_bn.gamma[c] * is;
            s_pa[c] = A;
            s_pb[c] = p.pre_bn.beta[c] - p.pre_bn.mean[c] * A;
        }
    } else if (PRE == IBINN_PRE_BNBWD) {
        for (int c = tid; c < p.Cin; c += IBINN_THREADS) {
            const float is = tc_invstd(p.pre_bn, c), k = p.pre_bn.gamma[c] * is;
            const float m1 = p.pre_sums[c] * p.pre_inv_count, m2 = p.pre_sums[p.Cin + c] * p.pre_inv_count;
            s_pa[c] = k;
            s_pb[c] = -k * m2 * is;
            s_pc[c] = -k * m1 + k * m2 * is * p.pre_bn.mean[c];
        }
    }
    if (EPI == IBINN_EPI_MASKSTATS) {
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
    DBG(1);

    // ---- weight producer starts streaming right away (overlaps the activation staging)
    const bool resident = (g.nst == g.kk);          // every tap fits: one barrier, no ring traffic
    if (warp == 1 && lane == 0) {
        if (resident) {
            tc::mbar_arrive_expect_tx(&bar_full[0], (uint32_t)g.kk * g.tap_bytes);
            for (int t = 0; t < g.kk; ++t)
                tc::bulk_g2s(wst + (size_t)t * g.tap_bytes, reinterpret_cast<const unsigned char*>(p.wprep) + (size_t)t * g.tap_bytes,
                             g.tap_bytes, &bar_full[0]);
        } else {
            for (int t = 0; t < g.nst; ++t) {
                tc::mbar_arrive_expect_tx(&bar_full[t], g.tap_bytes);
                tc::bulk_g2s(wst + (size_t)t * g.tap_bytes, reinterpret_cast<const unsigned char*>(p.wprep) + (size_t)t * g.tap_bytes,
                             g.tap_bytes, &bar_full[t]);
            }
        }
    }

    // ---- stage the activation window: 16-byte chunk = 4 channels of one position.  Tasks are taken
    // four at a time with all global loads issued before the first use (the shared-memory stores would
    // otherwise serialise the loads behind them).
    {
        const int nk4 = g.cin_p / 4;
        const int ntask = g.npos * nk4;
        const int b_x2 = (int)p.x2_bs, b_x = (int)p.x_bs;
        constexpr int U = 4;
        int i_u = tid, k_u = 0;                      // task (position, chunk) of this thread, advanced by 256 each time
        while (i_u >= g.npos) { i_u -= g.npos; ++k_u; }
        for (int c0 = tid; c0 < ntask; c0 += IBINN_THREADS * U) {
            float v[U][4], w2[U][4];
            int ti[U], tk[U], toff[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                ti[u] = i_u; tk[u] = k_u;
                const bool live = (c0 + u * IBINN_THREADS) < ntask;
                toff[u] = live ? s_pos[i_u] : -2;
                i_u += IBINN_THREADS;
                while (i_u >= g.npos) { i_u -= g.npos; ++k_u; }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ch = tk[u] * 4 + j;
                    v[u][j] = 0.f; w2[u][j] = 0.f;
                    if (toff[u] >= 0 && ch < p.Cin) {
                        const size_t o = (size_t)toff[u] + (size_t)ch * HsWs;
                        v[u][j] = __ldg(p.x + o);
                        if (PRE == IBINN_PRE_BNBWD) {
                            const int bb = toff[u] / b_x, in_img = toff[u] - bb * b_x;     // same (b, y, x) in x2, own batch stride
                            w2[u][j] = __ldg(p.x2 + (size_t)bb * b_x2 + in_img + (size_t)ch * HsWs);
                        } else if (PRE == IBINN_PRE_BN_RELU) {
                            if (p.pre_mask) w2[u][j] = __ldg(p.pre_mask + (size_t)(toff[u] / b_x) * p.Cin + ch);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (toff[u] == -2) continue;
                float h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ch = tk[u] * 4 + j;
                    float t = v[u][j];
                    if (toff[u] >= 0 && ch < p.Cin) {
                        if (PRE == IBINN_PRE_RELU) t = fmaxf(t, 0.f);
                        else if (PRE == IBINN_PRE_BN_RELU) {
                            t = fmaxf(fmaf(t, s_pa[ch], s_pb[ch]), 0.f);
                            if (p.pre_mask) t *= w2[u][j];
                        } else if (PRE == IBINN_PRE_BNBWD) t = fmaf(s_pa[ch], t, fmaf(s_pb[ch], w2[u][j], s_pc[ch]));
                    }
                    tc::split_tf32(t, h[j], l[j]);
                }
                *reinterpret_cast<float4*>(a_hi + (size_t)tk[u] * g.lbo_a + ti[u] * 16) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(a_lo + (size_t)tk[u] * g.lbo_a + ti[u] * 16) = make_float4(l[0], l[1], l[2], l[3]);
            }
        }
    }
    tc::fence_async_smem();
    __syncthreads();
    DBG(2);

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
        // descriptors differ only in the 14-bit start-address field: build once, then add (bytes >> 4)
        const uint64_t da_hi = tc::smem_desc(tc::smem_u32(a_hi), g.lbo_a), da_lo = tc::smem_desc(tc::smem_u32(a_lo), g.lbo_a);
        const uint64_t db0 = tc::smem_desc(tc::smem_u32(wst), g.lbo_b);
        const int nks = g.cin_p / 8;
        const uint32_t b_lo_off = ((uint32_t)(g.cin_p / 4) * g.lbo_b) >> 4;
        const uint32_t a_step = (uint32_t)(2 * g.lbo_a) >> 4, b_step = (uint32_t)(2 * g.lbo_b) >> 4, st_step = (uint32_t)g.tap_bytes >> 4;
        bool acc = false;
        for (int t = 0; t < g.kk; ++t) {
            const int s = t % g.nst;
            if (!resident || t == 0) {
                tc::mbar_wait(&bar_full[s], (t / g.nst) & 1);
                tc::fence_after_sync();
            }
            const int dy = g.pad ? t / 3 - 1 : 0, dx = g.pad ? t % 3 - 1 : 0;
            const uint32_t row_off = (uint32_t)(g.halo + dy * g.P + dx);          // 16-byte rows
            const uint64_t bh = db0 + (uint64_t)(s * st_step), bl = bh + b_lo_off;
#pragma unroll 1
            for (int pass = 0; pass < 3; ++pass) {
                uint64_t da = (pass == 2 ? da_lo : da_hi) + row_off;
                uint64_t db = pass == 1 ? bl : bh;
#pragma unroll 2
                for (int ks = 0; ks < nks; ++ks) {
                    tc::mma_tf32(tmem_d, da, db, idesc, acc);
                    acc = true;
                    da += a_step;
                    db += b_step;
                }
            }
            if (!resident) tc::mma_commit(&bar_empty[s]);
        }
        tc::mma_commit(&bar_done);
        DBG(3);
    }
    __syncwarp();
    tc::mbar_wait(&bar_done, 0);
    tc::fence_after_sync();
    DBG(4);

    // ---- epilogue: thread <-> output position q0 + 32*(warp%4) + lane, columns split between warp and warp+4
    const int wq = warp & 3, half = warp >> 2;
    const int ncol = g.np / 2;                       // 8, 16, 24 or 32 columns per thread
    const int col0 = half * ncol;
    const int row = wq * 32 + lane;
    const int q = q0 + row;
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
    constexpr bool STATS = (EPI == IBINN_EPI_BNSTATS || EPI == IBINN_EPI_MASKSTATS);
    if (STATS && half == 0) s_valid[row] = valid ? 1.f : 0.f;

#pragma unroll 1
    for (int c8 = 0; c8 < ncol; c8 += 8) {
        float v[8];
        tc::tmem_ld8(taddr + c8, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = col0 + c8 + i;
            float o = 0.f, xh = 0.f;
            if (valid && c < p.Cout) {
                o = v[i];
                if (EPI == IBINN_EPI_STORE || EPI == IBINN_EPI_RELUMASK) { if (p.bias) o += __ldg(p.bias + c); }
                if (EPI == IBINN_EPI_MASKSTATS) {
                    const float h = __ldg(p.m_h + mbase + (size_t)c * HWo);
                    o = fmaf(h, s_ea[c], s_eb[c]) > 0.f ? o : 0.f;
                    if (p.m_drop) o *= __ldg(p.m_drop + (size_t)bimg * p.Cout + c);
                    xh = (h - s_em[c]) * s_ei[c];
                } else if (EPI == IBINN_EPI_RELUMASK) {
                    o = __ldg(p.m_h + mbase + (size_t)c * HWo) > 0.f ? o : 0.f;
                }
                float st = o;
                if (EPI == IBINN_EPI_STORE || EPI == IBINN_EPI_RELUMASK) { if (p.addend) st += p.addend[abase + (size_t)c * HWo]; }
                p.y[obase + (size_t)c * HWo] = st;
            }
            if (STATS) {
                s_t1[c * TP + row] = o;
                if (EPI == IBINN_EPI_MASKSTATS) s_t2[c * TP + row] = o * xh;
            }
        }
    }
    DBG(8);
    if (STATS) {
        __syncthreads();
        // thread (c = tid/4, part = tid%4) reduces 32 rows of channel c; the 4 parts meet by shuffle
        const int c = tid >> 2, part = tid & 3;
        float a1 = 0.f, a2 = 0.f, cnt = 0.f;
        const float* t1 = s_t1 + c * TP + part * 32;
        const float* t2 = s_t2 + c * TP + part * 32;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            a1 += t1[r];
            cnt += s_valid[part * 32 + r];
            if (EPI == IBINN_EPI_MASKSTATS) a2 += t2[r];
        }
        a1 += __shfl_xor_sync(0xffffffffu, a1, 1); a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 1); cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
        float* rec = p.partials + (size_t)blockIdx.x * (2 * p.Cout + 4);
        if (EPI == IBINN_EPI_BNSTATS) {
            const float mean = cnt > 0.f ? a1 / cnt : 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                const float d = t1[r] - mean;
                a2 = fmaf(s_valid[part * 32 + r] * d, d, a2);
            }
            a2 += __shfl_xor_sync(0xffffffffu, a2, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
            if (part == 0 && c < p.Cout) { rec[c] = mean; rec[p.Cout + c] = a2; }
        } else {
            a2 += __shfl_xor_sync(0xffffffffu, a2, 1); a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
            if (part == 0 && c < p.Cout) { rec[c] = a1; rec[p.Cout + c] = a2; }
        }
        if (tid == 0) rec[2 * p.Cout] = cnt;
    }

    DBG(5);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, g.tmem_cols);
    DBG(6);
    if (!STATS) return;

    // ---- last CTA: combine the per-tile records in tile order (deterministic), like conv.cu
    const unsigned nblk = gridDim.x;
    if (!last_block_ticket(p.counter, nblk, &s_flag)) return;
    int cpad = 8;
    while (cpad < p.Cout) cpad <<= 1;
    const int tpc = IBINN_THREADS / cpad > 32 ? 32 : IBINN_THREADS / cpad;
    const int c = tid / tpc, j = tid % tpc;
    const bool cval = (c < p.Cout) && (tid < cpad * tpc);
    const size_t rs = 2 * p.Cout + 4;
    if (EPI == IBINN_EPI_BNSTATS) {
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

template <int PRE, int EPI>
static int tc_launch_inst(const ibinn_conv_args& a, const TcGeom& g, cudaStream_t st) {
    auto kfn = conv_tc_kernel<PRE, EPI>;
    int e = ibinn_set_smem(kfn, g.smem_bytes);
    if (e) return e;
    IBINN_LAUNCH(kfn, dim3(g.ntiles), dim3(IBINN_THREADS), g.smem_bytes, st, a, g);
    return ibinn_launch_status();
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
#define TC_CASE(PRE_, EPI_) if (a.pre == PRE_ && a.epi == EPI_) return tc_launch_inst<PRE_, EPI_>(a, g, st);
    TC_CASE(0, 0) TC_CASE(0, 1) TC_CASE(0, 2) TC_CASE(0, 3)
    TC_CASE(1, 0) TC_CASE(1, 1) TC_CASE(1, 2) TC_CASE(1, 3)
    TC_CASE(2, 0) TC_CASE(2, 1) TC_CASE(2, 2) TC_CASE(2, 3)
    TC_CASE(3, 0) TC_CASE(3, 1) TC_CASE(3, 2) TC_CASE(3, 3)
#undef TC_CASE
    return IBINN_EINVAL;
}

#else  // IBINN_EMU: the simulator build has no tensor cores; conv.cu keeps everything on its CUDA-core kernel
bool conv_tc_supported(const ibinn_conv_args&) { return false; }
int64_t conv_tc_partials_floats(const ibinn_conv_args&) { return 0; }
int64_t conv_tc_wprep_floats(const ibinn_conv_args&) { return 0; }
int conv_tc_launch(const ibinn_conv_args&, cudaStream_t) { return IBINN_EINVAL; }
#endif
