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
// The kernel is persistent and warp-specialised (10 warps): warps 0-3 = epilogue (TMEM lane quadrants), warps 4-7 =
// loaders (stage the activation window of the next tile), warp 8 = MMA issuer, warp 9 = filter producer; the
// epilogue warps help staging the first tile.
#include "common.cuh"
#include "conv_internal.cuh"
#include "tc.cuh"

#ifndef IBINN_EMU
#ifdef IBINN_DBG
__device__ unsigned long long ibinn_dbg[64];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define DBGL(i) do { ibinn_dbg[i] = gtime(); } while (0)
#define DBG(i) do { if (blockIdx.x == 0) { ibinn_dbg[i] = gtime(); if ((i) == 10 || (i) == 11) ibinn_dbg[(i) + 40] = clock64(); } } while (0)
extern "C" int ibinn_debug_read(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, ibinn_dbg, sizeof(unsigned long long) * 64); }
#else
#define DBG(i)
#define DBGL(i)
#endif
namespace {

constexpr int TM = 128;
constexpr int MAXC_TC = 256;
constexpr int TP = TM + 4;                  // row pitch of the statistics transposition tiles
constexpr int NTHREADS_TC = 320;            // warps 0-3 epilogue, 4-7 loaders, 8 MMA issuer, 9 weight producer
constexpr int MAX_RING = 9;

struct TcGeom {
    int pad, P, S, Qout, ntiles, npos, halo, cin_p, np, kk, tap_bytes, lbo_a, lbo_b, a_bytes;
    int nbuf, nst, resident, alias_stats, nstat, tmem_cols, grid;
    int off_w, off_stat, off_tab;           // byte offsets inside dynamic shared memory
    size_t smem_bytes;
};

static bool tc_geometry(const ibinn_conv_args& a, TcGeom& g) {
    if (a.Cout > 64 || a.Cin > MAXC_TC) return false;
    if (a.stride == 2) {
        // A stride-2 3x3 convolution is the stride-1 result at the even positions: the tiles walk the INPUT grid and the
        // epilogue keeps (stores, counts in the statistics) only even rows / columns.  Three quarters of the MMAs are thrown
        // away - the tensor pipe is idle most of the time in these kernels anyway (two such layers per network).
        if (a.ksize != 3 || a.upsample2 || (a.Hin & 1) || (a.Win & 1) || a.Hout != a.Hin / 2 || a.Wout != a.Win / 2) return false;
        if (a.epi != IBINN_EPI_STORE && a.epi != IBINN_EPI_BNSTATS) return false;
    } else if (a.stride != 1 || a.Hout != a.Hin || a.Wout != a.Win) return false;
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
    g.lbo_b = 2 * g.np * 16;                          // hi and lo rows of one k-chunk are adjacent
    g.a_bytes = (g.cin_p / 4) * g.lbo_a;              // one of hi / lo
    g.tap_bytes = (g.cin_p / 4) * g.lbo_b;            // hi + lo of one tap
    if (g.lbo_a >= (1 << 18)) return false;
    g.nstat = a.epi == IBINN_EPI_BNSTATS ? 1 : a.epi == IBINN_EPI_MASKSTATS ? 2 : 0;
    const int stat_bytes = g.nstat * g.np * TP * 4;
    const int red_bytes = g.nstat ? (NTHREADS_TC / 32) * 64 * 2 * 4 : 0;            // cross-warp combine of the last CTA
    const int tab_bytes = (3 * MAXC_TC + 4 * 64 + TM) * 4 + 256 + red_bytes;
    const int budget = 200 * 1024;
    // candidates, best first: double-buffered activations + resident filters ... single buffer + filter ring
    bool ok = false;
    if (g.ntiles > 148 && g.ntiles <= 2 * 148) {
        // Between one and two tiles per SM a persistent CTA pipelines nothing (tile 1 waits for tile 0's MMAs), and 40 of the
        // 148 CTAs would idle through the second tile.  Two co-resident single-tile CTAs per SM overlap their phases instead:
        // one activation buffer, statistics tiles borrowed from it, filters through a ring - within half an SM's shared memory.
        const int half = 108 * 1024;
        g.nbuf = 1;
        g.alias_stats = stat_bytes <= 2 * g.a_bytes;
        const long long fixed = 2LL * g.a_bytes + (g.alias_stats ? 0 : stat_bytes) + tab_bytes;
        int nst = (int)((half - fixed) / g.tap_bytes);
        if (nst > g.kk) nst = g.kk;
        if (nst > MAX_RING) nst = MAX_RING;
        if (half > fixed && (nst >= 3 || nst == g.kk)) {
            g.resident = (nst == g.kk);
            g.nst = nst;
            const long long op = 2LL * g.a_bytes + (long long)nst * g.tap_bytes;
            ok = true;
            g.off_w = 2 * g.a_bytes;
            g.off_stat = g.alias_stats ? 0 : (int)op;
            g.off_tab = (int)(op + (g.alias_stats ? 0 : stat_bytes));
            g.off_tab = (g.off_tab + 15) & ~15;
            g.smem_bytes = (size_t)g.off_tab + tab_bytes;
        }
    }
    for (int cand = 0; cand < 4 && !ok; ++cand) {
        g.nbuf = cand == 0 ? 2 : 1;
        g.resident = cand <= 2;
        g.alias_stats = cand >= 2;
        int nst = g.kk;
        if (!g.resident) {
            const long long room = (long long)budget - 2LL * g.a_bytes - tab_bytes;
            nst = (int)(room / g.tap_bytes);
            if (nst > MAX_RING) nst = MAX_RING;
            if (nst >= g.kk) nst = g.kk;
            if (nst < 2) continue;
        }
        g.nst = nst;
        long long op = (long long)g.nbuf * 2 * g.a_bytes + (long long)g.nst * g.tap_bytes;
        if (g.alias_stats && stat_bytes > g.nbuf * 2 * g.a_bytes) continue;     // borrowed tiles must stay inside the activation buffers
        const long long tot = op + (g.alias_stats ? 0 : stat_bytes) + tab_bytes;
        if (tot <= budget) {
            ok = true;
            g.off_w = g.nbuf * 2 * g.a_bytes;
            g.off_stat = g.alias_stats ? 0 : (int)op;
            g.off_tab = (int)(op + (g.alias_stats ? 0 : stat_bytes));
            g.off_tab = (g.off_tab + 15) & ~15;
            g.smem_bytes = (size_t)g.off_tab + tab_bytes;
        }
    }
    if (!ok) return false;
    const int cols = g.nbuf * 2 * g.np;               // per accumulator: [x_hi w_hi + x_lo w_hi | x_hi w_lo]
    g.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : 256;
    int per_sm = (int)((220 * 1024) / (g.smem_bytes + 2048));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 512 / g.tmem_cols) per_sm = 512 / g.tmem_cols;
    if (per_sm > 2) per_sm = 2;                       // __launch_bounds__(NTHREADS_TC, 2): two CTAs per SM are resident at most
    g.grid = g.ntiles < 148 * per_sm ? g.ntiles : 148 * per_sm;
    return true;
}

// ---- weight preparation: raw filters -> per tap [k4][hi rows 0..np-1 | lo rows np..2np-1][4] (canonical K-major rows of
// 16 bytes; hi and lo stacked along N so that ONE MMA of width 2*np forms x_hi*w_hi and x_hi*w_lo side by side)
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
        const size_t o = (size_t)tap * 2 * per_tap + (size_t)(k4 * 2 * np + n) * 4 + i;
        out[o] = hi;
        out[o + np * 4] = lo;
    }
}

// the same for a whole table of layers: blockIdx.y = layer
__global__ void __launch_bounds__(IBINN_THREADS) wprep_many_kernel(const ibinn_wprep_desc* __restrict__ table) {
    const ibinn_wprep_desc d = table[blockIdx.y];
    const int kk = d.ksize * d.ksize, cin_p = (d.Cin + 7) & ~7, np = (d.Cout + 15) & ~15;
    const int per_tap = (cin_p / 4) * np * 4;
    const int total = kk * per_tap;
    for (int e = blockIdx.x * IBINN_THREADS + threadIdx.x; e < total; e += gridDim.x * IBINN_THREADS) {
        const int i = e & 3, n = (e >> 2) % np, k4 = (e >> 2) / np % (cin_p / 4), tap = e / per_tap;
        const int ci = k4 * 4 + i;
        float v = 0.f;
        if (n < d.Cout && ci < d.Cin)
            v = d.transposed ? __ldg(d.w + ((size_t)ci * d.Cout + n) * kk + (kk - 1 - tap)) : __ldg(d.w + ((size_t)n * d.Cin + ci) * kk + tap);
        float hi, lo;
        tc::split_tf32(v, hi, lo);
        const size_t o = (size_t)tap * 2 * per_tap + (size_t)(k4 * 2 * np + n) * 4 + i;
        d.out[o] = hi;
        d.out[o + np * 4] = lo;
    }
}

__device__ __forceinline__ float tc_invstd(const ibinn_bn_ref& r, int c) {
    const float s = r.scale[c];
    return r.scale_is_variance ? rsqrtf(s + r.eps) : s;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Stage the activation window of one tile: global NCHW -> pre-op -> hi/lo split -> K-major rows of 16 bytes.
// Work item = (position, group of 4 channel-chunks); `lt` of `nlt` cooperating threads.
template <int PRE>
__device__ __forceinline__ void tc_stage_tile(const ibinn_conv_args& p, const TcGeom& g, int q0, unsigned char* a_hi, unsigned char* a_lo,
                                              const float* s_pa, const float* s_pb, const float* s_pc, int lt, int nlt) {
    const int nk4 = g.cin_p / 4, ngrp = (nk4 + 3) / 4;
    const int ntask = g.npos * ngrp;
    const int b_x2 = (int)p.x2_bs, b_x = (int)p.x_bs;
    const int HsWs = p.Hs * p.Ws;
    int i = lt, grp = 0;
    while (i >= g.npos) { i -= g.npos; ++grp; }
    for (int t = lt; t < ntask; t += nlt) {
        const int q = q0 - g.halo + i;
        int off = -1, bb = 0, in_img = 0;
        if (q >= 0 && q < g.Qout) {
            const int b = q / g.S, rem = q - b * g.S;
            const int yy = rem / g.P - g.pad, xx = rem % g.P;
            if (yy >= 0 && yy < p.Hin && xx < p.Win) {
                bb = b;
                if (p.upsample2) {
                    if (!((yy | xx) & 1) && (yy >> 1) < p.Hs && (xx >> 1) < p.Ws) { in_img = (yy >> 1) * p.Ws + (xx >> 1); off = b * b_x + in_img; }
                } else { in_img = yy * p.Ws + xx; off = b * b_x + in_img; }
            }
        }
        const int k0 = grp * 4;
        float v[4][4], w2[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int ch = (k0 + u) * 4 + jj;
                v[u][jj] = 0.f; w2[u][jj] = 0.f;
                if (off >= 0 && ch < p.Cin) {
                    v[u][jj] = __ldg(p.x + (size_t)off + (size_t)ch * HsWs);
                    if (PRE == IBINN_PRE_BNBWD) w2[u][jj] = __ldg(p.x2 + (size_t)bb * b_x2 + in_img + (size_t)ch * HsWs);
                    else if (PRE == IBINN_PRE_BN_RELU) { if (p.pre_mask) w2[u][jj] = __ldg(p.pre_mask + (size_t)bb * p.Cin + ch); }
                }
            }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (k0 + u >= nk4) continue;
            float h[4], l[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int ch = (k0 + u) * 4 + jj;
                float tv = v[u][jj];
                if (off >= 0 && ch < p.Cin) {
                    if (PRE == IBINN_PRE_RELU) tv = fmaxf(tv, 0.f);
                    else if (PRE == IBINN_PRE_BN_RELU) {
                        tv = fmaxf(fmaf(tv, s_pa[ch], s_pb[ch]), 0.f);
                        if (p.pre_mask) tv *= w2[u][jj];
                    } else if (PRE == IBINN_PRE_BNBWD) tv = fmaf(s_pa[ch], tv, fmaf(s_pb[ch], w2[u][jj], s_pc[ch]));
                }
                tc::split_tf32(tv, h[jj], l[jj]);
            }
            *reinterpret_cast<float4*>(a_hi + (size_t)(k0 + u) * g.lbo_a + i * 16) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(a_lo + (size_t)(k0 + u) * g.lbo_a + i * 16) = make_float4(l[0], l[1], l[2], l[3]);
        }
        i += nlt;
        while (i >= g.npos) { i -= g.npos; ++grp; }
    }
}

// Persistent, warp-specialised implicit-GEMM convolution.  Each CTA walks tiles blockIdx.x, +gridDim.x, ...
template <int PRE, int EPI>
__global__ void __launch_bounds__(NTHREADS_TC, 2) conv_tc_kernel(const ibinn_conv_args p, const TcGeom g) {
    extern __shared__ __align__(128) unsigned char tcs[];
    __shared__ __align__(8) uint64_t a_full[2], a_empty[2], acc_full[2], acc_empty[2], w_full[MAX_RING], w_empty[MAX_RING], a_help;
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_flag;
    __shared__ uint32_t s_tapoff[MAX_RING];
    constexpr bool STATS = (EPI == IBINN_EPI_BNSTATS || EPI == IBINN_EPI_MASKSTATS);

    unsigned char* wst = tcs + g.off_w;
    float* s_t1 = reinterpret_cast<float*>(tcs + g.off_stat);
    float* s_t2 = s_t1 + g.np * TP;
    float* s_pa = reinterpret_cast<float*>(tcs + g.off_tab);
    float* s_pb = s_pa + MAXC_TC;
    float* s_pc = s_pb + MAXC_TC;
    float* s_ea = s_pc + MAXC_TC;
    float* s_eb = s_ea + 64;
    float* s_em = s_eb + 64;
    float* s_ei = s_em + 64;
    float* s_valid = s_ei + 64;     // [128]

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform by construction: role branches stay on the uniform datapath
    const int HsWs = p.Hs * p.Ws;
    if (tid == 0) DBG(0);
    if (tid < g.kk) s_tapoff[tid] = (uint32_t)(g.halo + (g.pad ? (tid / 3 - 1) * g.P + (tid % 3 - 1) : 0));

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&a_full[i], 128); tc::mbar_init(&a_empty[i], 1);
            tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 128);
        }
        for (int i = 0; i < MAX_RING; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
        tc::mbar_init(&a_help, 128);
        tc::mbar_fence_init();
    }
    if (warp == 8) tc::tmem_alloc(&tmem_base_s, g.tmem_cols);
    // everything above is private to this CTA; from here on the kernel reads what its predecessor in the stream wrote
    ibinn_pdl_wait();
    ibinn_pdl_trigger();
    if (PRE == IBINN_PRE_BN_RELU) {
        for (int c = tid; c < p.Cin; c += NTHREADS_TC) {
            const float is = tc_invstd(p.pre_bn, c), A = p.pre_bn.gamma[c] * is;
            s_pa[c] = A;
            s_pb[c] = p.pre_bn.beta[c] - p.pre_bn.mean[c] * A;
        }
    } else if (PRE == IBINN_PRE_BNBWD) {
        for (int c = tid; c < p.Cin; c += NTHREADS_TC) {
            const float is = tc_invstd(p.pre_bn, c), k = p.pre_bn.gamma[c] * is;
            const float m1 = p.pre_sums[c] * p.pre_inv_count, m2 = p.pre_sums[p.Cin + c] * p.pre_inv_count;
            s_pa[c] = k;
            s_pb[c] = -k * m2 * is;
            s_pc[c] = -k * m1 + k * m2 * is * p.pre_bn.mean[c];
        }
    }
    if (EPI == IBINN_EPI_MASKSTATS) {
        for (int c = tid; c < 64; c += NTHREADS_TC) {
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
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;
    if (tid == 0) DBG(1);
    const int ntl = (g.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // tiles of this CTA

    if (warp == 9) {
        // ================= weight producer =================
        if (lane == 0) {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(p.wprep);
            if (g.resident) {
                for (int t = 0; t < g.kk; ++t) {
                    tc::mbar_arrive_expect_tx(&w_full[t], (uint32_t)g.tap_bytes);
                    tc::bulk_g2s(wst + (size_t)t * g.tap_bytes, src + (size_t)t * g.tap_bytes, g.tap_bytes, &w_full[t]);
                }
            } else {
                const int total = ntl * g.kk;
                for (int gi = 0; gi < total; ++gi) {
                    const int s = gi % g.nst, t = gi % g.kk;
                    if (gi >= g.nst) tc::mbar_wait(&w_empty[s], ((gi / g.nst) - 1) & 1);
                    tc::mbar_arrive_expect_tx(&w_full[s], g.tap_bytes);
                    tc::bulk_g2s(wst + (size_t)s * g.tap_bytes, src + (size_t)t * g.tap_bytes, g.tap_bytes, &w_full[s]);
                }
            }
        }
        __syncwarp();        // the warp reaches the closing __syncthreads as one
    } else if (warp == 8) {
        // ================= MMA issuer =================
        {
            // two MMAs per k-step: x_hi * [w_hi | w_lo] (N = 2 np) and x_lo * w_hi (N = np, onto the first np columns).
            // The issue loop is a latency-bound scalar chain: one elected lane runs it, descriptors advance by 32-bit adds
            // on their low words (address / LBO fields cannot carry), tap offsets come from a table.
            const uint32_t idesc2 = tc::idesc_tf32(TM, 2 * g.np), idesc1 = tc::idesc_tf32(TM, g.np);
            const uint64_t db0 = tc::smem_desc(tc::smem_u32(wst), g.lbo_b);
            const uint32_t hB = (uint32_t)(db0 >> 32), lB = (uint32_t)db0;
            const int nks = g.cin_p / 8;
            const uint32_t a_step = (uint32_t)(2 * g.lbo_a) >> 4, b_step = (uint32_t)(2 * g.lbo_b) >> 4, st_step = (uint32_t)g.tap_bytes >> 4;
            if (tc::elect_one()) {
                int gi = 0;
                for (int j = 0; j < ntl; ++j) {
                    const int buf = j % g.nbuf, n = j / g.nbuf;
                    const uint32_t a_base = tc::smem_u32(tcs + (size_t)buf * 2 * g.a_bytes);
                    const uint64_t da_hi = tc::smem_desc(a_base, g.lbo_a), da_lo = tc::smem_desc(a_base + g.a_bytes, g.lbo_a);
                    const uint32_t hA = (uint32_t)(da_hi >> 32), lAh = (uint32_t)da_hi, lAl = (uint32_t)da_lo;
                    const uint32_t acc_addr = tmem_d + buf * 2 * g.np;
                    if (n >= 1) tc::mbar_wait(&acc_empty[buf], (n - 1) & 1);      // epilogue has drained this accumulator
                    tc::mbar_wait(&a_full[buf], n & 1);                          // activation tile staged
                    if (j == 0) tc::mbar_wait(&a_help, 0);
                    tc::fence_after_sync();
                    if (j < 4) DBG(10 + 2 * j);
                    uint32_t acc = 0;
                    for (int t = 0; t < g.kk; ++t, ++gi) {
                        int s = t;
                        if (!g.resident) {
                            s = gi % g.nst;
                            tc::mbar_wait(&w_full[s], (gi / g.nst) & 1);
                            tc::fence_after_sync();
                        } else if (j == 0) {
                            tc::mbar_wait(&w_full[t], 0);                        // resident filters land tap by tap
                            tc::fence_after_sync();
                        }
                        const uint32_t row_off = s_tapoff[t];                    // halo + dy P + dx, in 16-byte rows
                        uint32_t ah = lAh + row_off, al = lAl + row_off, b = lB + (uint32_t)s * st_step;
#pragma unroll 4
                        for (int ks = 0; ks < nks; ++ks) {
                            tc::mma_tf32(acc_addr, ((uint64_t)hA << 32) | ah, ((uint64_t)hB << 32) | b, idesc2, acc);
                            tc::mma_tf32(acc_addr, ((uint64_t)hA << 32) | al, ((uint64_t)hB << 32) | b, idesc1, 1);
                            acc = 1;
                            ah += a_step;
                            al += a_step;
                            b += b_step;
                        }
                        if (!g.resident) tc::mma_commit(&w_empty[s]);
                    }
                    tc::mma_commit(&a_empty[buf]);
                    tc::mma_commit(&acc_full[buf]);
                    if (j < 4) DBG(11 + 2 * j);
                }
            }
        }
        __syncwarp();
    } else {
        // ================= warps 0-7: stage the activation window of each tile =================
        // Warps 4-7 (loaders) stage every tile; the still idle epilogue warps 0-3 help with the first one (256 threads,
        // signalled on the helper barrier).  One call site: the staging code exists once in the instruction stream.
        {
            const bool loader = warp >= 4;
            const int jend = loader ? ntl : 1;
            for (int j = 0; j < jend; ++j) {
                const int buf = j % g.nbuf, n = j / g.nbuf;
                const int q0 = ((int)blockIdx.x + j * (int)gridDim.x) * TM;
                unsigned char* a_hi = tcs + (size_t)buf * 2 * g.a_bytes;
                if (n >= 1) {
                    tc::mbar_wait(&a_empty[buf], (n - 1) & 1);                     // MMAs that read this buffer are done
                    if (STATS && g.alias_stats) tc::mbar_wait(&acc_empty[buf], (n - 1) & 1);   // ... and so is the epilogue that borrowed it
                }
                if (tid == 128 && j < 4) DBG(20 + 2 * j);
                tc_stage_tile<PRE>(p, g, q0, a_hi, a_hi + g.a_bytes, s_pa, s_pb, s_pc, j == 0 ? tid : tid - 128, j == 0 ? 256 : 128);
                tc::fence_async_smem();
                mbar_arrive(loader ? &a_full[buf] : &a_help);
                if (tid == 128 && j < 4) DBG(21 + 2 * j);
            }
        }
        if (warp < 4) {
            // ================= epilogue: thread <-> output row of the tile =================
            const int row = tid;                      // 0..127, TMEM lane
            const size_t HWo = (size_t)p.Hout * p.Wout;
            // statistics: channel sc is reduced by `nparts` adjacent lanes, lane `part` taking rows part, part + nparts, ...
            int nparts = 1;
            while (nparts * 2 * g.np <= TM) nparts *= 2;
            const int sc = tid / nparts, part = tid - sc * nparts, nrows = TM / nparts;
            float run_n = 0.f, run_mean = 0.f, run_m2 = 0.f, run_a1 = 0.f, run_a2 = 0.f;
            for (int j = 0; j < ntl; ++j) {
                const int buf = j % g.nbuf, n = j / g.nbuf;
                const int q = ((int)blockIdx.x + j * (int)gridDim.x) * TM + row;
                bool valid = false;
                size_t obase = 0, mbase = 0, abase = 0;
                int bimg = 0;
                if (q < g.Qout) {
                    const int b = q / g.S, rem = q - b * g.S;
                    const int yy = rem / g.P - g.pad, xx = rem % g.P;
                    if (yy >= 0 && yy < p.Hin && xx < p.Win && (p.stride == 1 || !((yy | xx) & 1))) {
                        valid = true;
                        bimg = b;
                        const int sh = p.stride - 1;
                        const size_t in_img = (size_t)(yy >> sh) * p.Wout + (xx >> sh);
                        obase = (size_t)b * p.y_bs + in_img;
                        mbase = (size_t)b * p.m_bs + in_img;
                        abase = (size_t)b * p.addend_bs + in_img;
                    }
                }
                tc::mbar_wait(&acc_full[buf], n & 1);
                tc::fence_after_sync();
                if (tid == 0 && j < 4) DBG(30 + 2 * j);
                if (STATS) s_valid[row] = valid ? 1.f : 0.f;
                const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + buf * 2 * g.np;
                // mask / addend operands of the next 8 columns are fetched while the current 8 are processed
                constexpr bool NEED_H = (EPI == IBINN_EPI_MASKSTATS || EPI == IBINN_EPI_RELUMASK);
                float hn[8], an[8];
                auto prefetch = [&](int c8n) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int c = c8n + i;
                        hn[i] = 0.f; an[i] = 0.f;
                        if (valid && c < p.Cout) {
                            if (NEED_H) hn[i] = __ldg(p.m_h + mbase + (size_t)c * HWo);
                            if ((EPI == IBINN_EPI_STORE || EPI == IBINN_EPI_RELUMASK) && p.addend) an[i] = p.addend[abase + (size_t)c * HWo];
                        }
                    }
                };
                prefetch(0);
#pragma unroll 1
                for (int c8 = 0; c8 < g.np; c8 += 8) {
                    float v[8], v2[8], hc[8], ac[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { hc[i] = hn[i]; ac[i] = an[i]; }
                    if (c8 + 8 < g.np) prefetch(c8 + 8);
                    tc::tmem_ld8(taddr + c8, v);
                    tc::tmem_ld8(taddr + g.np + c8, v2);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] += v2[i];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int c = c8 + i;
                        float o = 0.f, xh = 0.f;
                        if (valid && c < p.Cout) {
                            o = v[i];
                            if (EPI == IBINN_EPI_STORE || EPI == IBINN_EPI_RELUMASK) { if (p.bias) o += __ldg(p.bias + c); }
                            if (EPI == IBINN_EPI_MASKSTATS) {
                                const float h = hc[i];
                                o = fmaf(h, s_ea[c], s_eb[c]) > 0.f ? o : 0.f;
                                if (p.m_drop) o *= __ldg(p.m_drop + (size_t)bimg * p.Cout + c);
                                xh = (h - s_em[c]) * s_ei[c];
                            } else if (EPI == IBINN_EPI_RELUMASK) {
                                o = hc[i] > 0.f ? o : 0.f;
                            }
                            float st = o;
                            if (EPI == IBINN_EPI_STORE || EPI == IBINN_EPI_RELUMASK) st += ac[i];
                            p.y[obase + (size_t)c * HWo] = st;
                        }
                        if (STATS) {
                            s_t1[c * TP + row] = o;
                            if (EPI == IBINN_EPI_MASKSTATS) s_t2[c * TP + row] = o * xh;
                        }
                    }
                }
                if (tid == 0 && j == 0) DBG(60);
                tc::fence_before_sync();
                if (!(STATS && g.alias_stats)) mbar_arrive(&acc_empty[buf]);       // accumulator (TMEM) is free again
                if (STATS) {
                    epi_sync();
                    if (tid == 0 && j == 0) DBG(61);
                    // one pass; BatchNorm sums are taken about K = the column's first row (any value near the mean keeps
                    // a2 - a1^2/cnt free of cancellation)
                    float a1 = 0.f, a2 = 0.f, cnt = 0.f, K = 0.f;
                    if (sc < g.np) {
                        const float* t1 = s_t1 + sc * TP + part;
                        const float* t2 = s_t2 + sc * TP + part;
                        const float* vv = s_valid + part;
                        if (EPI == IBINN_EPI_BNSTATS) K = s_t1[sc * TP];
#pragma unroll 8
                        for (int r = 0; r < nrows; ++r) {
                            const float o = t1[r * nparts], w = vv[r * nparts];
                            cnt += w;
                            if (EPI == IBINN_EPI_BNSTATS) {
                                const float d = o - K;
                                a1 = fmaf(w, d, a1);
                                a2 = fmaf(w * d, d, a2);
                            } else {
                                a1 += o;
                                a2 += t2[r * nparts];
                            }
                        }
                    }
                    for (int o = nparts >> 1; o > 0; o >>= 1) {
                        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                    }
                    if (EPI == IBINN_EPI_BNSTATS) {
                        if (cnt > 0.f) {       // Chan: fold the tile (cnt, mean, M2) into the running (n, mean, M2)
                            const float d = a1 / cnt, mean = K + d, m2 = fmaxf(a2 - a1 * d, 0.f);
                            const float nn = run_n + cnt, dlt = mean - run_mean;
                            run_mean += dlt * cnt / nn;
                            run_m2 += m2 + dlt * dlt * run_n * cnt / nn;
                            run_n = nn;
                        }
                    } else {
                        run_a1 += a1;
                        run_a2 += a2;
                        run_n += cnt;
                    }
                    if (tid == 0 && j == 0) DBG(62);
                    epi_sync();                                                  // tiles may be overwritten by the next pass
                    if (g.alias_stats) mbar_arrive(&acc_empty[buf]);             // borrowed operand area is free again
                }
                if (tid == 0 && j < 4) DBG(31 + 2 * j);
            }
            if (STATS) {
                float* rec = p.partials + (size_t)blockIdx.x * (2 * p.Cout + 4);
                if (part == 0 && sc < p.Cout) {
                    if (EPI == IBINN_EPI_BNSTATS) { rec[sc] = run_mean; rec[p.Cout + sc] = run_m2; }
                    else { rec[sc] = run_a1; rec[p.Cout + sc] = run_a2; }
                }
                if (tid == 0) rec[2 * p.Cout] = run_n;
            }
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem_d, g.tmem_cols);
    if (tid == 0) DBG(2);
    if (!STATS) return;

    // ---- last CTA: combine the per-CTA records (fixed order: deterministic for a given grid).
    // With per-CTA (n_b, mean_b, M2_b):  N = sum n_b, S1 = sum n_b mean_b, S2 = sum (M2_b + n_b mean_b^2), mean = S1/N,
    // M2 = S2 - N mean^2, in double from fp32 records (the cancellation costs log10(mean^2/var) of 16 digits).
    // Every dependent global round trip is ~0.7 us here, so: ONE acq_rel ticket (the CTA barrier above makes it cover the
    // record stores of the whole CTA), the tail operands prefetched, and all record loads of a thread issued as one batch.
    const unsigned nblk = gridDim.x;
    if (tid == 0) {
        DBGL(5);
        unsigned t;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(p.counter) : "memory");
        s_flag = (t == nblk - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!s_flag) return;
    if (tid == 0) DBGL(4);
    const size_t rs = 2 * p.Cout + 4;
    const bool own = tid < p.Cout;                       // thread <-> channel for the tail
    long long nbt = 0;
    float rmean = 0.f, rvar = 0.f;
    if (EPI == IBINN_EPI_BNSTATS) {
        if (p.num_batches_tracked) nbt = *p.num_batches_tracked;
        if (own && p.running_mean) { rmean = p.running_mean[tid]; rvar = p.running_var[tid]; }
    } else if (own) {
        if (p.m_dbeta) rmean = p.m_dbeta[tid];           // accumulation targets, fetched ahead of the record loads
        if (p.m_dgamma) rvar = p.m_dgamma[tid];
    }
    // fp32 throughout, made safe by shifting: with K = mean of record 0 (any value near the mean works),
    //   s1 = sum n_b (m_b - K),  s2 = sum M2_b + n_b (m_b - K)^2,   mean = K + s1/N,   M2 = s2 - s1^2/N
    // the subtraction s2 - s1^2/N is benign because |s1/N| is a small fraction of the standard deviation.
    constexpr int NW = NTHREADS_TC / 32, UB = 15;
    float* s_red = s_valid + TM;                         // [NW][64][2] (+ one count per warp in s_cnt)
    __shared__ float s_cnt[NW];
    float Kown = 0.f;
    for (int ch0 = 0; ch0 < p.Cout; ch0 += 32) {         // warp <-> records, lane <-> channel
        const int ca = ch0 + lane;
        const bool va = ca < p.Cout;
        float Ka = 0.f;
        if (EPI == IBINN_EPI_BNSTATS) { if (va) Ka = p.partials[ca]; }
        float s0 = 0.f, s1a = 0.f, s2a = 0.f;
        for (unsigned b0 = warp; b0 < nblk; b0 += NW * UB) {
            float f0[UB], f1a[UB], f2a[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const bool in = b0 + u * NW < nblk;
                const float* rp = p.partials + (size_t)(b0 + u * NW) * rs;
                f0[u] = (in && EPI == IBINN_EPI_BNSTATS) ? rp[2 * p.Cout] : 0.f;
                f1a[u] = (in && va) ? rp[ca] : Ka;
                f2a[u] = (in && va) ? rp[p.Cout + ca] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                if (EPI == IBINN_EPI_BNSTATS) {
                    const float nb = f0[u], da = f1a[u] - Ka;
                    s0 += nb;
                    s1a = fmaf(nb, da, s1a);
                    s2a += fmaf(nb * da, da, f2a[u]);
                } else {
                    s1a += f1a[u];
                    s2a += f2a[u];
                }
            }
        }
        s_red[(warp * 64 + ca) * 2] = s1a;
        s_red[(warp * 64 + ca) * 2 + 1] = s2a;
        if (lane == 0) s_cnt[warp] = s0;
        if (tid == ca) Kown = Ka;                        // thread tid owns channel tid in the tail
    }
    if (tid == 0) DBGL(6);
    __syncthreads();
    if (tid == 0) DBGL(7);
    if (own) {
        float n = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) { n += s_cnt[w]; s1 += s_red[(w * 64 + tid) * 2]; s2 += s_red[(w * 64 + tid) * 2 + 1]; }
        if (EPI == IBINN_EPI_BNSTATS) {
            const float d = s1 / n, mean = Kown + d;
            float m2 = s2 - s1 * d;
            if (m2 < 0.f) m2 = 0.f;
            const float var_b = m2 / n;
            p.save_mean[tid] = mean;
            p.save_invstd[tid] = 1.0f / sqrtf(var_b + p.bn_eps);
            if (p.running_mean) {
                const float f = p.momentum >= 0.f ? p.momentum : 1.0f / (float)(nbt + 1);
                const float var_u = n > 1.f ? m2 / (n - 1.f) : var_b;
                p.running_mean[tid] = (1.0f - f) * rmean + f * mean;
                p.running_var[tid] = (1.0f - f) * rvar + f * var_u;
            }
        } else {
            p.sums_out[tid] = s1;
            p.sums_out[p.Cout + tid] = s2;
            if (p.m_dbeta) p.m_dbeta[tid] = rmean + s1;
            if (p.m_dgamma) p.m_dgamma[tid] = rvar + s2;
        }
    }
    if (EPI == IBINN_EPI_BNSTATS) { if (tid == 0 && p.num_batches_tracked) *p.num_batches_tracked = nbt + 1; }
    if (tid == 0) *p.counter = 0u;
    if (tid == 0) DBGL(3);
}

template <int PRE, int EPI>
static int tc_launch_inst(const ibinn_conv_args& a, const TcGeom& g, cudaStream_t st) {
    auto kfn = conv_tc_kernel<PRE, EPI>;
    int e = ibinn_set_smem(kfn, g.smem_bytes);
    if (e) return e;
    IBINN_LAUNCH_PDL(kfn, dim3(g.grid), dim3(NTHREADS_TC), g.smem_bytes, st, a, g);
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
    return (int64_t)g.grid * (2 * a.Cout + 4);
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
    int e = 0;
    if (!a.wprep_ready) {
        const int total = g.kk * (g.cin_p / 4) * g.np * 4;
        int nb = (total + IBINN_THREADS - 1) / IBINN_THREADS;
        if (nb > 148) nb = 148;
        IBINN_LAUNCH(wprep_kernel, dim3(nb), dim3(IBINN_THREADS), 0, st, a.w, a.wprep, a.Cin, a.Cout, g.cin_p, g.np, g.kk, a.w_transposed);
        e = ibinn_launch_status();
        if (e) return e;
    }
#define TC_CASE(PRE_, EPI_) if (a.pre == PRE_ && a.epi == EPI_) return tc_launch_inst<PRE_, EPI_>(a, g, st);
    TC_CASE(0, 0) TC_CASE(0, 1) TC_CASE(0, 2) TC_CASE(0, 3)
    TC_CASE(1, 0) TC_CASE(1, 1) TC_CASE(1, 2) TC_CASE(1, 3)
    TC_CASE(2, 0) TC_CASE(2, 1) TC_CASE(2, 2) TC_CASE(2, 3)
    TC_CASE(3, 0) TC_CASE(3, 1) TC_CASE(3, 2) TC_CASE(3, 3)
#undef TC_CASE
    return IBINN_EINVAL;
}

extern "C" int ibinn_wprep_many(const ibinn_wprep_desc* table, int n, ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!table) return IBINN_ENULL;
    if (n <= 0 || n > 65535) return IBINN_EINVAL;
    IBINN_LAUNCH(wprep_many_kernel, dim3(8, n), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, table);
    return ibinn_launch_status();
}

#else  // IBINN_EMU: the simulator build has no tensor cores; conv.cu keeps everything on its CUDA-core kernel
bool conv_tc_supported(const ibinn_conv_args&) { return false; }
int64_t conv_tc_partials_floats(const ibinn_conv_args&) { return 0; }
int64_t conv_tc_wprep_floats(const ibinn_conv_args&) { return 0; }
int conv_tc_launch(const ibinn_conv_args&, cudaStream_t) { return IBINN_EINVAL; }
extern "C" int ibinn_wprep_many(const ibinn_wprep_desc*, int, ibinn_stream_t) { return IBINN_EINVAL; }
#endif
