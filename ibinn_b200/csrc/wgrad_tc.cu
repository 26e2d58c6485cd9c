// Weight / bias gradients of the stride-1 subnet convolutions on tcgen05 (3xTF32).
//
//   dW[co][ci][ky][kx] = sum_q G[q][co] * X[q + (ky-1) P + (kx-1)][ci]        (q = padded pixel position)
//
// GEMM view: the reduction dimension K is the PIXEL axis.  NCHW keeps the pixels of one channel
// contiguous, so both operands are already K-major in global memory: a 16-byte chunk = 4 consecutive
// pixels of one channel.  The batch is one padded 1-D sequence with pitch P4 (multiple of 4 >= W + pad)
// and one shared zero row between images, so a row shift (dy) moves whole chunks and only the column
// shift (dx) needs its own staged copy.  One tcgen05.mma then covers several filter taps at once:
//   A rows (M = 128) = (dy copy, co):  G shifted by -dy rows        (2 copies for Cout <= 64, 3 for <= 40)
//   B rows (N <= 256) = (dx copy, ci): X shifted by dx in {-1,0,+1} (+ one row of ones -> dbias for free)
// i.e. N = 3 Cin >= 128 for the wide layers, which is where the tensor pipe reaches its full rate.
// CTAs split the pixel axis; each keeps its [128 x N] accumulators in TMEM for its whole range and
// writes ONE partial record; a second tiny kernel sums the records in CTA order (deterministic, no
// atomics) and accumulates into dW / dbias.
#include "common.cuh"
#include "conv_internal.cuh"
#include "tc.cuh"

#ifndef IBINN_EMU
#ifdef IBINN_DBG
__device__ unsigned long long ibinn_wdbg[64];
__device__ __forceinline__ unsigned long long wgtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define WDBG(i) do { if (blockIdx.x == 1) ibinn_wdbg[i] = wgtime(); } while (0)
extern "C" int ibinn_wdebug_read(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, ibinn_wdbg, sizeof(unsigned long long) * 64); }
#else
#define WDBG(i)
#endif
namespace {

constexpr int WT_GROUPS = 2;        // staging groups of 256 threads taking K blocks round-robin (warps 0-3 also run the epilogue)
constexpr int WT_STAGERS = 256 * WT_GROUPS;
constexpr int WT_THREADS = WT_STAGERS + 32;   // + warp 16: MMA issuer
constexpr int WT_MMA_WARP = WT_STAGERS / 32;
constexpr int KC_MAX = 32;          // K chunks per stage: 4 ... 32 (16 ... 128 pixels), chosen per layer
constexpr int WT_MAXC = 128;

struct WgTcGeom {
    int pad, kk, P4, rows_img, S4, Q, nblocks, ncta, bpc;     // bpc = K blocks per CTA
    int coutp, cinp, na, ntile, nd, np, ones_row, nstage, tmem_cols;
    int a_tile_bytes, b_bytes, stage_bytes, lbo_a, lbo_b, rec_floats;
    int sdec, sh, G8, sbo_a, kcs;          // chunks per padded row, chunk shift between dy slots, 8-row groups of G, A group stride, strip chunks
    int vec_g, vec_x, kc;
    size_t smem_bytes;
};

static inline int rup(int v, int m) { return (v + m - 1) / m * m; }

static bool wg_geometry(const ibinn_wgrad_args& a, WgTcGeom& g) {
    if (a.stride == 2) {
        // stride 2: the gradient is zero-inserted onto the input grid while it is staged (even rows / columns carry g, the
        // rest is zero), after which the layer is an ordinary stride-1 weight gradient over the input pixels
        if (a.ksize != 3 || (a.Hin & 1) || (a.Win & 1) || a.Hout != a.Hin / 2 || a.Wout != a.Win / 2) return false;
    } else if (a.stride != 1 || a.Hout != a.Hin || a.Wout != a.Win) return false;
    if (a.Cout > 64 || a.Cin > 80 || a.Cout > WT_MAXC || a.Cin > WT_MAXC) return false;
    g.pad = a.ksize / 2;
    g.kk = a.ksize * a.ksize;
    g.coutp = rup(a.Cout, 8);
    g.G8 = g.coutp / 8;
    // chunks per padded row: coprime with the number of 8-row groups of G (see the strip layout below)
    g.sdec = (a.Win + g.pad + 3) / 4;
    if (a.ksize == 3) {
        auto gcd = [](int x, int y) { while (y) { const int t = x % y; x = y; y = t; } return x; };
        while (gcd(g.sdec, g.G8) != 1) ++g.sdec;
    }
    g.P4 = 4 * g.sdec;
    g.sh = a.ksize == 3 ? g.sdec : 0;
    g.rows_img = a.Hin + g.pad;
    g.S4 = g.rows_img * g.P4;
    g.Q = a.B * g.S4 + g.pad * g.P4;
    g.cinp = rup(a.Cin, 8);
    g.nd = a.ksize;                                   // dx copies
    g.na = 128 / g.coutp < a.ksize ? 128 / g.coutp : a.ksize;   // dy copies stacked in one A tile
    g.ntile = (a.ksize + g.na - 1) / g.na;
    g.ones_row = a.dbias ? g.nd * g.cinp : -1;
    g.np = rup(g.nd * g.cinp + (a.dbias ? 8 : 0), 16);
    if (g.np > 256 || g.ntile * g.np > 512) return false;
    int cols = g.ntile * g.np;
    g.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    // A operand = ONE strip of G per K block, no row-shifted copies: chunk kappa, 8-row group gg of G live at byte
    // 128 (kappa G8 + gg sh) (+16 per row); read with LBO = 128 G8 and SBO = 128 sh, descriptor row (slot d, co) of chunk k lands
    // on strip chunk k + d sh, i.e. the gradient one padded row further.  (kappa, gg) -> kappa G8 + gg sh is injective because
    // gcd(G8, sh) = 1.  The strip covers chunks [k0 - sh, k0 + kc + sh).
    g.lbo_a = 128 * g.G8;
    g.sbo_a = a.ksize == 3 ? 128 * g.sh : 128;
    // B: K-chunk pitch = an ODD number of 16-byte rows (the eight lanes of a quarter-warp write the same row of eight
    // consecutive chunks, which must fall into eight different bank groups)
    g.lbo_b = (g.np + 1) * 16;
    // enough staging work per K block to hide the load latency (>= ~4 chunk tasks per staging thread),
    // but at least 2 stages in shared memory (one per staging group)
    const int tasks_per_chunk = a.Cout + g.nd * a.Cin;
    g.kc = tasks_per_chunk * 4 >= 1200 ? 4 : tasks_per_chunk * 8 >= 1200 ? 8 : tasks_per_chunk * 16 >= 1200 ? 16 : KC_MAX;
    for (;;) {
        g.kcs = g.kc + 2 * g.sh;
        const int gs = a.ksize == 3 ? g.sh : 1;
        const int need = g.kcs * g.G8 + (g.G8 - 1) * gs + 1;                                         // written by the staging threads
        const int reach = (g.ntile - 1) * g.na * g.sh * g.G8 + (g.kc - 1) * g.G8 + 15 * gs + 1;      // read by an M = 128 descriptor (rows past the data are don't-care)
        g.a_tile_bytes = 128 * (need > reach ? need : reach);     // one of hi / lo of the strip
        g.b_bytes = g.kc * g.lbo_b;
        g.stage_bytes = 2 * (g.a_tile_bytes + g.b_bytes);
        g.nstage = (200 * 1024) / g.stage_bytes;
        if (g.nstage >= 2 || g.kc == 4) break;
        g.kc /= 2;
    }
    if (g.nstage > 4) g.nstage = 4;
    if (g.nstage < 2) return false;
    g.nblocks = (g.Q + 4 * g.kc - 1) / (4 * g.kc);
    g.smem_bytes = (size_t)g.nstage * g.stage_bytes + (5 * WT_MAXC) * 4 + 256;
    // CTA caps (measured on the full step, side stream on).  The 8x8 stage (<= 16k padded pixels) gets at most 64 CTAs: its
    // dgrad chain occupies 64-81 SMs, so a weight-gradient kernel forked onto the side stream fits beside it, and the record
    // volume (one 9 Cout Cin record per CTA, written and re-read through L2/HBM) drops with the CTA count.  The 16x16 stage
    // gets 80: half-width weight-gradient kernels interleave with the main chain's launch gaps, tails and single-tile CTAs
    // (10.8 vs 11.1 ms per step); wider layers keep all SMs.
    const int cta_cap = g.Q <= 16384 ? 64 : g.Q <= 65536 ? 80 : 148;
    g.ncta = g.nblocks < cta_cap ? g.nblocks : cta_cap;
    g.bpc = (g.nblocks + g.ncta - 1) / g.ncta;
    g.ncta = (g.nblocks + g.bpc - 1) / g.bpc;
    g.rec_floats = rup(g.kk * a.Cout * a.Cin + (a.dbias ? a.Cout : 0), 4);      // records stay 16-byte aligned
    if (a.g_pre == IBINN_PRE_BNBWD && a.gh_bs != a.g_bs) return false;      // the staging code addresses g and gh with one offset
    if ((long long)a.B * a.g_bs >= (1ll << 31) || g.kcs > 64) return false;   // 32-bit position table of <= 64 strip chunks
    g.vec_g = (a.stride == 1) && (a.Win % 4 == 0) && (a.g_bs % 4 == 0) && (((uintptr_t)a.g & 15) == 0) &&
              (a.g_pre != IBINN_PRE_BNBWD || ((a.gh_bs % 4 == 0) && (((uintptr_t)a.gh & 15) == 0)));
    g.vec_x = (a.Win % 4 == 0) && (a.x_bs % 4 == 0) && (((uintptr_t)a.x & 15) == 0);
    return true;
}

__device__ __forceinline__ float wt_invstd(const ibinn_bn_ref& r, int c) {
    const float s = r.scale[c];
    return r.scale_is_variance ? rsqrtf(s + r.eps) : s;
}
__device__ __forceinline__ void wt_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wt_store_split(unsigned char* hi, unsigned char* lo, const float (&v)[4]) {
    float h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) tc::split_tf32(v[i], h[i], l[i]);
    *reinterpret_cast<float4*>(hi) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(lo) = make_float4(l[0], l[1], l[2], l[3]);
}

// GPRE / XPRE = pre-op of the gradient / activation operand, K3 = 3x3 filter (else 1x1), VEC = 16-byte global loads:
// compile-time so that each instantiation carries only its own staging path (the loop body is streamed through the
// instruction cache once per K block).
template <int GPRE, int XPRE, bool K3, bool VEC>
__global__ void __launch_bounds__(WT_THREADS, 1) wgrad_tc_kernel(const ibinn_wgrad_args p, const WgTcGeom g, float* __restrict__ partials) {
    extern __shared__ __align__(128) unsigned char wts[];
    constexpr int ND = K3 ? 3 : 1;                       // dx copies of the activation operand
    __shared__ __align__(8) uint64_t full[4], empty[4], done;
    __shared__ uint32_t tmem_base_s;
    float* s_ga = reinterpret_cast<float*>(wts + (size_t)g.nstage * g.stage_bytes);
    float* s_gb = s_ga + WT_MAXC;
    float* s_gc = s_gb + WT_MAXC;
    float* s_xa = s_gc + WT_MAXC;
    float* s_xb = s_xa + WT_MAXC;
    __shared__ int s_postab[WT_GROUPS][64];            // per staging group: g offset of every strip chunk of the current K block (-1: padding)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int HW = p.Hin * p.Win, HWg = p.Hout * p.Wout;      // channel strides of x and of g / gh
    if (tid == 0) WDBG(0);
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { tc::mbar_init(&full[i], 256); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(&done, 1);
        tc::mbar_fence_init();
    }
    if (warp == WT_MMA_WARP) tc::tmem_alloc(&tmem_base_s, g.tmem_cols);
    ibinn_pdl_wait();           // (common.cuh: the set-up above overlaps the tail of the previous launch in the stream)
    ibinn_pdl_trigger();
    if (GPRE == IBINN_PRE_BNBWD) {
        for (int c = tid; c < p.Cout; c += WT_THREADS) {
            const float is = wt_invstd(p.g_bn, c), k = p.g_bn.gamma[c] * is;
            const float m1 = p.g_sums[c] * p.g_inv_count, m2 = p.g_sums[p.Cout + c] * p.g_inv_count;
            s_ga[c] = k;
            s_gb[c] = -k * m2 * is;
            s_gc[c] = -k * m1 + k * m2 * is * p.g_bn.mean[c];
        }
    }
    if (XPRE == IBINN_PRE_BN_RELU) {
        for (int c = tid; c < p.Cin; c += WT_THREADS) {
            const float is = wt_invstd(p.x_bn, c), A = p.x_bn.gamma[c] * is;
            s_xa[c] = A;
            s_xb[c] = p.x_bn.beta[c] - p.x_bn.mean[c] * A;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;
    if (tid == 0) WDBG(1);
    const int blk0 = blockIdx.x * g.bpc;
    const int nblk = min(g.bpc, g.nblocks - blk0);

    const int KC = g.kc;
    if (warp == WT_MMA_WARP) {
        // ================= MMA issuer (whole warp walks the loop, one elected lane issues) =================
        const uint32_t idesc = tc::idesc_tf32(128, g.np);
        const uint32_t a_step = (uint32_t)(2 * g.lbo_a) >> 4, b_step = (uint32_t)(2 * g.lbo_b) >> 4;
        for (int i = 0; i < nblk; ++i) {
            const int s = i % g.nstage;
            tc::mbar_wait(&full[s], (i / g.nstage) & 1);
            tc::fence_after_sync();
            if (lane == 0 && i < 6) WDBG(10 + i);
            const uint32_t st = tc::smem_u32(wts + (size_t)s * g.stage_bytes);
            const uint32_t b_hi = st + 2 * g.a_tile_bytes, b_lo = b_hi + g.b_bytes;
            for (int t = 0; t < g.ntile; ++t) {
                // tile t holds the dy slots [t na, t na + na): its rows start na sh chunks further into the strip
                const uint32_t a_hi = st + t * g.na * g.sh * g.lbo_a, a_lo = a_hi + g.a_tile_bytes;
                const uint32_t acc = tmem_d + t * g.np;
#pragma unroll 1
                for (int pass = 0; pass < 3; ++pass) {
                    uint64_t da = tc::smem_desc_sbo(pass == 2 ? a_lo : a_hi, g.lbo_a, g.sbo_a);
                    uint64_t db = tc::smem_desc(pass == 1 ? b_lo : b_hi, g.lbo_b);
#pragma unroll 2
                    for (int ks = 0; ks < KC / 2; ++ks) {
                        if (tc::elect_one()) tc::mma_tf32(acc, da, db, idesc, (i | pass | ks) != 0);
                        da += a_step;
                        db += b_step;
                    }
                }
            }
            if (tc::elect_one()) tc::mma_commit(&empty[s]);
        }
        if (tc::elect_one()) tc::mma_commit(&done);
        if (lane == 0) WDBG(2);
        __syncwarp();
    } else {
        // ================= operand staging: WT_GROUPS groups of 256 threads take K blocks round-robin =================
        // B operand: 256 % KC == 0, so a thread always works on the same chunk column kc of a block and its pixel coordinates
        // are decoded once per block.  A operand: one strip of G per block (see wg_geometry), tasks = (strip chunk, channel).
        const int grp = tid >> 8, lt = tid & 255;
        const int kc = lt % KC, lane_t = lt / KC, tstride = 256 / KC;
        const int n_br = p.Cin + (g.ones_row >= 0 ? 1 : 0);    // B source rows: ci (+ the ones row)
        const int gs = K3 ? g.sh : 1;                          // A strip: 8-row groups are gs * 128 bytes apart
        const int a_dk = 256 / p.Cout, a_dc = 256 % p.Cout;    // strip task e -> (chunk e / Cout, channel e % Cout), advanced by 256
        // Position of this thread's chunk in the padded sequence, kept incrementally (no divisions in the block loop):
        // q = padded position, R0 = q / P4 (padded row), c0 = q % P4, bimg = R0 / rows_img, rin = R0 % rows_img.
        const int q_step = WT_GROUPS * 4 * KC, R_step = q_step / g.P4, c_step = q_step % g.P4;
        int R0, c0, bimg, rin;
        {
            const int q = (blk0 + grp) * 4 * KC + 4 * kc;
            R0 = q / g.P4; c0 = q - R0 * g.P4;
            bimg = R0 / g.rows_img; rin = R0 - bimg * g.rows_img;
        }
        for (int i = grp; i < nblk; i += WT_GROUPS) {
            const int s = i % g.nstage;
            if (i >= g.nstage) tc::mbar_wait(&empty[s], ((i / g.nstage) - 1) & 1);
            if (lt == 0 && i < 6) WDBG(20 + 2 * i);
            unsigned char* st = wts + (size_t)s * g.stage_bytes;
            unsigned char* b_hi = st + 2 * g.a_tile_bytes;
            const size_t b_lo_off = g.b_bytes;
            // g offsets of the strip chunks [k0 - sh, k0 + KC + sh) of this block (one division chain per chunk, not per task)
            int* s_pos = s_postab[grp];
            if (lt < g.kcs) {
                const int kq = (blk0 + i) * KC - g.sh + lt;
                int off = -1;
                if (kq >= 0) {
                    const int R = kq / g.sdec, c0k = (kq - R * g.sdec) * 4;
                    const int b = R / g.rows_img, y = R - b * g.rows_img - g.pad;
                    if (c0k < p.Win && b < p.B && y >= 0) {
                        if (p.stride == 1) off = b * (int)p.g_bs + y * p.Win + c0k;
                        else if (!(y & 1)) off = b * (int)p.g_bs + (y >> 1) * p.Wout + (c0k >> 1);
                    }
                }
                s_pos[lt] = off;
            }
            asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");
            const int bb = bimg, yy = rin - g.pad;
            const bool rowok = (yy >= 0 && bb < p.B);
            // ---- issue every global load of this block first, then transform and store
            constexpr int UB = 3;
            float w6[UB][6];
            float mk[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int ci = lane_t + u * tstride;
                mk[u] = 1.f;
#pragma unroll
                for (int j = 0; j < 6; ++j) w6[u][j] = 0.f;
                if (ci < p.Cin && rowok) {
                    const float* xp = p.x + (size_t)bb * p.x_bs + (size_t)ci * HW + (size_t)yy * p.Win;
                    if (VEC && c0 + 3 < p.Win) {
                        const float4 t4 = *reinterpret_cast<const float4*>(xp + c0);
                        w6[u][1] = t4.x; w6[u][2] = t4.y; w6[u][3] = t4.z; w6[u][4] = t4.w;
                        if (K3) {
                            if (c0 > 0) w6[u][0] = __ldg(xp + c0 - 1);
                            if (c0 + 4 < p.Win) w6[u][5] = __ldg(xp + c0 + 4);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            const int c = c0 - 1 + j;
                            if (c >= 0 && c < p.Win && (K3 || (j >= 1 && j <= 4))) w6[u][j] = __ldg(xp + c);
                        }
                    }
                    if (XPRE == IBINN_PRE_BN_RELU && p.x_mask) mk[u] = __ldg(p.x_mask + (size_t)bb * p.Cin + ci);
                }
            }
            if (lt == 0 && i == 0) WDBG(40);
            {
                // ---- A: the strip of G.  Lanes run along the channel (8 channels = one 128-byte row group: conflict-free
                // stores); every task is one 16-byte chunk = 4 pixels of one channel.
                constexpr int UA = 4;
                const int ntask = g.kcs * p.Cout;
                int kap = lt / p.Cout, co = lt - kap * p.Cout;
                unsigned char* a_hi = st;
                for (int e0 = lt; e0 < ntask; e0 += 256 * UA) {
                    float v[UA][4], h[UA][4];
                    int adr[UA], cch[UA], offs[UA];
#pragma unroll
                    for (int u = 0; u < UA; ++u) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) { v[u][j] = 0.f; h[u][j] = 0.f; }
                        adr[u] = -1; cch[u] = co; offs[u] = -1;
                        if (kap < g.kcs) {
                            adr[u] = 128 * (kap * g.G8 + (co >> 3) * gs) + (co & 7) * 16;
                            const int off = s_pos[kap];
                            offs[u] = off;
                            if (off >= 0) {
                                const float* gp = p.g + (size_t)off + (size_t)co * HWg;
                                const float* hp = p.gh + (size_t)off + (size_t)co * HWg;     // same (b,c,y,x): gh_bs == g_bs is checked on the host
                                if (VEC) {
                                    const float4 t4 = *reinterpret_cast<const float4*>(gp);
                                    v[u][0] = t4.x; v[u][1] = t4.y; v[u][2] = t4.z; v[u][3] = t4.w;
                                    if (GPRE == IBINN_PRE_BNBWD) {
                                        const float4 h4 = *reinterpret_cast<const float4*>(hp);
                                        h[u][0] = h4.x; h[u][1] = h4.y; h[u][2] = h4.z; h[u][3] = h4.w;
                                    }
                                } else {
                                    // scalar path: ragged widths, unaligned tensors, and stride 2 (g zero-inserted: even pixels only)
                                    const int R = ((blk0 + i) * KC - g.sh + kap) / g.sdec;
                                    const int c0k = (((blk0 + i) * KC - g.sh + kap) - R * g.sdec) * 4;
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        if (c0k + j < p.Win && (p.stride == 1 || !(j & 1))) {
                                            const int jj = p.stride == 1 ? j : j >> 1;
                                            v[u][j] = __ldg(gp + jj);
                                            if (GPRE == IBINN_PRE_BNBWD) h[u][j] = __ldg(hp + jj);
                                        } else h[u][j] = __int_as_float(0x7fc00000);      // marks "no pixel here" for the affine pre-op below
                                    }
                                }
                            }
                        }
                        kap += a_dk; co += a_dc;
                        if (co >= p.Cout) { co -= p.Cout; ++kap; }
                    }
#pragma unroll
                    for (int u = 0; u < UA; ++u) {
                        if (adr[u] < 0) continue;
                        float o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            o[j] = v[u][j];
                            // BN backward is affine (a*g + b*h + c): padded positions must stay exactly zero
                            if (GPRE == IBINN_PRE_BNBWD && offs[u] >= 0 && (VEC || h[u][j] == h[u][j]))
                                o[j] = fmaf(s_ga[cch[u]], v[u][j], fmaf(s_gb[cch[u]], h[u][j], s_gc[cch[u]]));
                        }
                        wt_store_split(a_hi + adr[u], a_hi + g.a_tile_bytes + adr[u], o);
                    }
                }
            }
            if (lt == 0 && i == 0) WDBG(42);
            // ---- B: activation copies shifted by dx (and the row of ones)
            for (int u0 = 0; u0 * tstride + lane_t < n_br; ++u0) {
                const int ci = lane_t + u0 * tstride;
                if (ci >= p.Cin) {                                   // the ones row: 1 at valid pixels
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = (rowok && c0 + j < p.Win) ? 1.f : 0.f;
                    unsigned char* hi = b_hi + (size_t)kc * g.lbo_b + (size_t)g.ones_row * 16;
                    wt_store_split(hi, hi + b_lo_off, o);
                    continue;
                }
                float x6[6];
                if (u0 < UB) {
#pragma unroll
                    for (int u = 0; u < UB; ++u) if (u == u0) {
#pragma unroll
                        for (int j = 0; j < 6; ++j) x6[j] = w6[u][j];
                    }
                } else {                                             // more than UB channels per thread: load late (rare shapes)
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        const int c = c0 - 1 + j;
                        x6[j] = (rowok && c >= 0 && c < p.Win) ? __ldg(p.x + (size_t)bb * p.x_bs + (size_t)ci * HW + (size_t)yy * p.Win + c) : 0.f;
                    }
                }
                float mkk = 1.f;
                if (u0 < UB) {
#pragma unroll
                    for (int u = 0; u < UB; ++u) if (u == u0) mkk = mk[u];
                } else if (rowok && XPRE == IBINN_PRE_BN_RELU && p.x_mask) mkk = __ldg(p.x_mask + (size_t)bb * p.Cin + ci);
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int c = c0 - 1 + j;
                    if (!rowok || c < 0 || c >= p.Win) { x6[j] = 0.f; continue; }
                    if (XPRE == IBINN_PRE_RELU) x6[j] = fmaxf(x6[j], 0.f);
                    else if (XPRE == IBINN_PRE_BN_RELU) x6[j] = fmaxf(fmaf(x6[j], s_xa[ci], s_xb[ci]), 0.f) * mkk;
                }
                float h6[6], l6[6];                                  // split once, the dx copies are windows of it
#pragma unroll
                for (int j = 0; j < 6; ++j) tc::split_tf32(x6[j], h6[j], l6[j]);
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    const int sh = K3 ? d : 1;                       // window offset: dx = d - 1 -> x6[sh .. sh+3]
                    unsigned char* hi = b_hi + (size_t)kc * g.lbo_b + (size_t)(d * g.cinp + ci) * 16;
                    *reinterpret_cast<float4*>(hi) = make_float4(h6[sh], h6[sh + 1], h6[sh + 2], h6[sh + 3]);
                    *reinterpret_cast<float4*>(hi + b_lo_off) = make_float4(l6[sh], l6[sh + 1], l6[sh + 2], l6[sh + 3]);
                }
            }
            if (lt == 0 && i == 0) WDBG(43);
            tc::fence_async_smem();
            if (lt == 0 && i == 0) WDBG(44);
            wt_arrive(&full[s]);
            if (lt == 0 && i < 6) WDBG(21 + 2 * i);
            // advance to this group's next block
            c0 += c_step; R0 += R_step; rin += R_step;
            if (c0 >= g.P4) { c0 -= g.P4; ++R0; ++rin; }
            while (rin >= g.rows_img) { rin -= g.rows_img; ++bimg; }
        }
    }

    // ---- epilogue: warps 0-3, thread <-> TMEM lane = A row (dy copy, co).  (Bound by the record volume written to L2 -
    // 9 Cout Cin floats per CTA - not by the store pattern: a transposition through shared memory was slower.)
    if (warp < 4) {
        tc::mbar_wait(&done, 0);
        tc::fence_after_sync();
        if (tid == 0) WDBG(3);
        const int row = tid;
        const int dyi = row / g.coutp, co = row - dyi * g.coutp;
        float* rec = partials + (size_t)blockIdx.x * g.rec_floats;
        const bool vec_st = (p.Cin % 4 == 0);
        for (int t = 0; t < g.ntile; ++t) {
            const int cp = K3 ? 2 - (t * g.na + dyi) : 0;         // dy slot -> ky (slot d holds the gradient d - 1 padded rows further)
            const bool rowv = (dyi < g.na) && (co < p.Cout) && cp >= 0 && (K3 || t * g.na + dyi == 0);
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + t * g.np;
            for (int d = 0; d < ND; ++d) {
                const int tap = K3 ? cp * 3 + d : 0;           // ky = cp, kx = d
                float* dst = rec + ((size_t)tap * p.Cout + co) * p.Cin;
#pragma unroll 1
                for (int c32 = 0; c32 < g.cinp; c32 += 32) {      // up to four 8-column loads in flight per wait
                    float v[4][8];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (c32 + 8 * k < g.cinp) tc::tmem_ld8(taddr + d * g.cinp + c32 + 8 * k, v[k]);
                    tc::tmem_ld_wait();
                    if (!rowv) continue;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int c8 = c32 + 8 * k;
                        if (c8 >= g.cinp) break;
                        if (vec_st && c8 + 8 <= p.Cin) {
                            *reinterpret_cast<float4*>(dst + c8) = make_float4(v[k][0], v[k][1], v[k][2], v[k][3]);
                            *reinterpret_cast<float4*>(dst + c8 + 4) = make_float4(v[k][4], v[k][5], v[k][6], v[k][7]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) if (c8 + j < p.Cin) dst[c8 + j] = v[k][j];
                        }
                    }
                }
            }
            if (g.ones_row >= 0) {
                float v[8];
                tc::tmem_ld8(taddr + g.ones_row, v);
                tc::tmem_ld_wait();
                if (rowv && cp == (K3 ? 1 : 0)) rec[g.kk * p.Cout * p.Cin + co] = v[0];     // the unshifted (dy = 0) copy sums G over the valid pixels
            }
        }
    }
    if (tid == 0) WDBG(4);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == WT_MMA_WARP) tc::tmem_dealloc(tmem_d, g.tmem_cols);
    if (tid == 0) WDBG(5);
}

// dw[co][ci][tap] += sum_cta rec[cta][tap][co][ci];  dbias[co] += sum_cta rec[cta][kk*Cout*Cin + co]
// block = 32 elements x 8 record slices (fixed summation order -> deterministic)
__global__ void __launch_bounds__(IBINN_THREADS) wgrad_reduce_kernel(const float* __restrict__ partials, int ncta, int rec_floats, float* __restrict__ dw,
                                                                    float* __restrict__ dbias, int Cout, int Cin, int kk) {
    __shared__ float s_part[8][33];
    const int n = kk * Cout * Cin;
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    for (int e0 = blockIdx.x * 32; e0 < rec_floats; e0 += gridDim.x * 32) {
        const int e = e0 + lane;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        if (e < rec_floats) {
            int c = slice;
            for (; c + 24 < ncta; c += 32) {
                s0 += __ldg(partials + (size_t)c * rec_floats + e);
                s1 += __ldg(partials + (size_t)(c + 8) * rec_floats + e);
                s2 += __ldg(partials + (size_t)(c + 16) * rec_floats + e);
                s3 += __ldg(partials + (size_t)(c + 24) * rec_floats + e);
            }
            for (; c < ncta; c += 8) s0 += __ldg(partials + (size_t)c * rec_floats + e);
        }
        __syncthreads();
        s_part[slice][lane] = (s0 + s1) + (s2 + s3);
        __syncthreads();
        if (slice == 0 && e < rec_floats) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += s_part[k][lane];
            if (e < n) {
                const int ci = e % Cin, co = (e / Cin) % Cout, tap = e / (Cin * Cout);
                dw[((size_t)co * Cin + ci) * kk + tap] += s;
            } else if (dbias && e - n < Cout) dbias[e - n] += s;
        }
    }
}

// the same for a table of layers: blockIdx.y = layer
__global__ void __launch_bounds__(IBINN_THREADS) wgrad_reduce_many_kernel(const ibinn_wreduce_desc* __restrict__ table) {
    __shared__ float s_part[8][33];
    const ibinn_wreduce_desc d = table[blockIdx.y];
    const int n = d.kk * d.Cout * d.Cin;
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    for (int e0 = blockIdx.x * 32; e0 < d.rec_floats; e0 += gridDim.x * 32) {
        const int e = e0 + lane;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        if (e < d.rec_floats) {
            int c = slice;
            for (; c + 24 < d.ncta; c += 32) {
                s0 += __ldg(d.partials + (size_t)c * d.rec_floats + e);
                s1 += __ldg(d.partials + (size_t)(c + 8) * d.rec_floats + e);
                s2 += __ldg(d.partials + (size_t)(c + 16) * d.rec_floats + e);
                s3 += __ldg(d.partials + (size_t)(c + 24) * d.rec_floats + e);
            }
            for (; c < d.ncta; c += 8) s0 += __ldg(d.partials + (size_t)c * d.rec_floats + e);
        }
        __syncthreads();
        s_part[slice][lane] = (s0 + s1) + (s2 + s3);
        __syncthreads();
        if (slice == 0 && e < d.rec_floats) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += s_part[k][lane];
            if (e < n) {
                const int ci = e % d.Cin, co = (e / d.Cin) % d.Cout, tap = e / (d.Cin * d.Cout);
                d.dw[((size_t)co * d.Cin + ci) * d.kk + tap] += s;
            } else if (d.dbias && e - n < d.Cout) d.dbias[e - n] += s;
        }
    }
}

}  // namespace

extern "C" int ibinn_conv_wgrad_records(const ibinn_wgrad_args* a, int32_t* ncta, int32_t* rec_floats) {
    if (!a || !ncta || !rec_floats) return IBINN_ENULL;
    WgTcGeom g;
    if (!wg_geometry(*a, g)) return IBINN_EINVAL;
    *ncta = g.ncta;
    *rec_floats = g.rec_floats;
    return 0;
}

extern "C" int ibinn_wgrad_reduce_many(const ibinn_wreduce_desc* table, int n, ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!table) return IBINN_ENULL;
    if (n <= 0 || n > 65535) return IBINN_EINVAL;
    IBINN_LAUNCH(wgrad_reduce_many_kernel, dim3(48, n), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, table);
    return ibinn_launch_status();
}

bool wgrad_tc_supported(const ibinn_wgrad_args& a) {
    WgTcGeom g;
    return wg_geometry(a, g);
}

int64_t wgrad_tc_workspace_floats(const ibinn_wgrad_args& a) {
    WgTcGeom g;
    if (!wg_geometry(a, g)) return 0;
    return (int64_t)g.ncta * g.rec_floats;
}

int wgrad_tc_launch(const ibinn_wgrad_args& a, float* workspace, cudaStream_t st) {
    WgTcGeom g;
    if (!wg_geometry(a, g)) return IBINN_EINVAL;
    if (!workspace) return IBINN_ENULL;
    const bool vec = g.vec_g && g.vec_x, k3 = a.ksize == 3;
    int e = IBINN_EINVAL;
#define WT_CASE(GP, XP, K3_, V_) \
    if (a.g_pre == GP && a.x_pre == XP && k3 == K3_ && vec == V_) { \
        auto kfn = wgrad_tc_kernel<GP, XP, K3_, V_>; \
        e = ibinn_set_smem(kfn, g.smem_bytes); \
        if (e) return e; \
        IBINN_LAUNCH_PDL(kfn, dim3(g.ncta), dim3(WT_THREADS), g.smem_bytes, st, a, g, workspace); \
        e = ibinn_launch_status(); \
    }
#define WT_CASES(GP, XP) WT_CASE(GP, XP, true, true) WT_CASE(GP, XP, true, false) WT_CASE(GP, XP, false, true) WT_CASE(GP, XP, false, false)
    WT_CASES(IBINN_PRE_NONE, IBINN_PRE_NONE) WT_CASES(IBINN_PRE_NONE, IBINN_PRE_RELU) WT_CASES(IBINN_PRE_NONE, IBINN_PRE_BN_RELU)
    WT_CASES(IBINN_PRE_BNBWD, IBINN_PRE_NONE) WT_CASES(IBINN_PRE_BNBWD, IBINN_PRE_RELU) WT_CASES(IBINN_PRE_BNBWD, IBINN_PRE_BN_RELU)
#undef WT_CASES
#undef WT_CASE
    if (e || a.defer_reduce) return e;
    int nb = (g.rec_floats + 31) / 32;
    if (nb > 148 * 8) nb = 148 * 8;
    IBINN_LAUNCH(wgrad_reduce_kernel, dim3(nb), dim3(IBINN_THREADS), 0, st, workspace, g.ncta, g.rec_floats, a.dw, a.dbias, a.Cout, a.Cin, g.kk);
    return ibinn_launch_status();
}

#else
bool wgrad_tc_supported(const ibinn_wgrad_args&) { return false; }
int64_t wgrad_tc_workspace_floats(const ibinn_wgrad_args&) { return 0; }
int wgrad_tc_launch(const ibinn_wgrad_args&, float*, cudaStream_t) { return IBINN_EINVAL; }
extern "C" int ibinn_conv_wgrad_records(const ibinn_wgrad_args*, int32_t*, int32_t*) { return IBINN_EINVAL; }
extern "C" int ibinn_wgrad_reduce_many(const ibinn_wreduce_desc*, int, ibinn_stream_t) { return IBINN_EINVAL; }
#endif
