// Fused forward of a RUN of AIO_Blocks (reference all_in_one_block.py:83-114 with the subnet of inn_architecture.py:68-92):
// one CTA owns one image and carries it through every block of the run with the activations resident in shared memory.
//
// Per block and image, four contractions run on tcgen05 as 3xTF32 (fp32-accurate) MMAs, all into ONE TMEM region (the column
// group of a 128-position tile is handed from one contraction to the next as its epilogue drains it):
//   conv1  3x3  c1    -> width     A = X1 image  (relu?(x1), hi / lo split, whole image + halo, K-major 16-byte rows)
//   conv2  3x3  width -> width     A = H1 image  (relu(bn1(conv1)), written by conv1's epilogue straight into operand layout)
//   conv3  1x1  width -> 2 c2      A = H2 tile   (relu(bn2(conv2)) [* dropout mask]) in a two-slot ring
//   perm   1x1  C     -> C         A = U tile    ([x1 | x2 exp(clamp tanh(alpha s)) + t], the coupling output) in the same slot
// The 3x3 taps are nine shifted descriptors onto the same image (one padded 1-D pixel sequence per image: pitch W+1, one zero
// row below), exactly as in conv_tc.cu; filters arrive as the operand images of ibinn_wprep_many through a ring of TMA bulk copies.
// Epilogues (warps 0 .. 4 NPART - 1; thread <-> TMEM lane <-> position, the NPART warps of a lane quadrant split the channels): BatchNorm + ReLU + split -> next operand image; coupling + log-det;
// ActNorm + store + the next block's X1 image.  The block input itself stays in shared memory as fp32 ([C][HW]) for the coupling.
//
// train = 0: running statistics; nothing couples CTAs; persistent over images.
// train = 1: batch statistics: per-image (mean, M2) records -> grid barrier -> every CTA combines the records in fixed order
//            (deterministic) -> BN coefficients; two barriers per block; raw conv outputs / subnet output / block output saved.
#include "common.cuh"
#include "tc.cuh"
#ifndef IBINN_EMU
#include <type_traits>
#endif

#ifndef IBINN_EMU
#ifdef IBINN_BLK_DBG
// -DIBINN_BLK_DBG: globaltimer stamps of CTA 0's first image (phase boundaries of the first two blocks), read with ibinn_blk_debug_read
__device__ unsigned long long ibinn_blk_dbg[256];
__device__ __forceinline__ unsigned long long bgtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define BDBG(cond, i) do { if ((cond) && blockIdx.x == 0) ibinn_blk_dbg[i] = bgtime(); } while (0)
#define BDBG_ADD(cond, i, v) do { if ((cond) && blockIdx.x == 0) ibinn_blk_dbg[i] += (v); } while (0)
extern "C" int ibinn_blk_debug_read(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, ibinn_blk_dbg, sizeof(unsigned long long) * 256); }
extern "C" int ibinn_blk_debug_clear() { unsigned long long z[256] = {0}; return (int)cudaMemcpyToSymbol(ibinn_blk_dbg, z, sizeof(z)); }
#else
#define BDBG(cond, i)
#define BDBG_ADD(cond, i, v)
#endif
namespace {

constexpr int TM = 128;
#ifndef IBINN_BLK_NPART
#define IBINN_BLK_NPART 4
#endif
constexpr int NPART = IBINN_BLK_NPART;    // epilogue warps per TMEM lane quadrant: they split the channels of a position (the epilogues are
                                          // dependent-latency chains - TMEM load, tanh / exp, shared-memory store - so warps are what hides them)
constexpr int NEPI = 128 * NPART;         // epilogue threads: warps 0 .. NW_EPI-1
constexpr int NW_EPI = NEPI / 32;
constexpr int WARP_MMA = NW_EPI, WARP_PROD = NW_EPI + 1;
constexpr int BT_THREADS = NEPI + 64;     // + the MMA issuer warp and the filter producer warp
constexpr int MAXT = 4;              // tiles of 128 positions per image
constexpr int MAXSLOT = 8;           // filter ring
constexpr int MAXW = 64;             // subnet width
constexpr int MAXCH = 64;            // block channels

struct BG {
    int C, c1, c2, H, W, HW, width, P, S, halo, ntile, rows;
    int cin_p, np3, npC, kC;
    int lbo_img;                     // bytes between k-chunks of the X1 / H1 images (rows * 16)
    int x_bytes, h_bytes, hu_bytes;  // ONE of hi / lo
    int nhu, nslot, slot_bytes;
    int tpc1, tpc2, nch1, nch2;      // filter taps per ring chunk and chunks per 3x3 convolution
    int tap1_bytes, tap2_bytes, w3_bytes, wp_bytes;
    int tile_cols, tmem_cols;
    int off_x, off_h, off_hu, off_ring, off_xs, off_tab;
    int smem_bytes;
};

__host__ __device__ inline int pad_to(int v, int m) { return (v + m - 1) / m * m; }

static bool bt_geometry(const ibinn_aio_stage_args& a, BG& g) {
    if (a.C < 2 || a.C > MAXCH || (a.width != 32 && a.width != 64)) return false;      // the statistics passes work on groups of 32 channels
    g.C = a.C; g.c2 = a.C / 2; g.c1 = a.C - g.c2; g.H = a.H; g.W = a.W; g.HW = a.H * a.W; g.width = a.width;
    g.P = a.W + 1;
    g.S = (a.H + 1) * g.P;
    g.halo = g.P + 1;
    g.ntile = (g.S + TM - 1) / TM;
    if (g.ntile > MAXT) return false;
    g.rows = g.halo + g.S;
    g.cin_p = pad_to(g.c1, 8);
    g.np3 = pad_to(2 * g.c2, 16);
    g.npC = pad_to(g.C, 16);
    g.kC = pad_to(g.C, 8);
    if (g.kC > g.width) return false;                 // U shares the H2 slot
    g.lbo_img = g.rows * 16;
    g.x_bytes = (g.cin_p / 4) * g.lbo_img;
    g.h_bytes = (g.width / 4) * g.lbo_img;
    g.hu_bytes = (g.width / 4) * TM * 16;
    g.nhu = g.ntile < 2 ? 1 : 2;
    g.tap1_bytes = (g.cin_p / 4) * 2 * g.width * 16;
    g.tap2_bytes = (g.width / 4) * 2 * g.width * 16;
    g.w3_bytes = (g.width / 4) * 2 * g.np3 * 16;
    g.wp_bytes = (g.kC / 4) * 2 * g.npC * 16;
    g.slot_bytes = g.tap2_bytes;
    if (g.w3_bytes > g.slot_bytes) g.slot_bytes = g.w3_bytes;
    if (g.wp_bytes > g.slot_bytes) g.slot_bytes = g.wp_bytes;
    if (g.tap1_bytes > g.slot_bytes) g.slot_bytes = g.tap1_bytes;
    // a ring chunk carries as many consecutive taps as fit: every tap is streamed ONCE per convolution and image (the MMA
    // loops run tap-outer, tile-inner)
    g.tpc1 = g.slot_bytes / g.tap1_bytes; if (g.tpc1 > 9) g.tpc1 = 9;
    g.tpc2 = g.slot_bytes / g.tap2_bytes; if (g.tpc2 > 9) g.tpc2 = 9;
    g.nch1 = (9 + g.tpc1 - 1) / g.tpc1;
    g.nch2 = (9 + g.tpc2 - 1) / g.tpc2;
    int cols = 2 * g.width;
    if (2 * g.np3 > cols) cols = 2 * g.np3;
    if (2 * g.npC > cols) cols = 2 * g.npC;
    g.tile_cols = cols;
    const int need = g.ntile * cols;
    if (need > 512) return false;
    g.tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    // shared memory: A-operand over-reads (garbage rows of discarded positions) must land inside the allocation, so the images
    // come first and the filter ring last
    int off = 0;
    g.off_x = off;  off += 2 * g.x_bytes;
    g.off_h = off;  off += 2 * g.h_bytes;
    g.off_hu = off; off += g.nhu * 2 * g.hu_bytes;
    g.off_ring = off;
    const int tab_bytes = (4 * MAXW + 3 * MAXCH + 24 * MAXW + 64 + MAXCH) * 4;
    const int xs_bytes = pad_to(g.C * g.HW * 4, 16);
    const int budget = 225 * 1024 - tab_bytes - xs_bytes - off;
    int nslot = budget / g.slot_bytes;
    if (nslot > MAXSLOT) nslot = MAXSLOT;
    if (nslot < 2) return false;
    g.nslot = nslot;
    off += nslot * g.slot_bytes;
    g.off_xs = off;  off += xs_bytes;
    g.off_tab = off; off += tab_bytes;
    g.smem_bytes = off;
    // the last tile's shifted views read up to (ntile*TM + 2*halo) rows of the last k-chunk of an image
    const int over = (g.ntile * TM + 2 * g.halo - g.rows) * 16;
    if (g.off_h + 2 * g.h_bytes + over > g.off_ring + nslot * g.slot_bytes) return false;
    if (g.lbo_img >= (1 << 18)) return false;
    return true;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
// One arrival per epilogue WARP (the barriers count NW_EPI): 512 arrivals on one shared-memory word serialise, and they sit on
// the hand-off from every epilogue to the MMA issuer.  Each lane has fenced its own writes; __syncwarp orders them before lane 0's
// (releasing) arrive.
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory"); }
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
    return __uint_as_float(r);
}
__device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }

// Column sums over the 32 lanes of a warp for 32 values per lane: lane l ends up with sum over lanes of v[l] (31 shuffles,
// fixed order).
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// grid-wide barrier of the epilogue threads of every CTA (train mode; all CTAs are co-resident: cooperative launch)
// `episode` = 1, 2, ... counts this launch's barriers: the counter starts at zero (the last CTA to leave the kernel resets it)
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned nblk, unsigned episode, int tid) {
    epi_sync();
    if (tid == 0) {
        unsigned long long old, cur;
        asm volatile("atom.add.release.gpu.global.u64 %0, [%1], 1;" : "=l"(old) : "l"(bar) : "memory");
        const unsigned long long target = (unsigned long long)episode * nblk;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(cur) : "l"(bar) : "memory");
        } while (cur < target);
    }
    epi_sync();
}

template <bool TRAIN>
__global__ void __launch_bounds__(BT_THREADS, 1) aio_stage_kernel(const ibinn_aio_stage_args p, const BG g) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t x1_full, acc_full[MAXT], h1_done[MAXT], hu_full[2], u_full[2], w_full[MAXSLOT], w_empty[MAXSLOT];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t s_tapoff[9];
    __shared__ __align__(16) ibinn_aio_block_desc s_desc[2];      // this block's and the next block's descriptor (prefetched one block ahead)

    unsigned char* x_hi = sm + g.off_x;
    unsigned char* x_lo = x_hi + g.x_bytes;
    unsigned char* h_hi = sm + g.off_h;
    unsigned char* h_lo = h_hi + g.h_bytes;
    unsigned char* hu = sm + g.off_hu;
    unsigned char* ring = sm + g.off_ring;
    float* xs = reinterpret_cast<float*>(sm + g.off_xs);          // [C][HW] fp32 block input
    float* tab = reinterpret_cast<float*>(sm + g.off_tab);
    float* s_A1 = tab;                  // [MAXW] BN1 scale
    float* s_B1 = s_A1 + MAXW;
    float* s_A2 = s_B1 + MAXW;
    float* s_B2 = s_A2 + MAXW;
    float* s_sc = s_B2 + MAXW;          // [MAXCH] exp(act_norm)
    float* s_of = s_sc + MAXCH;         // [MAXCH] act_offset
    float* s_b3 = s_of + MAXCH;         // [MAXCH] conv3 bias
    float* s_red = s_b3 + MAXCH;        // [24][MAXW] reduction scratch
    float* s_misc = s_red + 24 * MAXW;  // [64]
    float* s_an = s_misc + 64;          // [MAXCH] act_norm (log-scales) of the current block

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int nt = g.ntile;

    if (tid < 9) s_tapoff[tid] = (uint32_t)(g.halo + (tid / 3 - 1) * g.P + (tid % 3 - 1));
    if (tid == 0) {
        tc::mbar_init(&x1_full, NW_EPI);
        for (int i = 0; i < MAXT; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&h1_done[i], NW_EPI); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&hu_full[i], NW_EPI); tc::mbar_init(&u_full[i], NW_EPI); }
        for (int i = 0; i < MAXSLOT; ++i) { tc::mbar_init(&w_full[i], 1); tc::mbar_init(&w_empty[i], 1); }
        tc::mbar_fence_init();
    }
    if (warp == WARP_MMA) tc::tmem_alloc(&tmem_base_s, g.tmem_cols);
    // zero the top halo rows of both images once (positions < 0 = the zero padding above the first image row); everything
    // else is rewritten per image
    {
        const int nx = (g.cin_p / 4) * g.halo, nh = (g.width / 4) * g.halo;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < nx; i += BT_THREADS) {
            const int k = i / g.halo, r = i - k * g.halo;
            *reinterpret_cast<float4*>(x_hi + (size_t)k * g.lbo_img + r * 16) = z;
            *reinterpret_cast<float4*>(x_lo + (size_t)k * g.lbo_img + r * 16) = z;
        }
        for (int i = tid; i < nh; i += BT_THREADS) {
            const int k = i / g.halo, r = i - k * g.halo;
            *reinterpret_cast<float4*>(h_hi + (size_t)k * g.lbo_img + r * 16) = z;
            *reinterpret_cast<float4*>(h_lo + (size_t)k * g.lbo_img + r * 16) = z;
        }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;

    const int nimg = TRAIN ? 1 : (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == WARP_PROD) {
        // ================= filter producer: chunks in consumption order through the ring =================
        if (lane == 0) {
            uint32_t ps = 0, turn = 0;                       // ring slot and how often the ring has wrapped (no division per chunk)
            auto push = [&](const float* src, int bytes) {
                if (turn) tc::mbar_wait(&w_empty[ps], (turn - 1) & 1);
                tc::mbar_arrive_expect_tx(&w_full[ps], (uint32_t)bytes);
                tc::bulk_g2s(ring + (size_t)ps * g.slot_bytes, src, (uint32_t)bytes, &w_full[ps]);
                if (++ps == (uint32_t)g.nslot) { ps = 0; ++turn; }
            };
            for (int im = 0; im < nimg; ++im)
                for (int k = 0; k < p.nblk; ++k) {
                    const ibinn_aio_block_desc* d = p.blocks + k;
                    const float* w1 = d->wprep1; const float* w2 = d->wprep2; const float* w3 = d->wprep3; const float* wp = d->wprep_perm;
                    for (int c = 0; c < g.nch1; ++c) {
                        const int t0 = c * g.tpc1, n = (9 - t0 < g.tpc1) ? 9 - t0 : g.tpc1;
                        push(w1 + (size_t)t0 * (g.tap1_bytes / 4), n * g.tap1_bytes);
                    }
                    for (int c = 0; c < g.nch2; ++c) {
                        const int t0 = c * g.tpc2, n = (9 - t0 < g.tpc2) ? 9 - t0 : g.tpc2;
                        push(w2 + (size_t)t0 * (g.tap2_bytes / 4), n * g.tap2_bytes);
                    }
                    for (int s = 0; s <= nt; ++s) {
                        if (s < nt) push(w3, g.w3_bytes);
                        if (s >= 1) push(wp, g.wp_bytes);
                    }
                }
        }
        __syncwarp();
    } else if (warp == WARP_MMA) {
        // ================= MMA issuer =================
        if (tc::elect_one()) {
            uint32_t rslot = 0, rphase = 0, nblk_seen = 0, hu_cnt[2] = {0, 0}, u_cnt[2] = {0, 0};
            const uint32_t idW2 = tc::idesc_tf32(TM, 2 * g.width), idW1 = tc::idesc_tf32(TM, g.width);
            const uint32_t id32 = tc::idesc_tf32(TM, 2 * g.np3), id31 = tc::idesc_tf32(TM, g.np3);
            const uint32_t idC2 = tc::idesc_tf32(TM, 2 * g.npC), idC1 = tc::idesc_tf32(TM, g.npC);
            // one contraction of one tile: A (hi, lo images with `lbo_a`), nks k-steps, B chunk in ring slot with `lbo_b`
            auto gemm = [&](uint32_t acc_addr, uint32_t a_hi_addr, uint32_t a_lo_addr, uint32_t lbo_a, uint32_t b_addr, uint32_t lbo_b,
                            int nks, uint32_t id2, uint32_t id1, bool first) {
                const uint64_t dah = tc::smem_desc(a_hi_addr, lbo_a), dal = tc::smem_desc(a_lo_addr, lbo_a), db = tc::smem_desc(b_addr, lbo_b);
                const uint32_t hA = (uint32_t)(dah >> 32), hB = (uint32_t)(db >> 32);
                uint32_t ah = (uint32_t)dah, al = (uint32_t)dal, b = (uint32_t)db;
                const uint32_t a_step = (2 * lbo_a) >> 4, b_step = (2 * lbo_b) >> 4;
                uint32_t acc = first ? 0u : 1u;
                for (int ks = 0; ks < nks; ++ks) {
                    tc::mma_tf32(acc_addr, ((uint64_t)hA << 32) | ah, ((uint64_t)hB << 32) | b, id2, acc);
                    tc::mma_tf32(acc_addr, ((uint64_t)hA << 32) | al, ((uint64_t)hB << 32) | b, id1, 1);
                    acc = 1;
                    ah += a_step; al += a_step; b += b_step;
                }
            };
            // The 3x3 convolutions: 27 (tap, tile) pairs x nks k-steps.  The tensor pipe needs ~47 cycles per MMA of these shapes
            // (scripts/issue_probe.cu) and the issuing lane must not be slower: the descriptor low words advance by 32-bit adds only
            // (rows are 16 bytes apart: a row shift is +1), the k-loop is unrolled at compile time and the tap offset comes from
            // arithmetic instead of a shared-memory load (the generic routine above spent ~70 cycles per MMA).  Measured with the
            // IBINN_EXP_NOMMA / IBINN_EXP_SKIPLO builds below: conv1 + conv2 now cost 50 cycles per MMA inside the kernel.
            const uint32_t hI = (uint32_t)(tc::smem_desc(0, g.lbo_img) >> 32), lI = (uint32_t)tc::smem_desc(0, g.lbo_img);      // image operands
            const uint32_t hW = (uint32_t)(tc::smem_desc(0, 2 * g.width * 16) >> 32), lW = (uint32_t)tc::smem_desc(0, 2 * g.width * 16);
            const uint32_t stepI = (uint32_t)(2 * g.lbo_img) >> 4, stepW = (uint32_t)(2 * 2 * g.width * 16) >> 4;
            auto conv_taps = [&](auto nks_tag, uint32_t img_hi16, uint32_t img_lo16, uint32_t slot16, uint32_t tap_step16, int t0, int n) {
                constexpr int NKS = decltype(nks_tag)::value;
                for (int i = 0; i < n; ++i) {
                    const int tap = t0 + i, dy = (tap * 11) >> 5, dx = tap - 3 * dy;
                    const uint32_t toff = (uint32_t)(g.halo + (dy - 1) * g.P + (dx - 1));
                    const uint32_t b0 = lW + slot16 + (uint32_t)i * tap_step16;
                    const uint32_t first = tap == 0 ? 0u : 1u;
                    uint32_t acc_addr = tmem_d, ah0 = lI + img_hi16 + toff, al0 = lI + img_lo16 + toff;
                    for (int t = 0; t < nt; ++t) {
#pragma unroll
                        for (int ks = 0; ks < NKS; ++ks) {
#if !defined(IBINN_EXP_NOMMA)
                            tc::mma_tf32(acc_addr, ((uint64_t)hI << 32) | (ah0 + ks * stepI), ((uint64_t)hW << 32) | (b0 + ks * stepW), idW2, ks ? 1u : first);
#if !defined(IBINN_EXP_SKIPLO)
                            tc::mma_tf32(acc_addr, ((uint64_t)hI << 32) | (al0 + ks * stepI), ((uint64_t)hW << 32) | (b0 + ks * stepW), idW1, 1u);
#endif
#endif
                        }
                        acc_addr += g.tile_cols; ah0 += TM; al0 += TM;
                    }
                }
            };
#ifdef IBINN_BLK_DBG
            unsigned long long c_wait = 0, c_fence = 0, c_commit = 0;
#endif
            auto ring_wait = [&]() -> uint32_t {
                const uint32_t s = rslot;
#ifdef IBINN_BLK_DBG
                const long long q0 = clock64();
#endif
                tc::mbar_wait(&w_full[s], rphase);
#ifdef IBINN_BLK_DBG
                const long long q1 = clock64();
#endif
                // (no tcgen05.fence here: the chunk was written by the TMA engine and its arrival on the mbarrier is all the MMAs need;
                // the fences of this lane order TMEM accesses against the epilogue warps and sit at x1_full / h1_done / hu_full / u_full)
#ifdef IBINN_BLK_DBG
                c_wait += q1 - q0; c_fence += clock64() - q1;
#endif
                return s;
            };
            auto ring_release = [&](uint32_t s) {
#ifdef IBINN_BLK_DBG
                const long long q0 = clock64();
#endif
                tc::mma_commit(&w_empty[s]);
                if (++rslot == (uint32_t)g.nslot) { rslot = 0; rphase ^= 1u; }
#ifdef IBINN_BLK_DBG
                c_commit += clock64() - q0;
#endif
            };
            const uint32_t xh = tc::smem_u32(x_hi), xl = tc::smem_u32(x_lo), hh = tc::smem_u32(h_hi), hl = tc::smem_u32(h_lo);
            const uint32_t ring_a = tc::smem_u32(ring), hu_a = tc::smem_u32(hu);
            for (int im = 0; im < nimg; ++im)
                for (int k = 0; k < p.nblk; ++k, ++nblk_seen) {
                    // ---- conv1
                    const bool dbg = (im == 0 && k < 2);
                    const int db = 128 + 32 * k;
                    (void)dbg; (void)db;
                    BDBG(dbg, db + 0);
                    tc::mbar_wait(&x1_full, nblk_seen & 1);
                    tc::fence_after_sync();
                    BDBG(dbg, db + 1);
#ifdef IBINN_BLK_DBG
                    unsigned long long rw_acc = 0;      // (kept in a register: a global read-modify-write per chunk would distort the phase)
#endif
                    for (int c = 0; c < g.nch1; ++c) {
#ifdef IBINN_BLK_DBG
                        const long long w0 = clock64();
#endif
                        const uint32_t s = ring_wait();
#ifdef IBINN_BLK_DBG
                        rw_acc += (unsigned long long)(clock64() - w0);
#endif
                        const int t0 = c * g.tpc1, n = (9 - t0 < g.tpc1) ? 9 - t0 : g.tpc1;
                        const uint32_t slot16 = (ring_a + s * g.slot_bytes) >> 4, ts16 = (uint32_t)g.tap1_bytes >> 4;
                        switch (g.cin_p >> 3) {
                            case 1: conv_taps(std::integral_constant<int, 1>{}, xh >> 4, xl >> 4, slot16, ts16, t0, n); break;
                            case 2: conv_taps(std::integral_constant<int, 2>{}, xh >> 4, xl >> 4, slot16, ts16, t0, n); break;
                            case 3: conv_taps(std::integral_constant<int, 3>{}, xh >> 4, xl >> 4, slot16, ts16, t0, n); break;
                            default: conv_taps(std::integral_constant<int, 4>{}, xh >> 4, xl >> 4, slot16, ts16, t0, n); break;
                        }
                        ring_release(s);
                    }
                    for (int t = 0; t < nt; ++t) tc::mma_commit(&acc_full[t]);
                    BDBG(dbg, db + 2);
#ifdef IBINN_BLK_DBG
                    if (dbg && blockIdx.x == 0) ibinn_blk_dbg[200 + k] = rw_acc;
                    rw_acc = 0;
#endif
                    // ---- conv2: tap-outer, so every tile needs the whole H1 image
                    for (int t = 0; t < nt; ++t) tc::mbar_wait(&h1_done[t], nblk_seen & 1);
                    tc::fence_after_sync();
                    BDBG(dbg, db + 3);
#ifdef IBINN_BLK_DBG
                    c_wait = c_fence = c_commit = 0;
                    if (dbg && blockIdx.x == 0) ibinn_blk_dbg[212 + k] = (unsigned long long)clock64();
#endif
                    for (int c = 0; c < g.nch2; ++c) {
#ifdef IBINN_BLK_DBG
                        const long long w0 = clock64();
#endif
                        const uint32_t s = ring_wait();
#ifdef IBINN_BLK_DBG
                        rw_acc += (unsigned long long)(clock64() - w0);
                        if (c == 0) BDBG(dbg, 208 + k);
                        if (dbg && k == 1 && blockIdx.x == 0 && c < 9) ibinn_blk_dbg[220 + 2 * c] = (unsigned long long)clock64();
#endif
                        const int t0 = c * g.tpc2, n = (9 - t0 < g.tpc2) ? 9 - t0 : g.tpc2;
                        const uint32_t slot16 = (ring_a + s * g.slot_bytes) >> 4, ts16 = (uint32_t)g.tap2_bytes >> 4;
                        if (g.width == 32) conv_taps(std::integral_constant<int, 4>{}, hh >> 4, hl >> 4, slot16, ts16, t0, n);
                        else conv_taps(std::integral_constant<int, 8>{}, hh >> 4, hl >> 4, slot16, ts16, t0, n);
#ifdef IBINN_BLK_DBG
                        if (dbg && k == 1 && blockIdx.x == 0 && c < 9) ibinn_blk_dbg[221 + 2 * c] = (unsigned long long)clock64();
#endif
                        ring_release(s);
                    }
                    for (int t = 0; t < nt; ++t) tc::mma_commit(&acc_full[t]);
                    BDBG(dbg, db + 4);
#ifdef IBINN_BLK_DBG
                    if (dbg && blockIdx.x == 0) { ibinn_blk_dbg[216 + k] = (unsigned long long)clock64(); ibinn_blk_dbg[204 + k] = rw_acc;
                        if (k == 1) { ibinn_blk_dbg[240] = c_wait; ibinn_blk_dbg[241] = c_fence; ibinn_blk_dbg[242] = c_commit; } }
#endif
                    // ---- conv3 (tile s) and soft permutation (tile s-1), software-pipelined against the epilogue
                    for (int s = 0; s <= nt; ++s) {
                        if (s < nt) {
                            const int sl = s % g.nhu;
                            tc::mbar_wait(&hu_full[sl], hu_cnt[sl] & 1); ++hu_cnt[sl];
                            tc::fence_after_sync();
                            const uint32_t rs = ring_wait();
                            const uint32_t a0 = hu_a + sl * 2 * g.hu_bytes;
                            gemm(tmem_d + s * g.tile_cols, a0, a0 + g.hu_bytes, TM * 16, ring_a + rs * g.slot_bytes, 2 * g.np3 * 16, g.width / 8, id32, id31, true);
                            ring_release(rs);
                            tc::mma_commit(&acc_full[s]);
                        }
                        if (s >= 1) {
                            const int t = s - 1, sl = t % g.nhu;
                            tc::mbar_wait(&u_full[sl], u_cnt[sl] & 1); ++u_cnt[sl];
                            tc::fence_after_sync();
                            const uint32_t rs = ring_wait();
                            const uint32_t a0 = hu_a + sl * 2 * g.hu_bytes;
                            gemm(tmem_d + t * g.tile_cols, a0, a0 + g.hu_bytes, TM * 16, ring_a + rs * g.slot_bytes, 2 * g.npC * 16, g.kC / 8, idC2, idC1, true);
                            ring_release(rs);
                            tc::mma_commit(&acc_full[t]);
                        }
                    }
                    BDBG(dbg, db + 5);
                }
        }
        __syncwarp();
    } else {
        // ================= epilogue warps: staging + every epilogue; thread <-> position (TMEM lane) =================
        const int row = tid & 127;                            // 0..127: TMEM lane; the NPART warps of a lane quadrant split the channels
        const int half = tid >> 7;                            // 0..NPART-1
        const uint32_t lane_base = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
        const float inv_n = 1.0f / (float)g.HW;
        // position decode of this thread's row in every tile
        int pixs[MAXT];                                       // pixel index inside the image or -1
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
            const int q = t * TM + row;
            const int yy = q / g.P, xx = q - yy * g.P;
            pixs[t] = (t < nt && q < g.S && yy < g.H && xx < g.W) ? yy * g.W + xx : -1;
        }

        // BatchNorm coefficients of block k from given statistics (eval) -> s_A*, s_B*
        auto load_bn_eval = [&](const ibinn_bn_ref& r, float* sA, float* sB) {
            for (int c = tid; c < g.width; c += NEPI) {
                const float sc = r.scale[c];
                const float is = r.scale_is_variance ? rsqrtf(sc + r.eps) : sc;
                const float A = r.gamma[c] * is;
                sA[c] = A;
                sB[c] = r.beta[c] - r.mean[c] * A;
            }
        };

        unsigned n_episodes = 0;
        bool sdbg = false;      // (debug build) stamp the statistics phases of this block
        (void)sdbg;
        // ---- train: per-image statistics of the accumulators of all tiles (+ save the raw conv output), grid-wide combine
        auto batch_stats = [&](int kphase, const ibinn_bn_ref& r, float* sA, float* sB, float* raw_out, float* save, float* rmean, float* rvar,
                               int64_t* nbt, int b, int rec_slot) {
            const int W2 = g.width, ng = W2 / 32;
            // work items (tile, group of 32 channels) alternate between the two warps of a lane quadrant
            // pass 1a: per-channel sums over the image (invalid rows masked by selection); the raw values are stored on the way
            float csum[2] = {0.f, 0.f};                        // lane <-> channel (group of 32)
            for (int t = 0; t < nt; ++t) {
                tc::mbar_wait(&acc_full[t], kphase & 1);
                tc::fence_after_sync();
                const bool valid = pixs[t] >= 0;
                for (int gi = 0; gi < ng; ++gi) {
                    if ((t * ng + gi) % NPART != half) continue;
                    const int c0 = gi * 32;
                    float v[32];
#pragma unroll
                    for (int c8 = 0; c8 < 32; c8 += 8) {
                        float a8[8], b8[8];
                        tc::tmem_ld8(lane_base + t * g.tile_cols + c0 + c8, a8);
                        tc::tmem_ld8(lane_base + t * g.tile_cols + W2 + c0 + c8, b8);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[c8 + i] = valid ? a8[i] + b8[i] : 0.f;
                    }
                    if (valid && raw_out) {
                        float* dst = raw_out + ((size_t)b * W2 + c0) * g.HW + pixs[t];
#pragma unroll
                        for (int i = 0; i < 32; ++i) dst[(size_t)i * g.HW] = v[i];
                    }
                    const float cs = warp_colsum32(v, lane);
                    if (gi == 0) csum[0] += cs; else csum[1] += cs;
                }
            }
            BDBG(sdbg, 16 + 8 * kphase + 0);
            s_red[warp * MAXW + lane] = csum[0];
            if (ng > 1) s_red[warp * MAXW + 32 + lane] = csum[1];
            epi_sync();
            if (tid < W2) {
                float a = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < NW_EPI; ++w8) a += s_red[w8 * MAXW + tid];
                s_red[NW_EPI * MAXW + tid] = a * inv_n;
            }
            epi_sync();
            // pass 1b: M2 about the image mean
            float csq[2] = {0.f, 0.f};
            for (int t = 0; t < nt; ++t) {
                const bool valid = pixs[t] >= 0;
                for (int gi = 0; gi < ng; ++gi) {
                    if ((t * ng + gi) % NPART != half) continue;
                    const int c0 = gi * 32;
                    float v[32];
#pragma unroll
                    for (int c8 = 0; c8 < 32; c8 += 8) {
                        float a8[8], b8[8];
                        tc::tmem_ld8(lane_base + t * g.tile_cols + c0 + c8, a8);
                        tc::tmem_ld8(lane_base + t * g.tile_cols + W2 + c0 + c8, b8);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float dlt = (a8[i] + b8[i]) - s_red[NW_EPI * MAXW + c0 + c8 + i];
                            v[c8 + i] = valid ? dlt * dlt : 0.f;
                        }
                    }
                    const float cs = warp_colsum32(v, lane);
                    if (gi == 0) csq[0] += cs; else csq[1] += cs;
                }
            }
            BDBG(sdbg, 16 + 8 * kphase + 1);
            s_red[warp * MAXW + lane] = csq[0];
            if (ng > 1) s_red[warp * MAXW + 32 + lane] = csq[1];
            epi_sync();
            float* rec = p.stat_records + ((size_t)rec_slot * p.B + b) * 2 * W2;
            if (tid < W2) {
                float a = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < NW_EPI; ++w8) a += s_red[w8 * MAXW + tid];
                rec[tid] = s_red[NW_EPI * MAXW + tid];
                rec[W2 + tid] = a;
            }
            BDBG(sdbg, 16 + 8 * kphase + 2);
            grid_barrier(p.barrier, gridDim.x, ++n_episodes, tid);
            BDBG(sdbg, 16 + 8 * kphase + 3);
            // combine the B records in fixed order: with K = mean of record 0, d_b = mean_b - K:
            //   mean = K + sum d_b / B,   M2 = sum M2_b + n (sum d_b^2 - (sum d_b)^2 / B)
            {
                const int c = tid % W2, part = tid / W2, nparts = NEPI / W2;
                const float* base = p.stat_records + (size_t)rec_slot * p.B * 2 * W2;
                const float K = ldcg(base + c);
                const float gam = tid < W2 ? r.gamma[c] : 0.f, bet = tid < W2 ? r.beta[c] : 0.f;
                float s1 = 0.f, s2 = 0.f, m2 = 0.f;
                constexpr int UB = 16;          // loads in flight per thread: one L2 round trip per batch of 16 records
                for (int b0 = part; b0 < p.B; b0 += nparts * UB) {
                    float mb[UB], qb[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        const int bb = b0 + u * nparts;
                        mb[u] = bb < p.B ? ldcg(base + (size_t)bb * 2 * W2 + c) : 0.f;
                        qb[u] = bb < p.B ? ldcg(base + (size_t)bb * 2 * W2 + W2 + c) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        const int bb = b0 + u * nparts;
                        const float dlt = bb < p.B ? mb[u] - K : 0.f;
                        s1 += dlt;
                        s2 = fmaf(dlt, dlt, s2);
                        m2 += qb[u];
                    }
                }
                epi_sync();                                   // s_red is free again
                s_red[(part * 3 + 0) * W2 + c] = s1;        // [nparts][3][W2]: NEPI / W2 * 3 * W2 = 3 NEPI floats (<= 24 MAXW)
                s_red[(part * 3 + 1) * W2 + c] = s2;
                s_red[(part * 3 + 2) * W2 + c] = m2;
                epi_sync();
                if (tid < W2) {
                    float S1 = 0.f, S2 = 0.f, M2 = 0.f;
                    for (int q = 0; q < nparts; ++q) { S1 += s_red[(q * 3 + 0) * W2 + c]; S2 += s_red[(q * 3 + 1) * W2 + c]; M2 += s_red[(q * 3 + 2) * W2 + c]; }
                    const float Bf = (float)p.B, nimgpix = (float)g.HW, N = Bf * nimgpix;
                    const float dm = S1 / Bf, mean = K + dm;
                    float tot = M2 + nimgpix * (S2 - S1 * dm);
                    if (tot < 0.f) tot = 0.f;
                    const float var_b = tot / N;
                    const float is = 1.0f / sqrtf(var_b + r.eps);
                    const float A = gam * is;
                    sA[c] = A;
                    sB[c] = bet - mean * A;
                    if (blockIdx.x == 0) {
                        save[c] = mean;
                        save[W2 + c] = is;
                        if (rmean) {
                            const long long nb = nbt ? *nbt : 0;
                            const float f = p.momentum >= 0.f ? p.momentum : 1.0f / (float)(nb + 1);
                            const float var_u = N > 1.f ? tot / (N - 1.f) : var_b;
                            rmean[c] = (1.0f - f) * rmean[c] + f * mean;
                            rvar[c] = (1.0f - f) * rvar[c] + f * var_u;
                        }
                    }
                }
                epi_sync();
                BDBG(sdbg, 16 + 8 * kphase + 4);
                if (blockIdx.x == 0 && tid == 0 && nbt) *nbt = *nbt + 1;
            }
        };

        // BN + ReLU (+ dropout mask) + hi/lo split of tile t's accumulator -> operand rows starting at (dst_hi, dst_lo) with
        // k-chunk stride `lbo`; invalid positions are written as zeros (they are the zero padding of the next 3x3 conv)
        // (`store` = false: the row lies beyond the image; the TMEM loads are warp-collective and must still be executed)
        auto bn_relu_to_operand = [&](int t, const float* sA, const float* sB, const float* dmask, unsigned char* dst_hi, unsigned char* dst_lo, int lbo,
                                      bool store) {
            const bool valid = pixs[t] >= 0;
            const int W2 = g.width;
            for (int c8 = half * 8; c8 < W2; c8 += 8 * NPART) {       // the warps of a lane quadrant take alternate groups of 8 channels
                float a8[8], b8[8], hi[8], lo[8];
                tc::tmem_ld8(lane_base + t * g.tile_cols + c8, a8);
                tc::tmem_ld8(lane_base + t * g.tile_cols + W2 + c8, b8);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float v = fmaxf(fmaf(a8[i] + b8[i], sA[c8 + i], sB[c8 + i]), 0.f);
                    if (dmask) v *= dmask[c8 + i];
                    v = valid ? v : 0.f;
                    tc::split_tf32(v, hi[i], lo[i]);
                }
                if (store) {
                    *reinterpret_cast<float4*>(dst_hi + (size_t)(c8 / 4) * lbo) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(dst_hi + (size_t)(c8 / 4 + 1) * lbo) = make_float4(hi[4], hi[5], hi[6], hi[7]);
                    *reinterpret_cast<float4*>(dst_lo + (size_t)(c8 / 4) * lbo) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<float4*>(dst_lo + (size_t)(c8 / 4 + 1) * lbo) = make_float4(lo[4], lo[5], lo[6], lo[7]);
                }
            }
        };

        // ---- per-block descriptor and tables, prefetched ONE BLOCK AHEAD: the descriptor into shared memory, the raw table values
        // (BatchNorm parameters / running statistics, ActNorm, conv3 bias: one item per thread) into registers; at the top of a block
        // only arithmetic and shared-memory stores remain (they were 2 us of dependent global-load latency per block)
        float pre[4] = {0.f, 0.f, 0.f, 0.f};
        auto prefetch_block = [&](int kn, int buf) {
            const ibinn_aio_block_desc* dn = p.blocks + kn;
            if (tid < (int)(sizeof(ibinn_aio_block_desc) / 4))
                reinterpret_cast<uint32_t*>(&s_desc[buf])[tid] = reinterpret_cast<const uint32_t*>(dn)[tid];
            if (tid < 2 * g.width) {
                const ibinn_bn_ref& r = (tid < g.width) ? dn->bn1 : dn->bn2;
                const int c = tid < g.width ? tid : tid - g.width;
                pre[0] = r.gamma[c]; pre[1] = r.beta[c];
                if (!TRAIN) { pre[2] = r.mean[c]; pre[3] = r.scale[c]; }
            } else if (tid >= 128 && tid < 128 + g.C) {
                pre[0] = dn->act_norm[tid - 128]; pre[1] = dn->act_offset[tid - 128];
            } else if (tid >= 192 && tid < 192 + 2 * g.c2) {
                pre[0] = dn->bias3[tid - 192];
            }
        };
        prefetch_block(0, 0);
        epi_sync();

        uint32_t blk_seen = 0;
        for (int im = 0; im < nimg; ++im) {
            const int b = (int)blockIdx.x + im * (int)gridDim.x;
            float jsum = 0.f, san_tot = 0.f;
            // ---- stage the first block's input: fp32 copy for the coupling + X1 operand image
            {
                const float* xb = p.x + (size_t)b * g.C * g.HW;
                for (int i = tid; i < g.C * g.HW; i += NEPI) xs[i] = __ldg(xb + i);
                epi_sync();
            }
            for (int k = 0; k < p.nblk; ++k, ++blk_seen) {
                const ibinn_aio_block_desc* d = &s_desc[blk_seen & 1];
                const bool last = (k == p.nblk - 1);
                epi_sync();                                   // every thread is done with the previous block's tables
                const bool dbg = (im == 0 && k < 2 && tid == 0);
                const int db = 32 * k;
                (void)dbg; (void)db;
                sdbg = (im == 0 && k == 1 && tid == 0);
                BDBG(dbg, db + 0);
                // X1 image of this block from xs (the previous block's epilogue has already written it for k > 0)
                if (k == 0) {
                    const int nk4 = g.cin_p / 4;
                    const bool rf = d->relu_first != 0;
                    for (int i = tid; i < nk4 * g.S; i += NEPI) {
                        const int k4 = i / g.S, q = i - k4 * g.S;
                        const int yy = q / g.P, xx = q - yy * g.P;
                        float hi[4], lo[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int ch = k4 * 4 + j;
                            float v = 0.f;
                            if (ch < g.c1 && yy < g.H && xx < g.W) {
                                v = xs[ch * g.HW + yy * g.W + xx];
                                if (rf) v = fmaxf(v, 0.f);
                            }
                            tc::split_tf32(v, hi[j], lo[j]);
                        }
                        *reinterpret_cast<float4*>(x_hi + (size_t)k4 * g.lbo_img + (g.halo + q) * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<float4*>(x_lo + (size_t)k4 * g.lbo_img + (g.halo + q) * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                // per-block tables from the values prefetched during the previous block
                if (tid < 2 * g.width) {
                    if (!TRAIN) {
                        const ibinn_bn_ref& r = (tid < g.width) ? d->bn1 : d->bn2;
                        const int c = tid < g.width ? tid : tid - g.width;
                        const float is = r.scale_is_variance ? rsqrtf(pre[3] + r.eps) : pre[3];
                        const float A = pre[0] * is;
                        (tid < g.width ? s_A1 : s_A2)[c] = A;
                        (tid < g.width ? s_B1 : s_B2)[c] = pre[1] - pre[2] * A;
                    }
                } else if (tid >= 128 && tid < 128 + g.C) {
                    s_sc[tid - 128] = expf(pre[0]); s_of[tid - 128] = pre[1]; s_an[tid - 128] = pre[0];
                } else if (tid >= 192 && tid < 192 + 2 * g.c2) {
                    s_b3[tid - 192] = pre[0];
                }
                tc::fence_async_smem();
                warp_arrive(&x1_full, lane);
                // the NEXT block's descriptor / table values (the first block of the next image after the last block)
                {
                    const int kn = last ? 0 : k + 1;
                    if (!last || im + 1 < nimg) prefetch_block(kn, (blk_seen + 1) & 1);
                }
                epi_sync();                                   // tables visible to all epilogue threads
                if (warp == 3) {
                    float v = 0.f;
                    for (int c = lane; c < g.C; c += 32) v += s_an[c];
                    san_tot += warp_sum(v);                  // the same value in every lane of warp 3
                }
                BDBG(dbg, db + 1);

                const float* dmask = (TRAIN && d->drop_mask) ? d->drop_mask + (size_t)b * g.width : nullptr;
                const bool rf_next = !last && s_desc[(blk_seen + 1) & 1].relu_first != 0;      // leading ReLU of the NEXT block's subnet (prefetched descriptor)
                // ---- B: conv1 accumulators -> H1 image
                if (TRAIN) batch_stats(0, d->bn1, s_A1, s_B1, d->h1, d->save, d->running_mean1, d->running_var1, d->nbt1, b, 0);
                for (int t = 0; t < nt; ++t) {
                    if (!TRAIN) { tc::mbar_wait(&acc_full[t], 0); tc::fence_after_sync(); }
                    if (t == 0) BDBG(dbg, db + 2);
                    const int q = t * TM + row;
                    bn_relu_to_operand(t, s_A1, s_B1, nullptr, h_hi + (g.halo + q) * 16, h_lo + (g.halo + q) * 16, g.lbo_img, q < g.S);
                    tc::fence_before_sync();
                    tc::fence_async_smem();
                    warp_arrive(&h1_done[t], lane);
                }
                BDBG(dbg, db + 3);
                // ---- train: statistics of conv2 need all of its tiles before any of them can be normalised
                if (TRAIN) batch_stats(1, d->bn2, s_A2, s_B2, d->h2, d->save + 2 * g.width, d->running_mean2, d->running_var2, d->nbt2, b, 1);
                BDBG(dbg, db + 4);
                // ---- pipelined tail: H(s-2), D(s), F(s-1)
                for (int s = 0; s < nt + 2; ++s) {
                    if (s >= 2) {
                        // H: ActNorm on the permuted tile, store, and the next block's inputs
                        const int t = s - 2;
                        tc::mbar_wait(&acc_full[t], 1);
                        tc::fence_after_sync();
                        if (s == 2) BDBG(dbg, db + 10);
                        const int pix = pixs[t];
                        const int q = t * TM + row;
                        float* og = (d->out && pix >= 0) ? d->out + (size_t)b * g.C * g.HW + pix : nullptr;
                        const int nk4 = g.cin_p / 4;
                        for (int c4 = half * 4; c4 < g.npC; c4 += 4 * NPART) {      // groups of 4 channels: C = 12 / 48 split evenly over the NPART warps
                            float a4[4], b4[4], o4[4];
                            tc::tmem_ld4(lane_base + t * g.tile_cols + c4, a4);
                            tc::tmem_ld4(lane_base + t * g.tile_cols + g.npC + c4, b4);
                            tc::tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int c = c4 + i;
                                o4[i] = 0.f;
                                if (c < g.C) {
                                    o4[i] = fmaf(a4[i] + b4[i], s_sc[c], s_of[c]);
                                    if (pix >= 0) {
                                        if (og) og[(size_t)c * g.HW] = o4[i];
                                        if (!last) xs[c * g.HW + pix] = o4[i];
                                    }
                                }
                            }
                            const int k4 = c4 / 4;
                            if (!last && q < g.S && k4 < nk4) {
                                // next block's X1 image: [relu](out[:c1]), zeros at the padding positions
                                const bool rf = rf_next;
                                float hi[4], lo[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int c = c4 + j;
                                    float v = (c < g.c1 && pix >= 0) ? o4[j] : 0.f;
                                    if (rf) v = fmaxf(v, 0.f);
                                    tc::split_tf32(v, hi[j], lo[j]);
                                }
                                *reinterpret_cast<float4*>(x_hi + (size_t)k4 * g.lbo_img + (g.halo + q) * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                                *reinterpret_cast<float4*>(x_lo + (size_t)k4 * g.lbo_img + (g.halo + q) * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                            }
                        }
                        tc::fence_before_sync();
                    }
                    if (s < nt) {
                        // D: conv2 accumulators -> H2 tile in slot s % nhu
                        const int t = s, sl = t % g.nhu;
                        tc::mbar_wait(&acc_full[t], 1);
                        tc::fence_after_sync();
                        if (s == 0) BDBG(dbg, db + 5);
                        unsigned char* dh = hu + (size_t)sl * 2 * g.hu_bytes + row * 16;
                        bn_relu_to_operand(t, s_A2, s_B2, dmask, dh, dh + g.hu_bytes, TM * 16, true);
                        tc::fence_before_sync();
                        tc::fence_async_smem();
                        warp_arrive(&hu_full[sl], lane);
                        if (s == 0) BDBG(dbg, db + 6);
                    }
                    if (s >= 1 && s - 1 < nt) {
                        // F: subnet output -> affine coupling -> U tile (same slot), log-det
                        const int t = s - 1, sl = t % g.nhu;
                        tc::mbar_wait(&acc_full[t], 0);
                        tc::fence_after_sync();
                        if (s == 1) BDBG(dbg, db + 7);
                        const int pix = pixs[t];
                        const uint32_t tb = lane_base + t * g.tile_cols;
                        unsigned char* uh = hu + (size_t)sl * 2 * g.hu_bytes + row * 16;
                        float* ag = (TRAIN && d->a && pix >= 0) ? d->a + (size_t)b * 2 * g.c2 * g.HW + pix : nullptr;
                        for (int k4 = half; k4 < g.kC / 4; k4 += NPART) {
                            float sv[4], tv[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int c = k4 * 4 + j, jj = c - g.c1;
                                sv[j] = 0.f; tv[j] = 0.f;
                                if (jj >= 0 && c < g.C) {
                                    sv[j] = tmem_ld1(tb + jj) + tmem_ld1(tb + g.np3 + jj);
                                    tv[j] = tmem_ld1(tb + g.c2 + jj) + tmem_ld1(tb + g.np3 + g.c2 + jj);
                                }
                            }
                            tc::tmem_ld_wait();
                            float hi[4], lo[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int c = k4 * 4 + j, jj = c - g.c1;
                                float u = 0.f;
                                if (c < g.C && pix >= 0) {
                                    u = xs[c * g.HW + pix];
                                    if (jj >= 0) {
                                        const float sa = sv[j] + s_b3[jj], ta = tv[j] + s_b3[g.c2 + jj];
                                        if (ag) { ag[(size_t)jj * g.HW] = sa; ag[(size_t)(g.c2 + jj) * g.HW] = ta; }
                                        const float sig = p.clamp * tanhf(p.alpha * sa);
                                        u = fmaf(u, expf(sig), ta);
                                        jsum += sig;
                                    }
                                }
                                tc::split_tf32(u, hi[j], lo[j]);
                            }
                            *reinterpret_cast<float4*>(uh + (size_t)k4 * TM * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                            *reinterpret_cast<float4*>(uh + g.hu_bytes + (size_t)k4 * TM * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                        }
                        tc::fence_before_sync();
                        tc::fence_async_smem();
                        warp_arrive(&u_full[sl], lane);
                        if (s == 1) BDBG(dbg, db + 8);
                    }
                }
                BDBG(dbg, db + 9);
                // xs / X1 of the next block are complete in shared memory; the arrive on x1_full at the top of the next
                // iteration (after its epi_sync) publishes them to the MMA issuer
            }
            // ---- log-det of the run for this image: sum of the coupling log-scales + HW * sum of the ActNorm log-scales
            {
                const float w = warp_sum(jsum);
                epi_sync();
                if (lane == 0) s_misc[warp] = w;
                if (warp == 3 && lane == 0) s_misc[NW_EPI] = san_tot;
                epi_sync();
                if (tid == 0) {
                    float js = 0.f;
#pragma unroll
                    for (int w8 = 0; w8 < NW_EPI; w8 += 4) js += (s_misc[w8] + s_misc[w8 + 1]) + (s_misc[w8 + 2] + s_misc[w8 + 3]);
                    p.jac[b] = js + (float)g.HW * s_misc[NW_EPI];
                }
                epi_sync();
            }
        }
    }

    if (TRAIN && tid == 0) {
        // every CTA is past its last barrier when it gets here: the last one out returns both counters to zero for the next launch
        const unsigned long long old = atomicAdd(p.barrier + 1, 1ull);
        if (old == (unsigned long long)gridDim.x - 1ull) {
            p.barrier[0] = 0ull;
            p.barrier[1] = 0ull;
            __threadfence();
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == WARP_MMA) tc::tmem_dealloc(tmem_d, g.tmem_cols);
}

}  // namespace

extern "C" size_t ibinn_sizeof_aio_block_desc(void) { return sizeof(ibinn_aio_block_desc); }
extern "C" size_t ibinn_sizeof_aio_stage_args(void) { return sizeof(ibinn_aio_stage_args); }

extern "C" int ibinn_aio_stage_supported(const ibinn_aio_stage_args* a) {
    if (!a) return 0;
    BG g;
    if (!bt_geometry(*a, g)) return 0;
    if (a->nblk < 1 || a->B < 1) return 0;
    if (a->train && a->B > 148) return 0;          // one co-resident CTA per image (grid barrier)
    return 1;
}

extern "C" int64_t ibinn_aio_stage_workspace_floats(const ibinn_aio_stage_args* a) {
    if (!a) return IBINN_ENULL;
    return a->train ? (int64_t)2 * a->B * 2 * a->width : 0;
}

extern "C" int ibinn_aio_stage_forward(const ibinn_aio_stage_args* ap, ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!ap) return IBINN_ENULL;
    const ibinn_aio_stage_args& a = *ap;
    if (!a.x || !a.jac || !a.blocks) return IBINN_ENULL;
    if (a.train && (!a.stat_records || !a.barrier)) return IBINN_ENULL;
    BG g;
    if (!ibinn_aio_stage_supported(ap) || !bt_geometry(a, g)) return IBINN_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = a.train ? a.B : (a.B < 148 ? a.B : 148);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(BT_THREADS);
    cfg.dynamicSmemBytes = (size_t)g.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = a.train ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ++ibinn_launch_counter;
    cudaError_t ce;
    if (a.train) {
        auto kfn = aio_stage_kernel<true>;
        e = ibinn_set_smem(kfn, g.smem_bytes);
        if (e) return e;
        ce = cudaLaunchKernelEx(&cfg, kfn, a, g);
    } else {
        auto kfn = aio_stage_kernel<false>;
        e = ibinn_set_smem(kfn, g.smem_bytes);
        if (e) return e;
        ce = cudaLaunchKernelEx(&cfg, kfn, a, g);
    }
    if (ce != cudaSuccess) {
        snprintf(ibinn_errbuf, sizeof(ibinn_errbuf), "%s:%d launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(ce));
        cudaGetLastError();
        return (int)ce;
    }
    return ibinn_launch_status();
}

#else  // IBINN_EMU: no tensor cores in the simulator build; callers use the per-op entry points
extern "C" size_t ibinn_sizeof_aio_block_desc(void) { return sizeof(ibinn_aio_block_desc); }
extern "C" size_t ibinn_sizeof_aio_stage_args(void) { return sizeof(ibinn_aio_stage_args); }
extern "C" int ibinn_aio_stage_supported(const ibinn_aio_stage_args*) { return 0; }
extern "C" int64_t ibinn_aio_stage_workspace_floats(const ibinn_aio_stage_args*) { return 0; }
extern "C" int ibinn_aio_stage_forward(const ibinn_aio_stage_args*, ibinn_stream_t) { return IBINN_EINVAL; }
#endif
