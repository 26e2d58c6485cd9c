// Subnet convolutions of the IB-INN coupling blocks: exact-fp32 direct convolution on the
// CUDA cores with the surrounding BatchNorm / ReLU / Dropout2d folded into the operand staging
// and the epilogue.  Replaces reference inn_architecture.py:68-110 (forward) and the autograd
// backward of those layers.  See include/ibinn_b200.h for the contract.
//
// Tiling: one CTA = one image x TH output rows x (up to 64) output channels; 256 threads =
// NCT channel-threads x NPT pixel-threads, a thread owns PX consecutive pixels x 8 channels.
// The input tile (with halo, pre-op applied, zero padded) and the weight slab of CI_T input
// channels are staged in shared memory; weights are read as broadcast float4.
#include "common.cuh"
#include "conv_internal.cuh"

namespace {

constexpr int COR = 8;    // output channels per thread
constexpr int CI_T = 8;   // input channels staged per iteration
constexpr int MAXC = 256; // max channels for per-channel coefficient tables

struct ConvGeom {
    int nct, npt, tpr, th, n_ytiles, n_cotiles, ri, ciw, pitch, cop, px;
    size_t smem_bytes;
};

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

static bool conv_geometry(const ibinn_conv_args& a, int px, ConvGeom& g) {
    int cout_tile = a.Cout < 64 ? a.Cout : 64;
    int nct = cout_tile <= 8 ? 1 : cout_tile <= 16 ? 2 : cout_tile <= 32 ? 4 : 8;
    g.px = px;
    g.nct = nct;
    g.npt = IBINN_THREADS / nct;
    g.tpr = ceil_div(a.Wout, px);
    if (g.tpr > g.npt) return false;
    g.th = g.npt / g.tpr;
    if (g.th > a.Hout) g.th = a.Hout;
    g.n_ytiles = ceil_div(a.Hout, g.th);
    g.cop = nct * COR;
    g.n_cotiles = ceil_div(a.Cout, g.cop);
    g.ri = (g.th - 1) * a.stride + a.ksize;
    g.ciw = (g.tpr * px - 1) * a.stride + a.ksize;
    g.pitch = g.ciw | 1;
    size_t fl = 3 * MAXC + 4 * 64                      // pre coefficient tables, epilogue tables
                + (size_t)CI_T * g.ri * g.pitch        // input tile
                + (size_t)CI_T * a.ksize * a.ksize * g.cop  // weight slab
                + 8 * COR + 8;                          // reduction scratch
    g.smem_bytes = fl * sizeof(float);
    return true;
}

// pick the widest pixel tile that does not leave most pixel-threads idle
static bool conv_pick(const ibinn_conv_args& a, ConvGeom& g) {
    const int cand[3] = {4, 2, 1};
    bool have = false;
    double best = -1;
    ConvGeom bg;
    for (int i = 0; i < 3; ++i) {
        ConvGeom t;
        if (!conv_geometry(a, cand[i], t)) continue;
        // utilisation of pixel threads: useful pixels / (threads * px)
        double useful = (double)a.Hout * a.Wout;
        double slots = (double)t.n_ytiles * t.npt * cand[i];
        double u = useful / slots * (cand[i] == 4 ? 1.0 : cand[i] == 2 ? 0.93 : 0.8);
        if (u > best) { best = u; bg = t; have = true; }
    }
    if (have) g = bg;
    return have;
}

struct PreTab { float* a; float* b; float* c; };

__device__ __forceinline__ float conv_load_pre(const ibinn_conv_args& p, const PreTab& t, int b, int c, int iy, int ix) {
    int sy = iy, sx = ix;
    if (p.upsample2) {
        if ((iy | ix) & 1) return 0.f;
        sy = iy >> 1; sx = ix >> 1;
        if (sy >= p.Hs || sx >= p.Ws) return 0.f;
    }
    const size_t off = ((size_t)c * p.Hs + sy) * p.Ws + sx;
    float v = __ldg(p.x + (size_t)b * p.x_bs + off);
    switch (p.pre) {
        case IBINN_PRE_RELU: v = fmaxf(v, 0.f); break;
        case IBINN_PRE_BN_RELU:
            v = fmaxf(fmaf(v, t.a[c], t.b[c]), 0.f);
            if (p.pre_mask) v *= __ldg(p.pre_mask + (size_t)b * p.Cin + c);
            break;
        case IBINN_PRE_BNBWD:
            v = fmaf(t.a[c], v, fmaf(t.b[c], __ldg(p.x2 + (size_t)b * p.x2_bs + off), t.c[c]));
            break;
        default: break;
    }
    return v;
}

__device__ __forceinline__ float bn_invstd(const ibinn_bn_ref& r, int c) {
    float s = r.scale[c];
    return r.scale_is_variance ? rsqrtf(s + r.eps) : s;
}

// sum over the pixel-threads that share a channel-thread index; every thread of the group gets it
template <int N>
__device__ __forceinline__ void group_sum(float (&v)[N], float* s_red, int npt) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int wpg = npt >> 5;                  // warps per channel-thread group
    const int g0 = (w / wpg) * wpg;
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) s_red[w * N + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) {
        float t = 0.f;
        for (int i = 0; i < wpg; ++i) t += s_red[(g0 + i) * N + k];
        v[k] = t;
    }
}

template <int KS, int PX, int S>
__global__ void __launch_bounds__(IBINN_THREADS) conv_kernel(const ibinn_conv_args p, const ConvGeom g) {
    constexpr int KK = KS * KS;
    constexpr int PAD = KS / 2;
    constexpr int XW = (PX - 1) * S + KS;
    IBINN_DYN_SMEM(float, smem);
    __shared__ int s_flag;

    float* s_pa = smem;
    float* s_pb = s_pa + MAXC;
    float* s_pc = s_pb + MAXC;
    float* s_ea = s_pc + MAXC;     // epilogue: scale, shift, mean, invstd  (4 x 64)
    float* s_eb = s_ea + 64;
    float* s_em = s_eb + 64;
    float* s_ei = s_em + 64;
    float* s_in = s_ei + 64;
    float* s_w = s_in + (size_t)CI_T * g.ri * g.pitch;
    float* s_red = s_w + (size_t)CI_T * KK * g.cop;

    const int tid = threadIdx.x;
    const int ct = tid / g.npt, pt = tid % g.npt;
    const int ty = pt / g.tpr, tx = pt % g.tpr;
    const int ytile = blockIdx.x % g.n_ytiles, cotile = blockIdx.x / g.n_ytiles;
    const int b = blockIdx.y;
    const int y0 = ytile * g.th;
    const int co0 = cotile * g.cop;
    const int yy = y0 + ty;
    const bool pix_thread = (ty < g.th) && (yy < p.Hout);
    const PreTab pt_tab{s_pa, s_pb, s_pc};

    // ---- per-channel coefficient tables
    if (p.pre == IBINN_PRE_BN_RELU) {
        for (int c = tid; c < p.Cin; c += IBINN_THREADS) {
            float is = bn_invstd(p.pre_bn, c);
            float A = p.pre_bn.gamma[c] * is;
            s_pa[c] = A;
            s_pb[c] = p.pre_bn.beta[c] - p.pre_bn.mean[c] * A;
        }
    } else if (p.pre == IBINN_PRE_BNBWD) {
        // g_raw = gamma*is*(g - m1 - xhat*m2), xhat = (h-mean)*is, m1 = sum g / N, m2 = sum g*xhat / N
        for (int c = tid; c < p.Cin; c += IBINN_THREADS) {
            float is = bn_invstd(p.pre_bn, c);
            float k = p.pre_bn.gamma[c] * is;
            float m1 = p.pre_sums[c] * p.pre_inv_count;
            float m2 = p.pre_sums[p.Cin + c] * p.pre_inv_count;
            float mean = p.pre_bn.mean[c];
            s_pa[c] = k;
            s_pb[c] = -k * m2 * is;
            s_pc[c] = -k * m1 + k * m2 * is * mean;
        }
    }
    if (p.epi == IBINN_EPI_MASKSTATS) {
        for (int c = tid; c < g.cop; c += IBINN_THREADS) {
            int cg = co0 + c;
            float A = 0.f, Bc = 0.f, mean = 0.f, is = 0.f;
            if (cg < p.Cout) {
                is = bn_invstd(p.m_bn, cg);
                mean = p.m_bn.mean[cg];
                A = p.m_bn.gamma[cg] * is;
                Bc = p.m_bn.beta[cg] - mean * A;
            }
            s_ea[c] = A; s_eb[c] = Bc; s_em[c] = mean; s_ei[c] = is;
        }
    }

    float acc[PX][COR];
#pragma unroll
    for (int i = 0; i < PX; ++i)
#pragma unroll
        for (int k = 0; k < COR; ++k) acc[i][k] = 0.f;

    const int tile_n = g.ri * g.ciw;
    for (int c0 = 0; c0 < p.Cin; c0 += CI_T) {
        const int nci = min(CI_T, p.Cin - c0);
        __syncthreads();
        for (int e = tid; e < nci * tile_n; e += IBINN_THREADS) {
            const int c = e / tile_n;
            const int rem = e - c * tile_n;
            const int r = rem / g.ciw;
            const int col = rem - r * g.ciw;
            const int iy = y0 * S - PAD + r, ix = col - PAD;
            float v = 0.f;
            if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win) v = conv_load_pre(p, pt_tab, b, c0 + c, iy, ix);
            s_in[(c * g.ri + r) * g.pitch + col] = v;
        }
        for (int e = tid; e < nci * KK * g.cop; e += IBINN_THREADS) {
            const int co = e % g.cop;
            const int t = (e / g.cop) % KK;
            const int c = e / (g.cop * KK);
            const int cg = co0 + co;
            float v = 0.f;
            if (cg < p.Cout) {
                v = p.w_transposed ? __ldg(p.w + ((size_t)(c0 + c) * p.Cout + cg) * KK + (KK - 1 - t))
                                   : __ldg(p.w + ((size_t)cg * p.Cin + (c0 + c)) * KK + t);
            }
            s_w[(c * KK + t) * g.cop + co] = v;
        }
        __syncthreads();
        if (pix_thread) {
            for (int c = 0; c < nci; ++c) {
#pragma unroll
                for (int ky = 0; ky < KS; ++ky) {
                    const float* row = s_in + (c * g.ri + ty * S + ky) * g.pitch + tx * PX * S;
                    float xin[XW];
#pragma unroll
                    for (int j = 0; j < XW; ++j) xin[j] = row[j];
#pragma unroll
                    for (int kx = 0; kx < KS; ++kx) {
                        const float* wp = s_w + (c * KK + ky * KS + kx) * g.cop + ct * COR;
                        const float4 w0 = ld4(wp), w1 = ld4(wp + 4);
                        const float wv[COR] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                        for (int i = 0; i < PX; ++i)
#pragma unroll
                            for (int k = 0; k < COR; ++k) acc[i][k] = fmaf(xin[i * S + kx], wv[k], acc[i][k]);
                    }
                }
            }
        }
    }

    // ---- epilogue
    const int xbase = tx * PX;
    const size_t HWo = (size_t)p.Hout * p.Wout;
    bool pvalid[PX];
#pragma unroll
    for (int i = 0; i < PX; ++i) pvalid[i] = pix_thread && (xbase + i < p.Wout);

    float s1[COR], s2[COR];
#pragma unroll
    for (int k = 0; k < COR; ++k) { s1[k] = 0.f; s2[k] = 0.f; }

#pragma unroll
    for (int k = 0; k < COR; ++k) {
        const int cl = ct * COR + k;
        const int cg = co0 + cl;
        const bool cvalid = cg < p.Cout;
        const size_t obase = (size_t)cg * HWo + (size_t)yy * p.Wout + xbase;
        float bias = (cvalid && p.bias) ? __ldg(p.bias + cg) : 0.f;
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            if (!(pvalid[i] && cvalid)) continue;
            float v = acc[i][k] + bias;
            if (p.epi == IBINN_EPI_MASKSTATS) {
                const float h = __ldg(p.m_h + (size_t)b * p.m_bs + obase + i);
                const float act = fmaf(h, s_ea[cl], s_eb[cl]);
                v = act > 0.f ? v : 0.f;
                if (p.m_drop) v *= __ldg(p.m_drop + (size_t)b * p.Cout + cg);
                s1[k] += v;
                s2[k] += v * ((h - s_em[cl]) * s_ei[cl]);
            } else if (p.epi == IBINN_EPI_RELUMASK) {
                const float h = __ldg(p.m_h + (size_t)b * p.m_bs + obase + i);
                v = h > 0.f ? v : 0.f;
            } else if (p.epi == IBINN_EPI_BNSTATS) {
                s1[k] += v;
            }
            if (p.addend) v += p.addend[(size_t)b * p.addend_bs + obase + i];
            p.y[(size_t)b * p.y_bs + obase + i] = v;
        }
    }

    if (p.epi != IBINN_EPI_BNSTATS && p.epi != IBINN_EPI_MASKSTATS) return;

    // ---- per-CTA partial statistics -> global, last CTA combines them in a fixed order
    const int rows_here = min(g.th, p.Hout - y0);
    const float n_here = (float)(rows_here * p.Wout);
    const unsigned nblk = gridDim.x * gridDim.y;
    const unsigned blk = blockIdx.y * gridDim.x + blockIdx.x;
    if (p.epi == IBINN_EPI_BNSTATS) {
        group_sum<COR>(s1, s_red, g.npt);
#pragma unroll
        for (int k = 0; k < COR; ++k) {
            const float m = s1[k] / n_here;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < PX; ++i)
                if (pvalid[i]) { float d = acc[i][k] - m; q = fmaf(d, d, q); }
            s2[k] = q;
            s1[k] = m;
        }
        group_sum<COR>(s2, s_red, g.npt);
    } else {
        group_sum<COR>(s1, s_red, g.npt);
        group_sum<COR>(s2, s_red, g.npt);
    }
    if (pt == 0) {
#pragma unroll
        for (int k = 0; k < COR; ++k) {
            const int cg = co0 + ct * COR + k;
            if (cg < p.Cout) {
                p.partials[((size_t)blk * 2 + 0) * p.Cout + cg] = s1[k];
                p.partials[((size_t)blk * 2 + 1) * p.Cout + cg] = s2[k];
            }
        }
    }
    if (!last_block_ticket(p.counter, nblk, &s_flag)) return;

    int cpad = 8;
    while (cpad < p.Cout) cpad <<= 1;
    const int tpc = IBINN_THREADS / cpad > 32 ? 32 : IBINN_THREADS / cpad;   // threads per channel (power of 2)
    const int c = tid / tpc, j = tid % tpc;
    const bool cval = (c < p.Cout) && (tid < cpad * tpc);
    if (p.epi == IBINN_EPI_BNSTATS) {
        double n = 0., mean = 0., m2 = 0.;
        if (cval) {
            for (unsigned bb = j; bb < nblk; bb += tpc) {
                const int yt = (bb % gridDim.x) % g.n_ytiles;
                const double nb = (double)(min(g.th, p.Hout - yt * g.th) * p.Wout);
                const double mb = p.partials[((size_t)bb * 2 + 0) * p.Cout + c];
                const double qb = p.partials[((size_t)bb * 2 + 1) * p.Cout + c];
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
                // symmetric form so that both partners compute identical values
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
                a1 += (double)p.partials[((size_t)bb * 2 + 0) * p.Cout + c];
                a2 += (double)p.partials[((size_t)bb * 2 + 1) * p.Cout + c];
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

static int conv_validate(const ibinn_conv_args& a) {
    if (!a.x || !a.w || !a.y) return IBINN_ENULL;
    if (a.B <= 0 || a.Cin <= 0 || a.Cout <= 0 || a.Hout <= 0 || a.Wout <= 0) return IBINN_EINVAL;
    if (a.B > 65535 || a.Cin > MAXC) return IBINN_EINVAL;
    if (!(a.ksize == 1 || a.ksize == 3) || !(a.stride == 1 || a.stride == 2)) return IBINN_EINVAL;
    if (a.ksize == 1 && a.stride != 1) return IBINN_EINVAL;
    const int pad = a.ksize / 2;
    if (a.Hout != (a.Hin + 2 * pad - a.ksize) / a.stride + 1) return IBINN_EINVAL;
    if (a.Wout != (a.Win + 2 * pad - a.ksize) / a.stride + 1) return IBINN_EINVAL;
    if (a.upsample2) {
        if (a.stride != 1) return IBINN_EINVAL;
        if ((a.Hin + 1) / 2 != a.Hs || (a.Win + 1) / 2 != a.Ws) return IBINN_EINVAL;
    } else if (a.Hin != a.Hs || a.Win != a.Ws) return IBINN_EINVAL;
    if (a.pre == IBINN_PRE_BNBWD && (!a.x2 || !a.pre_sums)) return IBINN_ENULL;
    if ((a.pre == IBINN_PRE_BN_RELU || a.pre == IBINN_PRE_BNBWD) && (!a.pre_bn.mean || !a.pre_bn.scale || !a.pre_bn.gamma)) return IBINN_ENULL;
    if (a.pre == IBINN_PRE_BN_RELU && !a.pre_bn.beta) return IBINN_ENULL;
    if (a.epi == IBINN_EPI_BNSTATS) {
        if (!a.save_mean || !a.save_invstd || !a.partials || !a.counter) return IBINN_ENULL;
        if (a.Cout > 64 || a.bias || a.addend) return IBINN_EINVAL;
    } else if (a.epi == IBINN_EPI_MASKSTATS) {
        if (!a.m_h || !a.sums_out || !a.partials || !a.counter || !a.m_bn.mean || !a.m_bn.scale || !a.m_bn.gamma || !a.m_bn.beta) return IBINN_ENULL;
        if (a.Cout > 64 || a.bias || a.addend) return IBINN_EINVAL;
    } else if (a.epi == IBINN_EPI_RELUMASK) {
        if (!a.m_h) return IBINN_ENULL;
    } else if (a.epi != IBINN_EPI_STORE) return IBINN_EINVAL;
    return 0;
}

template <int KS, int PX, int S>
static int conv_launch(const ibinn_conv_args& a, const ConvGeom& g, cudaStream_t st) {
    auto kfn = conv_kernel<KS, PX, S>;
    int e = ibinn_set_smem(kfn, g.smem_bytes);
    if (e) return e;
    dim3 grid(g.n_ytiles * g.n_cotiles, a.B);
    IBINN_LAUNCH(kfn, grid, dim3(IBINN_THREADS), g.smem_bytes, st, a, g);
    return ibinn_launch_status();
}

}  // namespace

extern "C" size_t ibinn_sizeof_conv_args(void) { return sizeof(ibinn_conv_args); }

extern "C" int64_t ibinn_conv_wprep_floats(const ibinn_conv_args* a) {
    if (!a) return IBINN_ENULL;
    return conv_tc_supported(*a) ? conv_tc_wprep_floats(*a) : 0;
}

extern "C" int64_t ibinn_conv_partials_floats(const ibinn_conv_args* a) {
    if (!a) return IBINN_ENULL;
    if (a->epi != IBINN_EPI_BNSTATS && a->epi != IBINN_EPI_MASKSTATS) return 0;
    if (conv_tc_supported(*a)) return conv_tc_partials_floats(*a);
    ConvGeom g;
    if (!conv_pick(*a, g)) return IBINN_EINVAL;
    return (int64_t)g.n_ytiles * g.n_cotiles * a->B * 2 * a->Cout;
}

extern "C" int ibinn_conv_forward(const ibinn_conv_args* ap, ibinn_stream_t stream) {
    if (!ap) return IBINN_ENULL;
    int e = ibinn_check_device();
    if (e) return e;
    const ibinn_conv_args& a = *ap;
    e = conv_validate(a);
    if (e) return e;
    cudaStream_t st = (cudaStream_t)stream;
    if (conv_tc_supported(a)) return conv_tc_launch(a, st);      // tcgen05 implicit GEMM (stride 1)
    ConvGeom g;
    if (!conv_pick(a, g)) return IBINN_EINVAL;
#define IBINN_CONV_CASE(KS_, PX_, S_) \
    if (a.ksize == KS_ && g.px == PX_ && a.stride == S_) return conv_launch<KS_, PX_, S_>(a, g, st);
    IBINN_CONV_CASE(3, 4, 1) IBINN_CONV_CASE(3, 2, 1) IBINN_CONV_CASE(3, 1, 1)
    IBINN_CONV_CASE(3, 4, 2) IBINN_CONV_CASE(3, 2, 2) IBINN_CONV_CASE(3, 1, 2)
    IBINN_CONV_CASE(1, 4, 1) IBINN_CONV_CASE(1, 2, 1) IBINN_CONV_CASE(1, 1, 1)
#undef IBINN_CONV_CASE
    return IBINN_EINVAL;
}
