// Weight / bias gradients of the subnet convolutions (autograd backward of reference
// inn_architecture.py:77,81,86,97,101,105), exact fp32 on the CUDA cores.
//
//   dW[co][ci][ky][kx] += sum_{b,y,x} G[b,co,y,x] * X[b,ci, S*y+ky-pad, S*x+kx-pad]
//
// G and X are produced on the fly from what the forward saved: G = batch-norm backward of the
// incoming gradient (or the gradient itself), X = relu / relu(bn(.))*dropout of the conv input.
// Work decomposition: a thread owns an (8 co x 4 ci x 1 tap) register tile; a CTA owns up to 256
// such tiles times PP pixel partitions and loops over its images in row bands.  Both operands are
// staged channels-last ([pixel][C+4], conflict-free 128-bit transposing stores) so that the inner
// loop is three LDS.128 per 32 FMA.  Partial tiles are reduced across partitions in shared memory
// and accumulated to dW with one fp32 atomic per element per CTA.
#include "common.cuh"
#include "conv_internal.cuh"

namespace {

struct WgGeom {
    int nco8, nci4, ntiles, tc, pp, ntgroups, ipg, npixgroups, bh, xr, xw, cop, cip;
    size_t smem_bytes;
};

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

static bool wgrad_geometry(const ibinn_wgrad_args& a, WgGeom& g) {
    const int KK = a.ksize * a.ksize;
    g.nco8 = cdiv(a.Cout, 8);
    g.nci4 = cdiv(a.Cin, 4);
    g.ntiles = KK * g.nci4 * g.nco8;
    g.ntgroups = cdiv(g.ntiles, IBINN_THREADS);
    g.tc = cdiv(g.ntiles, g.ntgroups);
    g.pp = IBINN_THREADS / g.tc;
    if (g.pp < 1) g.pp = 1;
    g.cop = g.nco8 * 8 + 4;
    g.cip = g.nci4 * 4 + 4;
    // images per CTA: aim at >= ~148 CTAs
    int ipg = (int)(((long long)a.B * g.ntgroups) / 148);
    if (ipg < 1) ipg = 1;
    if (ipg > 8) ipg = 8;
    g.ipg = ipg;
    g.npixgroups = cdiv(a.B, ipg);
    // band height: fit the staging buffers in ~96 KB
    int bh = a.Hout;
    for (;;) {
        g.xr = (bh - 1) * a.stride + a.ksize;
        g.xw = (a.Wout - 1) * a.stride + a.ksize;
        size_t fl = (size_t)bh * a.Wout * g.cop + (size_t)g.xr * g.xw * g.cip + (size_t)bh * a.Wout;
        size_t red = (size_t)IBINN_THREADS * 33;
        if (fl < red) fl = red;
        g.smem_bytes = (fl + 3 * 256 + 64) * sizeof(float);
        if (g.smem_bytes <= 96 * 1024 || bh == 1) break;
        bh = (bh + 1) / 2;
    }
    g.bh = bh;
    return g.smem_bytes <= 200 * 1024;
}

__device__ __forceinline__ float wg_invstd(const ibinn_bn_ref& r, int c) {
    float s = r.scale[c];
    return r.scale_is_variance ? rsqrtf(s + r.eps) : s;
}

__global__ void __launch_bounds__(IBINN_THREADS) wgrad_kernel(const ibinn_wgrad_args p, const WgGeom g) {
    IBINN_DYN_SMEM(float, smem);
    const int KS = p.ksize, KK = KS * KS, PAD = KS / 2, S = p.stride;
    float* s_ga = smem;            // G pre-op coefficients (3 x 256)
    float* s_gb = s_ga + 256;
    float* s_gc = s_gb + 256;
    float* s_buf = s_gc + 256;     // staging / reduction area
    float* s_G = s_buf;
    float* s_X = s_G + (size_t)g.bh * p.Wout * g.cop;
    int* s_off = reinterpret_cast<int*>(s_X + (size_t)g.xr * g.xw * g.cip);

    const int tid = threadIdx.x;
    const int tl = tid % g.tc, pp = tid / g.tc;
    const int tile = blockIdx.x * g.tc + tl;
    const bool active = (pp < g.pp) && (tile < g.ntiles);
    const int co8 = tile % g.nco8;
    const int ci4 = (tile / g.nco8) % g.nci4;
    const int tap = tile / (g.nco8 * g.nci4);
    const int ky = tap / KS, kx = tap % KS;
    const bool want_bias = active && p.dbias && (tap == KK / 2) && (ci4 == 0);

    // coefficient tables.  X pre-op (BN_RELU) coefficients go to s_ga/s_gb too when G needs none.
    // G: BNBWD -> g_raw = a*g + b*h + c
    __shared__ float s_xa[256], s_xb[256];
    if (p.g_pre == IBINN_PRE_BNBWD) {
        for (int c = tid; c < p.Cout; c += IBINN_THREADS) {
            float is = wg_invstd(p.g_bn, c);
            float k = p.g_bn.gamma[c] * is;
            float m1 = p.g_sums[c] * p.g_inv_count;
            float m2 = p.g_sums[p.Cout + c] * p.g_inv_count;
            s_ga[c] = k;
            s_gb[c] = -k * m2 * is;
            s_gc[c] = -k * m1 + k * m2 * is * p.g_bn.mean[c];
        }
    }
    if (p.x_pre == IBINN_PRE_BN_RELU) {
        for (int c = tid; c < p.Cin; c += IBINN_THREADS) {
            float is = wg_invstd(p.x_bn, c);
            float A = p.x_bn.gamma[c] * is;
            s_xa[c] = A;
            s_xb[c] = p.x_bn.beta[c] - p.x_bn.mean[c] * A;
        }
    }

    float acc[8][4];
    float bsum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        bsum[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }

    const int b_begin = blockIdx.y * g.ipg;
    const int b_end = min(p.B, b_begin + g.ipg);
    const int nco4 = g.nco8 * 2, nci4 = g.nci4;
    const size_t HWo = (size_t)p.Hout * p.Wout, HWi = (size_t)p.Hin * p.Win;

    for (int b = b_begin; b < b_end; ++b) {
        for (int yb = 0; yb < p.Hout; yb += g.bh) {
            const int rows = min(g.bh, p.Hout - yb);
            const int npx = rows * p.Wout;
            __syncthreads();
            // ---- stage G: [q][co], 4 channels per thread, 128-bit stores
            for (int e = tid; e < npx * nco4; e += IBINN_THREADS) {
                const int q = e % npx, c4 = e / npx;
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = c4 * 4 + k;
                    float t = 0.f;
                    if (c < p.Cout) {
                        const size_t off = (size_t)c * HWo + (size_t)yb * p.Wout + q;
                        t = __ldg(p.g + (size_t)b * p.g_bs + off);
                        if (p.g_pre == IBINN_PRE_BNBWD)
                            t = fmaf(s_ga[c], t, fmaf(s_gb[c], __ldg(p.gh + (size_t)b * p.gh_bs + off), s_gc[c]));
                    }
                    v[k] = t;
                }
                st4(s_G + (size_t)q * g.cop + c4 * 4, make_float4(v[0], v[1], v[2], v[3]));
            }
            // ---- stage X: [r][col][ci] with halo / zero padding
            const int xrows = (rows - 1) * S + KS;
            const int nxp = xrows * g.xw;
            const int iy0 = yb * S - PAD;
            for (int e = tid; e < nxp * nci4; e += IBINN_THREADS) {
                const int q = e % nxp, c4 = e / nxp;
                const int r = q / g.xw, col = q - r * g.xw;
                const int iy = iy0 + r, ix = col - PAD;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int c = c4 * 4 + k;
                        if (c < p.Cin) {
                            float t = __ldg(p.x + (size_t)b * p.x_bs + (size_t)c * HWi + (size_t)iy * p.Win + ix);
                            if (p.x_pre == IBINN_PRE_RELU) t = fmaxf(t, 0.f);
                            else if (p.x_pre == IBINN_PRE_BN_RELU) {
                                t = fmaxf(fmaf(t, s_xa[c], s_xb[c]), 0.f);
                                if (p.x_mask) t *= __ldg(p.x_mask + (size_t)b * p.Cin + c);
                            }
                            v[k] = t;
                        }
                    }
                }
                st4(s_X + (size_t)q * g.cip + c4 * 4, make_float4(v[0], v[1], v[2], v[3]));
            }
            for (int q = tid; q < npx; q += IBINN_THREADS) {
                const int r = q / p.Wout, c = q - r * p.Wout;
                s_off[q] = (r * S) * g.xw + c * S;
            }
            __syncthreads();
            if (active) {
                const float* gp = s_G + co8 * 8;
                const float* xp = s_X + (size_t)(ky * g.xw + kx) * g.cip + ci4 * 4;
                for (int q = pp; q < npx; q += g.pp) {
                    const float4 g0 = ld4(gp + (size_t)q * g.cop), g1 = ld4(gp + (size_t)q * g.cop + 4);
                    const float4 xv = ld4(xp + (size_t)s_off[q] * g.cip);
                    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                    const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], xx[j], acc[i][j]);
                    if (want_bias) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) bsum[i] += gv[i];
                    }
                }
            }
        }
    }

    // ---- reduce the PP pixel partitions, then accumulate to global
    __syncthreads();
    float* s_r = s_buf;   // [tid][33]
    if (g.pp > 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s_r[tid * 33 + i * 4 + j] = active ? acc[i][j] : 0.f;
        __syncthreads();
        if (pp == 0 && active) {
            for (int q = 1; q < g.pp; ++q)
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] += s_r[(q * g.tc + tl) * 33 + i * 4 + j];
        }
        __syncthreads();
        if (p.dbias) {
#pragma unroll
            for (int i = 0; i < 8; ++i) s_r[tid * 33 + i] = want_bias ? bsum[i] : 0.f;
            __syncthreads();
            if (pp == 0 && want_bias) {
                for (int q = 1; q < g.pp; ++q)
#pragma unroll
                    for (int i = 0; i < 8; ++i) bsum[i] += s_r[(q * g.tc + tl) * 33 + i];
            }
        }
    }
    if (pp == 0 && active) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int co = co8 * 8 + i;
            if (co >= p.Cout) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = ci4 * 4 + j;
                if (ci < p.Cin) atomicAdd(p.dw + ((size_t)co * p.Cin + ci) * KK + tap, acc[i][j]);
            }
            if (want_bias) atomicAdd(p.dbias + co, bsum[i]);
        }
    }
}

}  // namespace

extern "C" size_t ibinn_sizeof_wgrad_args(void) { return sizeof(ibinn_wgrad_args); }

extern "C" int64_t ibinn_conv_wgrad_workspace_floats(const ibinn_wgrad_args* a) {
    if (!a) return IBINN_ENULL;
    return wgrad_tc_supported(*a) ? wgrad_tc_workspace_floats(*a) : 0;
}

extern "C" int ibinn_conv_wgrad(const ibinn_wgrad_args* ap, ibinn_stream_t stream) {
    if (!ap) return IBINN_ENULL;
    int e = ibinn_check_device();
    if (e) return e;
    const ibinn_wgrad_args& a = *ap;
    if (!a.g || !a.x || !a.dw) return IBINN_ENULL;
    if (a.B <= 0 || a.Cin <= 0 || a.Cout <= 0 || a.Cin > 256 || a.Cout > 256) return IBINN_EINVAL;
    if (!(a.ksize == 1 || a.ksize == 3) || !(a.stride == 1 || a.stride == 2)) return IBINN_EINVAL;
    const int pad = a.ksize / 2;
    if (a.Hout != (a.Hin + 2 * pad - a.ksize) / a.stride + 1 || a.Wout != (a.Win + 2 * pad - a.ksize) / a.stride + 1) return IBINN_EINVAL;
    if (a.g_pre == IBINN_PRE_BNBWD) {
        if (!a.gh || !a.g_sums || !a.g_bn.mean || !a.g_bn.scale || !a.g_bn.gamma) return IBINN_ENULL;
    } else if (a.g_pre != IBINN_PRE_NONE) return IBINN_EINVAL;
    if (a.x_pre == IBINN_PRE_BN_RELU) {
        if (!a.x_bn.mean || !a.x_bn.scale || !a.x_bn.gamma || !a.x_bn.beta) return IBINN_ENULL;
    } else if (a.x_pre != IBINN_PRE_NONE && a.x_pre != IBINN_PRE_RELU) return IBINN_EINVAL;
    if (wgrad_tc_supported(a)) return wgrad_tc_launch(a, a.workspace, (cudaStream_t)stream);     // tcgen05 path (stride 1)
    WgGeom g;
    if (!wgrad_geometry(a, g)) return IBINN_EINVAL;
    if (g.npixgroups > 65535) return IBINN_EINVAL;
    e = ibinn_set_smem(wgrad_kernel, g.smem_bytes);
    if (e) return e;
    dim3 grid(g.ntgroups, g.npixgroups);
    IBINN_LAUNCH(wgrad_kernel, grid, dim3(IBINN_THREADS), g.smem_bytes, (cudaStream_t)stream, a, g);
    return ibinn_launch_status();
}
