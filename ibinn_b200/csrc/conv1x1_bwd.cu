// Backward of the subnets' 1x1 output convolution (reference inn_architecture.py:88,106: `nn.Conv2d(width, cout, 1)` on
// relu(bn2(h2)) [* Dropout2d mask]) in ONE launch: what ibinn_conv_forward (dgrad, EPI_MASKSTATS) and ibinn_conv_wgrad did in two
// launches on two streams.  Per tile of PPC consecutive (image, pixel) positions:
//   g_in[w]  = (sum_j W[j][w] g[j]) * [bn(h)[w] > 0] * drop[b][w]          gradient w.r.t. the BatchNorm output, masked
//   sums     = (sum g_in, sum g_in * xhat)  per channel                      BatchNorm-backward sums (dbeta, dgamma): per-CTA records,
//                                                                            combined in fixed order by the last CTA
//   dW[j][w] = sum_pos g[j] * relu(bn(h))[w] * drop,  dbias[j] = sum g[j]    per-CTA records in the layout of ibinn_wgrad_reduce_many
// Exact fp32 on the CUDA cores: the three contractions are 2 c2 x width (12 x 32 ... 48 x 64) per position and this kernel is bound
// by its launch / load / ticket latencies, not by arithmetic.
#include "common.cuh"

namespace {

struct C1P {
    ibinn_conv1x1_bwd_args a;
    int ppc, ncta, rec_floats, total;      // positions per CTA, grid, floats per dW record, B * HW
};

static bool c1_geometry(const ibinn_conv1x1_bwd_args& a, C1P& p) {
    if (a.B <= 0 || a.Cin <= 0 || a.Cout <= 0 || a.Cout > 64 || (a.Cin != 16 && a.Cin != 32 && a.Cin != 64)) return false;      // the subnet widths
    const int HW = a.H * a.W;
    if (HW & 3) return false;
    if ((int64_t)a.B * HW >= (1ll << 31)) return false;
    p.a = a;
    p.total = a.B * HW;
    p.ppc = HW >= 1024 ? 256 : HW >= 256 ? 128 : 64;
    p.ncta = (p.total + p.ppc - 1) / p.ppc;
    p.rec_floats = (a.Cout * a.Cin + a.Cout + 3) & ~3;
    return true;
}

__global__ void __launch_bounds__(IBINN_THREADS) conv1x1_bwd_kernel(const C1P p) {
    IBINN_DYN_SMEM(float, sm);
    __shared__ int s_flag;
    const ibinn_conv1x1_bwd_args& a = p.a;
    const int Cin = a.Cin, Cout = a.Cout, HW = a.H * a.W, PPC = p.ppc, nq = PPC / 4;
    float* s_g = sm;                               // [Cout][PPC]   upstream gradient
    float* s_act = s_g + Cout * PPC;               // [Cin][PPC]    relu(bn(h)) * drop
    float* s_xh = s_act + Cin * PPC;               // [Cin][PPC]    xhat
    float* s_w = s_xh + Cin * PPC;                 // [Cout][Cin]
    float* s_A = s_w + Cout * Cin;                 // [Cin] gamma * invstd
    float* s_Bc = s_A + Cin;                       // [Cin] beta - mean * A
    float* s_mean = s_Bc + Cin;                    // [Cin]
    float* s_is = s_mean + Cin;                    // [Cin]
    float* s_actT = s_is + Cin;                    // [PPC][Cin + 1]  the activation again, position-major (weight-gradient phase)
    const int TP1 = Cin + 1;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int e0 = blockIdx.x * PPC;               // first (image, pixel) position of this CTA

    for (int c = tid; c < Cin; c += IBINN_THREADS) {
        const float sc = a.bn.scale[c];
        const float is = a.bn.scale_is_variance ? rsqrtf(sc + a.bn.eps) : sc;
        const float A = a.bn.gamma[c] * is, mean = a.bn.mean[c];
        s_A[c] = A; s_Bc[c] = a.bn.beta[c] - mean * A; s_mean[c] = mean; s_is[c] = is;
    }
    for (int i = tid; i < Cout * Cin; i += IBINN_THREADS) s_w[i] = a.w[i];
    __syncthreads();
    // ---- phase 1: stage g and the activation of this tile (float4 along the pixels)
    for (int i = tid; i < (Cout + Cin) * nq; i += IBINN_THREADS) {
        const int c = i / nq, q = i - c * nq;
        const int e = e0 + 4 * q;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < Cout) {
            if (e < p.total) {
                const int b = e / HW, px = e - b * HW;
                v = *reinterpret_cast<const float4*>(a.g + ((size_t)b * Cout + c) * HW + px);
            }
            *reinterpret_cast<float4*>(s_g + c * PPC + 4 * q) = v;
        } else {
            const int w = c - Cout;
            float4 act = v, xh = v;
            if (e < p.total) {
                const int b = e / HW, px = e - b * HW;
                const float4 h = *reinterpret_cast<const float4*>(a.h + (size_t)b * a.h_bs + (size_t)w * HW + px);
                const float dm = a.drop ? a.drop[(size_t)b * Cin + w] : 1.f;
                const float A = s_A[w], Bc = s_Bc[w], mean = s_mean[w], is = s_is[w];
                act = make_float4(fmaxf(fmaf(h.x, A, Bc), 0.f) * dm, fmaxf(fmaf(h.y, A, Bc), 0.f) * dm, fmaxf(fmaf(h.z, A, Bc), 0.f) * dm,
                                  fmaxf(fmaf(h.w, A, Bc), 0.f) * dm);
                // the decision of the forward pass (pre-activation > 0) is kept in the sign bit of xhat's companion: act > 0 <=> kept
                // unless the dropout mask is zero, where the gradient vanishes as well
                xh = make_float4((h.x - mean) * is, (h.y - mean) * is, (h.z - mean) * is, (h.w - mean) * is);
                if (dm == 0.f) act = make_float4(0.f, 0.f, 0.f, 0.f);
                // keep the ReLU decision separately for dm > 0: act > 0 exactly when the pre-activation was > 0
            }
            *reinterpret_cast<float4*>(s_act + w * PPC + 4 * q) = act;
            *reinterpret_cast<float4*>(s_xh + w * PPC + 4 * q) = xh;
            s_actT[(4 * q + 0) * TP1 + w] = act.x; s_actT[(4 * q + 1) * TP1 + w] = act.y;
            s_actT[(4 * q + 2) * TP1 + w] = act.z; s_actT[(4 * q + 3) * TP1 + w] = act.w;
        }
    }
    __syncthreads();
    // ---- phase 2: g_in = (W^T g) * mask * drop, its store and the BatchNorm-backward sums.  warp <-> channel w, lane <-> pixel quad
    for (int w = wid; w < Cin; w += IBINN_THREADS / 32) {
        float s1 = 0.f, s2 = 0.f;
        for (int q = lane; q < nq; q += 32) {
            const int e = e0 + 4 * q;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
            for (int j = 0; j < Cout; ++j) {
                const float wj = s_w[j * Cin + w];
                const float4 gv = *reinterpret_cast<const float4*>(s_g + j * PPC + 4 * q);
                acc[0] = fmaf(wj, gv.x, acc[0]); acc[1] = fmaf(wj, gv.y, acc[1]); acc[2] = fmaf(wj, gv.z, acc[2]); acc[3] = fmaf(wj, gv.w, acc[3]);
            }
            if (e < p.total) {
                const int b = e / HW, px = e - b * HW;
                const float dm = a.drop ? a.drop[(size_t)b * Cin + w] : 1.f;
                const float4 act = *reinterpret_cast<const float4*>(s_act + w * PPC + 4 * q);
                const float4 xh = *reinterpret_cast<const float4*>(s_xh + w * PPC + 4 * q);
                float o[4];
                o[0] = act.x > 0.f ? acc[0] * dm : 0.f; o[1] = act.y > 0.f ? acc[1] * dm : 0.f;
                o[2] = act.z > 0.f ? acc[2] * dm : 0.f; o[3] = act.w > 0.f ? acc[3] * dm : 0.f;
                *reinterpret_cast<float4*>(a.g_in + ((size_t)b * Cin + w) * HW + px) = make_float4(o[0], o[1], o[2], o[3]);
                s1 += (o[0] + o[1]) + (o[2] + o[3]);
                s2 += fmaf(o[0], xh.x, fmaf(o[1], xh.y, fmaf(o[2], xh.z, o[3] * xh.w)));
            }
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) {
            a.partials[(size_t)blockIdx.x * 2 * Cin + w] = s1;
            a.partials[(size_t)blockIdx.x * 2 * Cin + Cin + w] = s2;
        }
    }
    // ---- phase 3: weight-gradient record of this tile.  thread <-> input channel w (consecutive lanes: conflict-free reads of the
    // position-major activation), thread groups split the output channels j (their g values are warp-wide broadcasts); every
    // thread sums over the tile's positions itself: no cross-lane reduction
    {
        float* rec = a.wrecords + (size_t)blockIdx.x * p.rec_floats;
        const int ngrp = IBINN_THREADS / Cin, w = tid % Cin, grp = tid / Cin;       // Cin in {4..64}, a multiple of 4 (and IBINN_THREADS % Cin == 0 for 16 / 32 / 64)
        if (grp < ngrp) {
            for (int j0 = grp; j0 < Cout; j0 += ngrp * 8) {
                float acc[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[u] = 0.f;
                for (int px = 0; px < PPC; ++px) {
                    const float av = s_actT[px * TP1 + w];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int j = j0 + u * ngrp;
                        if (j < Cout) acc[u] = fmaf(s_g[j * PPC + px], av, acc[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = j0 + u * ngrp;
                    if (j < Cout) rec[j * Cin + w] = acc[u];
                }
            }
        }
        // bias gradient: sum over the tile's positions of g[j]
        for (int j = wid; j < Cout; j += IBINN_THREADS / 32) {
            float sb = 0.f;
            for (int px = lane; px < PPC; px += 32) sb += s_g[j * PPC + px];
            sb = warp_sum(sb);
            if (lane == 0) rec[Cout * Cin + j] = sb;
        }
    }
    // ---- last CTA: BatchNorm-backward sums over all tiles, fixed order
    if (!last_block_ticket(a.counter, gridDim.x, &s_flag)) return;
    // (all 256 threads: channel c = tid % 2Cin, record slice = tid / 2Cin, eight loads in flight; a serial loop per channel cost
    //  one L2 round trip per four records: 20 - 40 us at 256 - 512 tiles)
    {
        const int nch = 2 * Cin, nsl = IBINN_THREADS / nch, c = tid % nch, sl = tid / nch;
        float acc = 0.f;
        if (sl < nsl) {
            for (int r0 = sl; r0 < (int)gridDim.x; r0 += nsl * 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int r = r0 + u * nsl;
                    v[u] = r < (int)gridDim.x ? __ldcg(a.partials + (size_t)r * nch + c) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += v[u];
            }
        }
        __syncthreads();
        float* s_red = sm;                          // the operand tiles are dead by now
        if (sl < nsl) s_red[sl * nch + c] = acc;
        __syncthreads();
        if (tid < nch) {
            float t = 0.f;
            for (int q = 0; q < nsl; ++q) t += s_red[q * nch + tid];
            a.sums_out[tid] = t;
            if (tid < Cin) { if (a.dbeta) a.dbeta[tid] += t; }
            else if (a.dgamma) a.dgamma[tid - Cin] += t;
        }
    }
    if (tid == 0) *a.counter = 0u;
}

static size_t c1_smem(const C1P& p) {
    return (size_t)((p.a.Cout + 2 * p.a.Cin) * p.ppc + p.a.Cout * p.a.Cin + 4 * p.a.Cin + (p.a.Cin + 1) * p.ppc + 16) * sizeof(float);
}

}  // namespace

extern "C" size_t ibinn_sizeof_conv1x1_bwd_args(void) { return sizeof(ibinn_conv1x1_bwd_args); }

extern "C" int ibinn_conv1x1_bwd_geometry(const ibinn_conv1x1_bwd_args* a, int32_t* ncta, int32_t* rec_floats) {
    if (!a || !ncta || !rec_floats) return IBINN_ENULL;
    C1P p;
    if (!c1_geometry(*a, p)) return IBINN_EINVAL;
    *ncta = p.ncta;
    *rec_floats = p.rec_floats;
    return 0;
}

extern "C" int ibinn_conv1x1_backward(const ibinn_conv1x1_bwd_args* ap, ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!ap) return IBINN_ENULL;
    const ibinn_conv1x1_bwd_args& a = *ap;
    if (!a.g || !a.h || !a.w || !a.g_in || !a.sums_out || !a.partials || !a.counter || !a.wrecords || !a.bn.mean || !a.bn.scale || !a.bn.gamma || !a.bn.beta)
        return IBINN_ENULL;
    C1P p;
    if (!c1_geometry(a, p)) return IBINN_EINVAL;
    if (((reinterpret_cast<uintptr_t>(a.g) | reinterpret_cast<uintptr_t>(a.h) | reinterpret_cast<uintptr_t>(a.g_in)) & 15) != 0 || (a.h_bs & 3)) return IBINN_EALIGN;
    const size_t smem = c1_smem(p);
    e = ibinn_set_smem(conv1x1_bwd_kernel, smem);
    if (e) return e;
    IBINN_LAUNCH(conv1x1_bwd_kernel, dim3(p.ncta), dim3(IBINN_THREADS), smem, (cudaStream_t)stream, p);
    return ibinn_launch_status();
}
