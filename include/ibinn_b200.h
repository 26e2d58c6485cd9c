/* ibinn_b200 -- C ABI of the B200-native IB-INN hot path (sm_100a only).
 *
 * The reference (vislearn/IB-INN) is pure Python/PyTorch and has no FFI of its own; each entry
 * point below replaces the stock ATen/cuDNN/cuBLAS calls that the cited reference lines issue.
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer on the current device, fp32 unless noted, contiguous NCHW
 *     within an image; `*_bs` is the batch stride in floats (lets a kernel address a channel slice
 *     of a wider tensor without a copy);
 *   - the caller owns all buffers, including workspaces; kernels never allocate, free or sync;
 *   - all work is enqueued on `stream` (a cudaStream_t); safe under CUDA-graph capture;
 *   - return 0 on success, a negative IBINN_E* for a rejected call (checked before launch),
 *     a positive value = cudaError_t of the launch.  Nothing throws or exits;
 *   - `counter` arguments point at one zero-initialised uint32 that the kernel returns to zero;
 *   - there is no CPU fallback: on a device that is not sm_100 every call returns IBINN_EARCH.
 */
#ifndef IBINN_B200_H
#define IBINN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IBINN_OK 0
#define IBINN_EINVAL (-1)   /* bad shape / unsupported size */
#define IBINN_EALIGN (-2)   /* pointer or stride not aligned as required */
#define IBINN_EARCH (-3)    /* device is not sm_100 (B200) */
#define IBINN_ENULL (-4)    /* required pointer is NULL */

typedef void* ibinn_stream_t; /* cudaStream_t */

int ibinn_abi_version(void);
/* text of the most recent failure reported by this library ("" if none); not thread-safe, diagnostics only */
const char* ibinn_last_error(void);
/* kernels enqueued by this library in this process so far (host-side counter) */
int64_t ibinn_launch_count(void);
/* 0 if the current device is sm_100, else IBINN_EARCH (or the cudaError_t). */
int ibinn_check_device(void);

/* ---------------------------------------------------------------- subnet convolutions
 * Replaces nn.Conv2d / nn.BatchNorm2d / nn.ReLU / nn.Dropout2d of basic_residual_block and
 * strided_residual_block (reference inn_architecture.py:68-110) and their autograd backward.
 * One kernel = [pre-op on the input while it is staged] -> KSxKS convolution (pad KS/2,
 * stride 1|2, no groups) -> [epilogue].  dgrad is the same kernel with `w_transposed`.
 */
enum {
    IBINN_PRE_NONE = 0,
    IBINN_PRE_RELU = 1,     /* relu(x)                                     inn_architecture.py:72 */
    IBINN_PRE_BN_RELU = 2,  /* relu(bn(x)) [* dropout mask (B,Cin)]        inn_architecture.py:78-84 */
    IBINN_PRE_BNBWD = 3     /* batch-norm backward applied to (x=grad, x2=raw fwd activation) */
};
enum {
    IBINN_EPI_STORE = 0,     /* y = conv (+bias) (+addend) */
    IBINN_EPI_BNSTATS = 1,   /* y = conv; batch mean / inv-std of y -> save_*, running stats updated */
    IBINN_EPI_MASKSTATS = 2, /* y = conv * [bn(m_h) > 0] (* drop mask); sums_out = (sum y, sum y*xhat) */
    IBINN_EPI_RELUMASK = 3   /* y = conv * [m_h > 0] (+addend) */
};

typedef struct ibinn_bn_ref {  /* a batch-norm layer as seen by a kernel */
    const float* mean;         /* (C) batch or running mean */
    const float* scale;        /* (C) inv-std, or the variance if var_is_variance */
    const float* gamma;        /* (C) */
    const float* beta;         /* (C) */
    float eps;
    int32_t scale_is_variance;
} ibinn_bn_ref;

typedef struct ibinn_conv_args {
    const float* x;  int64_t x_bs;      /* input  (B, Cin, Hs, Ws)                     */
    const float* x2; int64_t x2_bs;     /* PRE_BNBWD: raw forward activation, same geometry */
    const float* w;                     /* (Cout,Cin,KS,KS); w_transposed: (Cin,Cout,KS,KS) read flipped */
    const float* bias;                  /* (Cout) or NULL */
    const float* addend; int64_t addend_bs; /* (B,Cout,Hout,Wout) or NULL; may alias y */
    float* y; int64_t y_bs;             /* output (B, Cout, Hout, Wout) */
    int32_t B, Cin, Cout, Hs, Ws, Hout, Wout;
    int32_t ksize, stride;              /* ksize 1|3, stride 1|2 */
    int32_t upsample2;                  /* input is zero-inserted to (2Hs [-1], 2Ws [-1]) first: dgrad of a stride-2 conv */
    int32_t Hin, Win;                   /* logical input size (== Hs,Ws unless upsample2) */
    int32_t w_transposed;
    int32_t pre, epi;
    ibinn_bn_ref pre_bn;                /* PRE_BN_RELU, PRE_BNBWD */
    const float* pre_mask;              /* PRE_BN_RELU: (B,Cin) dropout mask (0 or 1/(1-p)) or NULL */
    const float* pre_sums;              /* PRE_BNBWD: (2,Cin) = sum g, sum g*xhat */
    float pre_inv_count;                /* PRE_BNBWD: 1 / (B*Hs*Ws) */
    /* EPI_BNSTATS */
    float* save_mean; float* save_invstd;        /* (Cout) outputs */
    float* running_mean; float* running_var;     /* (Cout) in/out or NULL */
    int64_t* num_batches_tracked;                /* in/out or NULL */
    float momentum;                              /* < 0: cumulative average 1/num_batches_tracked */
    float bn_eps;
    /* EPI_MASKSTATS / EPI_RELUMASK */
    const float* m_h; int64_t m_bs;              /* activation that decides the mask (B,Cout,Hout,Wout) */
    ibinn_bn_ref m_bn;                           /* MASKSTATS: its batch norm */
    const float* m_drop;                         /* (B,Cout) dropout mask or NULL */
    float* sums_out;                             /* (2,Cout) */
    float* m_dgamma; float* m_dbeta;             /* (Cout) accumulated: += sums_out[1], += sums_out[0]; or NULL */
    /* workspaces */
    float* partials;                             /* ibinn_conv_partials_floats() floats (BNSTATS, MASKSTATS) */
    uint32_t* counter;
    float* wprep;                                /* ibinn_conv_wprep_floats() floats, 16-byte aligned: tensor-core operand image of w */
    int32_t wprep_ready;                         /* 1: `wprep` already holds the image of `w` (ibinn_wprep_many), skip the prep launch */
} ibinn_conv_args;

size_t ibinn_sizeof_conv_args(void);
/* number of floats `partials` must hold for this call (0 if the epilogue needs none) */
int64_t ibinn_conv_partials_floats(const ibinn_conv_args* a);
/* number of floats `wprep` must hold (0: this call runs on the CUDA-core kernel and needs none) */
int64_t ibinn_conv_wprep_floats(const ibinn_conv_args* a);
int ibinn_conv_forward(const ibinn_conv_args* a, ibinn_stream_t stream);

/* Batched filter preparation for the tensor-core convolutions: one launch turns n raw filter tensors into
 * their operand images (hi/lo tf32 split, per-tap K-major rows).  `table` is a DEVICE array. */
typedef struct ibinn_wprep_desc {
    const float* w;      /* (Cout,Cin,KS,KS), or (Cin,Cout,KS,KS) read flipped when transposed */
    float* out;          /* ibinn_conv_wprep_floats() floats */
    int32_t Cin, Cout, ksize, transposed;
} ibinn_wprep_desc;
int ibinn_wprep_many(const ibinn_wprep_desc* table, int n, ibinn_stream_t stream);

typedef struct ibinn_wgrad_args {
    const float* g;  int64_t g_bs;      /* grad w.r.t. conv output (B,Cout,Hout,Wout)        */
    const float* gh; int64_t gh_bs;     /* g_pre == PRE_BNBWD: raw conv output of the forward */
    int32_t g_pre;                      /* IBINN_PRE_NONE | IBINN_PRE_BNBWD */
    ibinn_bn_ref g_bn; const float* g_sums; float g_inv_count;
    const float* x;  int64_t x_bs;      /* conv input before its pre-op (B,Cin,Hin,Win) */
    int32_t x_pre;                      /* NONE | RELU | BN_RELU */
    ibinn_bn_ref x_bn; const float* x_mask;
    float* dw;                          /* (Cout,Cin,KS,KS), ACCUMULATED into */
    float* dbias;                       /* (Cout) accumulated, or NULL */
    int32_t B, Cin, Cout, Hin, Win, Hout, Wout, ksize, stride;
    float* workspace;                   /* ibinn_conv_wgrad_workspace_floats() floats (per-CTA partial records), or NULL if 0 */
    int32_t defer_reduce;               /* 1: leave the partial records in `workspace`; the caller sums them later with ibinn_wgrad_reduce_many */
} ibinn_wgrad_args;

size_t ibinn_sizeof_wgrad_args(void);
/* floats of `workspace` this call needs (0: CUDA-core kernel, accumulates with atomics) */
int64_t ibinn_conv_wgrad_workspace_floats(const ibinn_wgrad_args* a);
int ibinn_conv_wgrad(const ibinn_wgrad_args* a, ibinn_stream_t stream);
/* Deferred reduction of the tensor-core wgrad records of many layers in ONE launch.  `ibinn_conv_wgrad_records`
 * tells how the records of a call are laid out; `table` is a DEVICE array. */
typedef struct ibinn_wreduce_desc {
    const float* partials; float* dw; float* dbias;
    int32_t ncta, rec_floats, Cout, Cin, kk;
} ibinn_wreduce_desc;
int ibinn_conv_wgrad_records(const ibinn_wgrad_args* a, int32_t* ncta, int32_t* rec_floats);
int ibinn_wgrad_reduce_many(const ibinn_wreduce_desc* table, int n, ibinn_stream_t stream);

/* ---------------------------------------------------------------- AIO_Block coupling
 * Replaces reference all_in_one_block.py:61-111 (`log_e`, `affine`, `permute`, `forward`):
 *   x1|x2 = split(x, [C - C/2, C/2]);  sig = clamp*tanh(alpha*a[:, :C/2]);  y2 = x2*exp(sig) + a[:, C/2:]
 *   out = (w . [x1|y2]) * exp(act_norm) + act_offset ;  jac[b] = sum(sig) + H*W*sum(act_norm)   (:113-114)
 * all tensors contiguous; x,out (B,C,H,W); a (B,2*(C/2),H,W); w (C,C) as [out][in]; one CTA per sample. */
int ibinn_aio_coupling_forward(const float* x, const float* a, const float* w, const float* act_norm,
                               const float* act_offset, float* out, float* jac, int B, int C, int H, int W,
                               float clamp, float alpha, ibinn_stream_t stream);
/* hand-derived backward.  `out` is the forward output.  g_act_norm /
 * g_act_offset (C) are accumulated into.  A sample is processed by ibinn_aio_backward_records(...)/B CTAs; each CTA
 * writes one record of 2*C floats (sum g, sum g*(y - offset) + pixels*g_jac) to `partials`
 * (ibinn_aio_backward_records(B,C,H,W) * 2*C floats).  defer_reduce = 0: the last CTA adds the records to g_act_offset /
 * g_act_norm; 1: the caller does it for many blocks at once with ibinn_aio_param_reduce_many. */
int ibinn_aio_backward_records(int B, int C, int H, int W);
int ibinn_aio_coupling_backward(const float* g_out, const float* g_jac, const float* x, const float* a,
                                const float* out, const float* w, const float* act_norm, const float* act_offset,
                                float* g_x, float* g_a, float* g_act_norm, float* g_act_offset, float* partials,
                                uint32_t* counter, int B, int C, int H, int W, float clamp, float alpha,
                                int defer_reduce, ibinn_stream_t stream);
typedef struct {
    const float* partials;   /* nrec records of 2*C floats */
    float* g_act_offset;     /* (C), accumulated into; may be NULL */
    float* g_act_norm;       /* (C), accumulated into; may be NULL */
    int32_t nrec, C;
} ibinn_aio_reduce_desc;
/* g_act_offset[c] += sum_r partials[r][c],  g_act_norm[c] += sum_r partials[r][C + c]  for every entry of the table
 * (device memory, n entries), records summed in index order. */
int ibinn_aio_param_reduce_many(const ibinn_aio_reduce_desc* table, int n, ibinn_stream_t stream);
/* Backward of the subnets' 1x1 output convolution (inn_architecture.py:88,106) in one launch: the masked input gradient with its
 * BatchNorm-backward sums (what ibinn_conv_forward with w_transposed + EPI_MASKSTATS computes) AND the weight / bias gradient
 * records (what ibinn_conv_wgrad computes), exact fp32.  g (B,Cout,H,W) contiguous; h (B,Cin,H,W) raw input of the preceding
 * BatchNorm, batch stride h_bs; w (Cout,Cin); g_in (B,Cin,H,W) contiguous out; sums_out (2,Cin); dgamma / dbeta (Cin) accumulated or
 * NULL; partials: ncta * 2 * Cin floats; wrecords: ncta * rec_floats floats, to be summed into dw / dbias by ibinn_wgrad_reduce_many
 * (ibinn_wreduce_desc with kk = 1); ncta / rec_floats from ibinn_conv1x1_bwd_geometry.  Needs H*W % 4 == 0, Cin % 4 == 0, <= 64 channels. */
typedef struct ibinn_conv1x1_bwd_args {
    const float* g; const float* h; int64_t h_bs;
    ibinn_bn_ref bn;
    const float* drop;          /* (B,Cin) Dropout2d mask or NULL */
    const float* w;
    float* g_in; float* sums_out; float* dgamma; float* dbeta;
    float* partials; uint32_t* counter; float* wrecords;
    int32_t B, Cin, Cout, H, W;
} ibinn_conv1x1_bwd_args;
size_t ibinn_sizeof_conv1x1_bwd_args(void);
int ibinn_conv1x1_bwd_geometry(const ibinn_conv1x1_bwd_args* a, int32_t* ncta, int32_t* rec_floats);
int ibinn_conv1x1_backward(const ibinn_conv1x1_bwd_args* a, ibinn_stream_t stream);

/* ---------------------------------------------------------------- fused AIO_Block forward (one launch for a RUN of blocks)
 * Replaces, for a run of `nblk` consecutive AIO_Blocks of one resolution stage, everything the entry points above do in
 * 4 launches per block (reference all_in_one_block.py:83-114 + inn_architecture.py:68-92): conv3x3 -> BN -> ReLU -> conv3x3 ->
 * BN -> ReLU -> [Dropout2d] -> conv1x1 -> affine coupling -> soft permutation -> ActNorm.  One CTA owns one image and walks the
 * whole run with the activations resident in shared memory; the four contractions of a block (incl. the soft permutation) are
 * 3xTF32 tcgen05 MMAs into one TMEM region.  train = 0 (running statistics): no inter-CTA dependency, persistent over images.
 * train = 1 (batch statistics): two grid-wide barriers per block (cooperative launch, B <= number of SMs), the raw conv
 * outputs, subnet output, block output and batch statistics are saved for the backward entry points.
 * `blocks` is a DEVICE array of nblk descriptors; filter images come from ibinn_wprep_many (conv3: the (2c2,width,1,1) filter;
 * perm: the (C,C,1,1) matrix w as a 1x1 filter). */
typedef struct ibinn_aio_block_desc {
    const float* wprep1; const float* wprep2; const float* wprep3; const float* wprep_perm;
    const float* bias3;                 /* (2*(C/2)) */
    ibinn_bn_ref bn1, bn2;              /* train = 0: running mean / variance (scale_is_variance = 1); train = 1: gamma, beta, eps */
    const float* act_norm; const float* act_offset;   /* (C) */
    const float* drop_mask;             /* train: (B,width) Dropout2d mask (0 or 1/(1-p)) or NULL */
    float* out;                         /* (B,C,H,W) block output, or NULL (train = 0: only needed for the last block of the run) */
    float* h1; float* h2; float* a;     /* train: (B,width,H,W) x2 raw conv outputs, (B,2*(C/2),H,W) subnet output */
    float* save;                        /* train: (4,width) = batch mean1, inv-std1, mean2, inv-std2 */
    float* running_mean1; float* running_var1; int64_t* nbt1;   /* train: updated in place (may be NULL) */
    float* running_mean2; float* running_var2; int64_t* nbt2;
    int32_t relu_first;                 /* leading ReLU of the subnet (inn_architecture.py:72) */
    int32_t reserved;
} ibinn_aio_block_desc;

typedef struct ibinn_aio_stage_args {
    const float* x;                     /* (B,C,H,W) contiguous: input of the first block of the run */
    float* jac;                         /* (B): sum of the log-dets of the run's blocks */
    const ibinn_aio_block_desc* blocks; /* device memory */
    int32_t nblk, B, C, H, W, width, train;
    float clamp, alpha, momentum;       /* momentum < 0: cumulative average */
    float* stat_records;                /* train: ibinn_aio_stage_workspace_floats() floats */
    unsigned long long* barrier;        /* train: TWO zero-initialised 64-bit counters (arrivals, exits); the kernel leaves them at zero */
} ibinn_aio_stage_args;

size_t ibinn_sizeof_aio_block_desc(void);
size_t ibinn_sizeof_aio_stage_args(void);
/* 1 if this geometry / mode runs on the fused kernel, 0 if the caller must use the per-op entry points */
int ibinn_aio_stage_supported(const ibinn_aio_stage_args* a);
int64_t ibinn_aio_stage_workspace_floats(const ibinn_aio_stage_args* a);
int ibinn_aio_stage_forward(const ibinn_aio_stage_args* a, ibinn_stream_t stream);

/* inverse direction (all_in_one_block.py:69,93-103): u = w_inv . ((y - act_offset) * exp(-act_norm)),
 * then, with a = subnet(u[:, :C-C/2]):  u[:, C-C/2:] <- (u2 - t) * exp(-sig) in place; jac = -(...) */
int ibinn_aio_unpermute(const float* y, const float* w_inv, const float* act_norm, const float* act_offset, float* u,
                        int B, int C, int H, int W, ibinn_stream_t stream);
int ibinn_aio_affine_inverse(float* u, const float* a, const float* act_norm, float* jac, int B, int C, int H, int W,
                             float clamp, float alpha, ibinn_stream_t stream);

/* ---------------------------------------------------------------- DownsampleCouplingBlock halves
 * Replaces reference downsampling_coupling_block.py:49-66 (`down`/`up` grouped stride-2 convs with one-hot
 * kernels + `affine`):  out[b,4c+2dy+dx,h,w] = x[b,c,2h+dy,2w+dx] * exp(clamp*tanh(alpha*a[b,4c+2dy+dx,h,w]))
 *                                              + a[b,4cp+4c+2dy+dx,h,w]
 * x: (B,cp,H,W) slice with batch stride x_bs; a: (B,8cp,H/2,W/2) contiguous; out: (B,4cp,H/2,W/2) slice, out_bs.
 * jac[b] (+)= sum of the log-scales. */
int ibinn_s2d_coupling_forward(const float* x, int64_t x_bs, const float* a, float* out, int64_t out_bs, float* jac,
                               int jac_accumulate, int B, int cp, int H, int W, float clamp, float alpha,
                               ibinn_stream_t stream);
int ibinn_s2d_coupling_inverse(const float* y, int64_t y_bs, const float* a, float* x, int64_t x_bs, float* jac,
                               int jac_accumulate, int B, int cp, int H, int W, float clamp, float alpha,
                               ibinn_stream_t stream);
int ibinn_s2d_coupling_backward(const float* g_out, int64_t g_bs, const float* g_jac, const float* x, int64_t x_bs,
                                const float* a, float* g_x, int64_t gx_bs, float* g_a, int B, int cp, int H, int W,
                                float clamp, float alpha, ibinn_stream_t stream);

/* ---------------------------------------------------------------- GLOWCouplingBlock halves (FrEIA v0.2; reference
 * call site inn_architecture.py:161):  y = exp(le(r[:, :L])) * x + r[:, L:],  le(s) = clamp*0.636*atan(s/clamp)
 * x,y: (B,L) with row strides; r: (B,2L) contiguous; jac[b] (+)= sum le. */
int ibinn_glow_affine_forward(const float* x, int64_t x_rs, const float* r, float* y, int64_t y_rs, float* jac,
                              int jac_accumulate, int B, int L, float clamp, ibinn_stream_t stream);
int ibinn_glow_affine_inverse(const float* y, int64_t y_rs, const float* r, float* x, int64_t x_rs, float* jac,
                              int jac_accumulate, int B, int L, float clamp, ibinn_stream_t stream);
int ibinn_glow_affine_backward(const float* g_y, int64_t gy_rs, const float* g_jac, const float* x, int64_t x_rs,
                               const float* r, float* g_x, int64_t gx_rs, float* g_r, int B, int L, float clamp,
                               ibinn_stream_t stream);

/* ---------------------------------------------------------------- data-movement nodes
 * DCTPooling2d (reference dct_transform.py:81-101): out[b,(kw*N+kh)*C+c] = scale * sum M[kh][h] M[kw][w] x[b,c,h,w].
 * inverse / backward: x[b,c,h,w] = scale * sum A(h,kh) A(w,kw) t[b,(kw*N+kh)*C+c], A = M (rev, M = inv_weight)
 * or M^T (autograd backward, M = weight). */
int ibinn_dct2d_forward(const float* x, const float* M, float* out, int B, int C, int N, float scale, ibinn_stream_t stream);
int ibinn_dct2d_inverse(const float* t, const float* M, float* x, int B, int C, int N, float scale, int transpose_M,
                        ibinn_stream_t stream);
/* HaarDownsampling (FrEIA v0.2; reference call site inn_architecture.py:127): (B,C,H,W) <-> (B,4C,H/2,W/2) */
int ibinn_haar_forward(const float* x, float* y, int B, int C, int H, int W, float fac, ibinn_stream_t stream);
int ibinn_haar_inverse(const float* y, float* x, int B, int C, int H, int W, float fac, ibinn_stream_t stream);
/* PermuteRandom (FrEIA v0.2; reference call site inn_architecture.py:160): out[b][i] = x[b][index[i]] */
int ibinn_permute_features(const float* x, const int32_t* index, float* out, int B, int D, ibinn_stream_t stream);

/* ---------------------------------------------------------------- fully connected subnets (fc_constr,
 * reference inn_architecture.py:112-120) and their autograd backward.
 * Small-batch GEMMs are split along the contraction to fill the SMs.  With `workspace` (ibinn_linear_workspace_floats(M, N, Kc)
 * floats; (M, N) = shape of the result, Kc = contraction length: forward (B, N, K), dgrad (B, K, N), wgrad (N, K, B)) and `tickets`
 * (ibinn_linear_tickets(M, N, Kc) zero-initialised uint32 slots, returned to zero) the partial results are summed in split order
 * by the last CTA of each output tile: bit-reproducible.  workspace = NULL: fp32 atomics on the result (not reproducible). */
int64_t ibinn_linear_workspace_floats(int M, int N, int Kc);
int ibinn_linear_tickets(int M, int N, int Kc);
int ibinn_linear_forward(const float* x, int64_t x_rs, const float* x_mask, int x_pre, const float* w, const float* bias,
                         float* y, int B, int K, int N, float* workspace, uint32_t* tickets, ibinn_stream_t stream);
int ibinn_linear_dgrad(const float* g, const float* h, const float* g_mask, int g_pre, const float* w, float* dx,
                       int64_t dx_rs, int accumulate, int B, int K, int N, float* workspace, uint32_t* tickets, ibinn_stream_t stream);
int ibinn_linear_wgrad(const float* g, const float* h, const float* g_mask, int g_pre, const float* x, int64_t x_rs,
                       const float* x_mask, int x_pre, float* dw, float* db, int B, int K, int N, float* workspace, uint32_t* tickets,
                       ibinn_stream_t stream);

/* ---------------------------------------------------------------- GMM / Information-Bottleneck head
 * Replaces reference model.py:113-126 (`cluster_distances`) and model.py:144-163 (loss dict).
 * z (B,D), jac (B), mu (C,D), phi (C), y (B,C) soft one-hot or NULL.  Per-sample outputs L_x, logits (B,C),
 * L_cNLL, L_y, correct (0/1), lse; means[5] = mean(L_x), mean(logits), mean(L_cNLL), mean(L_y), acc.
 * `partials` is scratch of ibinn_gmm_head_partials_floats(B) floats (contents unspecified afterwards: per-sample records, or per-CTA
 * fp64 records at the scoring batches when it is 8-byte aligned); `counter` one zero-initialised uint32, left at zero. */
int64_t ibinn_gmm_head_partials_floats(int B);
int ibinn_gmm_head_forward(const float* z, const float* jac, const float* mu, const float* phi, const float* y, float* L_x,
                           float* logits, float* L_cNLL, float* L_y, float* correct, float* lse, float* means,
                           float* partials, uint32_t* counter, int B, int C, int D, ibinn_stream_t stream);
/* The same outputs at the scoring batches (B >= 1024, C <= 112, D % 64 == 0) with the distance GEMM z . mu^T on tcgen05 (3xTF32,
 * fp32-accurate; z streamed once from HBM, split-K over the SMs).  `workspace`: ibinn_gmm_head_gemm_workspace_floats() floats,
 * 16-byte aligned; its head holds the prepared class means (operand image, centre, squared norms): pass mu_prepared = 1 when the
 * same workspace was used with the same `mu` before (scoring: the class means do not change between batches).
 * Per-sample results agree with ibinn_gmm_head_forward to fp32 rounding (not bit for bit: different summation). */
int ibinn_gmm_head_gemm_supported(int B, int C, int D);
int64_t ibinn_gmm_head_gemm_workspace_floats(int B, int C, int D);
int ibinn_gmm_head_forward_gemm(const float* z, const float* jac, const float* mu, const float* phi, const float* y, float* L_x,
                                float* logits, float* L_cNLL, float* L_y, float* correct, float* lse, float* means,
                                float* partials, uint32_t* counter, float* workspace, int mu_prepared, int B, int C, int D,
                                ibinn_stream_t stream);
/* upstream grads g_* may be NULL; s_* = element stride (0 broadcasts one scalar); they are multiplied by
 * scale_vec / scale_logits (1/B and 1/(B*C) when the means were the outputs).  G_ws, Glw_ws: (B,C) workspaces. */
int ibinn_gmm_head_backward(const float* z, const float* mu, const float* phi, const float* y, const float* logits,
                            const float* lse, const float* g_Lx, int s_Lx, const float* g_LcNLL, int s_Lc,
                            const float* g_Ly, int s_Ly, const float* g_logits, int s_logits, float scale_vec,
                            float scale_logits, float* G_ws, float* Glw_ws, float* g_z, float* g_jac, float* g_mu,
                            int accumulate_mu, float* g_phi, int accumulate_phi, int B, int C, int D,
                            ibinn_stream_t stream);

/* ---------------------------------------------------------------- training-data pipeline
 * Replaces the per-image CPU work of reference data.py:47-90 (`Augmentor`) and data.py:152-160 (the torchvision transform chain
 * of the CIFAR training set) for a batch whose raw images live in HBM: images (N,H,W,C) uint8; index (B) int32 or NULL;
 * params (B,8) = flip, brightness, contrast, saturation, rotation angle [rad], crop_x, crop_y in [0,4], unused -- NULL selects the
 * deterministic test transform (ToTensor + normalisation); unif (B,C,H,W) U[0,1) or NULL; gauss (B,C+pad,H,W) N(0,1) or NULL;
 * out (B,C+pad,H,W) fp32 NCHW.  beta / gamma are HOST arrays of C floats (data.py:107-110). */
int ibinn_augment_batch(const unsigned char* images, const int32_t* index, const float* params, const float* unif, const float* gauss,
                        float* out, int B, int H, int W, int C, int pad_channels, const float* beta, const float* gamma, float sigma,
                        float pad_sigma, int tanh_augmentation, ibinn_stream_t stream);

/* ---------------------------------------------------------------- clip_grad_norm_ + SGD (reference train.py:109-110)
 * out_norm_coef[0] = sqrt(sum grad^2 + sum extra_sumsq), [1] = min(1, max_norm/(norm+1e-6)). */
int64_t ibinn_grad_norm_partials_doubles(void);
int ibinn_grad_norm_clip(const float* grad, int64_t n, const float* extra_sumsq, int n_extra, float max_norm,
                         double* partials, uint32_t* counter, float* out_norm_coef, ibinn_stream_t stream);
/* hyper = {lr, momentum, weight_decay} in device memory; clip_coef device scalar or NULL.
 * d = grad*coef*grad_scale + wd*p; buf = momentum*buf + d; p -= lr*buf   (torch.optim.SGD, dampening 0) */
int ibinn_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, const float* hyper,
                   const float* clip_coef, float grad_scale, ibinn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
