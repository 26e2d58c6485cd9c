// Training-data pipeline on the GPU: what the reference does per image on 6 CPU workers with PIL
// (reference data.py:47-90 `Augmentor`, data.py:152-160 the torchvision transform chain of the CIFAR training set) as ONE kernel
// over a batch whose raw uint8 images already live in HBM (the whole CIFAR training set is 150 MB):
//   RandomHorizontalFlip -> ColorJitter(brightness, contrast, saturation) -> Pad(8, edge) -> RandomRotation (nearest) ->
//   CenterCrop(36) -> RandomCrop(32) -> ToTensor -> + U[0,1)/256 -> + sigma N(0,1) -> gamma (x - beta) -> [atanh] -> [noise channels]
// One CTA per output image; every output pixel is pulled through the inverse of the geometric chain (one gather, no intermediate
// images); the random decisions arrive as a (B, 8) parameter row per sample, the per-element noise as optional tensors, so the
// kernel itself is deterministic (torch's Philox streams are not part of parity).
// Deviation from PIL, stated: colour arithmetic stays in fp32 (PIL rounds to uint8 after each of the three enhance steps), and the
// three jitter steps run in the fixed order brightness, contrast, saturation (torchvision shuffles them per image).
#include "common.cuh"

namespace {

struct AugP {
    const unsigned char* images;   // (N, H, W, C) uint8
    const int* index;              // (B) image index per output sample, or NULL (identity)
    const float* params;           // (B, 8): flip, brightness, contrast, saturation, angle [rad], crop_x, crop_y, unused; NULL = test transform
    const float* unif;             // (B, C, H, W) U[0,1) or NULL
    const float* gauss;            // (B, C + pad, H, W) N(0,1) or NULL
    float* out;                    // (B, C + pad, H, W)
    int B, H, W, C, pad;
    float sigma, pad_sigma, do_tanh;
    float beta[4], gamma[4];
};

__device__ __forceinline__ float clip255(float v) { return fminf(fmaxf(v, 0.f), 255.f); }
__device__ __forceinline__ float gray(float r, float g, float b) { return 0.299f * r + 0.587f * g + 0.114f * b; }

__global__ void __launch_bounds__(IBINN_THREADS) augment_kernel(const AugP p) {
    __shared__ float s_scr[40];
    const int b = blockIdx.x, tid = threadIdx.x, HW = p.H * p.W, C = p.C;
    const int src_i = p.index ? p.index[b] : b;
    const unsigned char* img = p.images + (size_t)src_i * HW * C;
    float flip = 0.f, br = 1.f, ct = 1.f, sat = 1.f, ang = 0.f, cx = 2.f, cy = 2.f;
    if (p.params) {
        const float* q = p.params + (size_t)b * 8;
        flip = q[0]; br = q[1]; ct = q[2]; sat = q[3]; ang = q[4]; cx = q[5]; cy = q[6];
    }
    // mean grey level of the brightness-adjusted image: the pivot of the contrast step (PIL ImageEnhance.Contrast)
    float mean_gray = 0.f;
    if (p.params && C == 3) {
        float s = 0.f;
        for (int i = tid; i < HW; i += IBINN_THREADS)
            s += gray(clip255(br * img[i * 3]), clip255(br * img[i * 3 + 1]), clip255(br * img[i * 3 + 2]));
        mean_gray = floorf(block_sum(s, s_scr) / (float)HW + 0.5f);
    }
    const float cs = cosf(ang), sn = sinf(ang);
    const float pc = 0.5f * (float)(p.W + 16);            // centre of the padded image (48 x 48 for CIFAR)
    const int off = (p.W + 16 - (p.W + 4)) / 2;           // CenterCrop(36) inside the padded image
    float* ob = p.out + (size_t)b * (C + p.pad) * HW;
    for (int i = tid; i < HW; i += IBINN_THREADS) {
        const int y = i / p.W, x = i - y * p.W;
        float v[3] = {0.f, 0.f, 0.f};
        bool inside = true;
        int sy = y, sx = x;
        if (p.params) {
            // output pixel -> RandomCrop offset -> CenterCrop offset -> padded image -> inverse rotation (nearest) -> un-pad (edge) -> un-flip
            const float X = (float)(x + (int)cx + off) + 0.5f - pc, Y = (float)(y + (int)cy + off) + 0.5f - pc;
            const float xs = cs * X - sn * Y + pc, ys = sn * X + cs * Y + pc;
            const int ix = (int)floorf(xs), iy = (int)floorf(ys);
            inside = ix >= 0 && ix < p.W + 16 && iy >= 0 && iy < p.H + 16;      // outside the rotated canvas: filled with black
            sx = min(max(ix - 8, 0), p.W - 1);
            sy = min(max(iy - 8, 0), p.H - 1);
            if (flip > 0.5f) sx = p.W - 1 - sx;
        }
        if (inside)
            for (int c = 0; c < C; ++c) v[c] = (float)img[(sy * p.W + sx) * C + c];
        if (p.params && C == 3 && inside) {
            float r = clip255(br * v[0]), g = clip255(br * v[1]), bl = clip255(br * v[2]);
            r = clip255(fmaf(ct, r - mean_gray, mean_gray)); g = clip255(fmaf(ct, g - mean_gray, mean_gray)); bl = clip255(fmaf(ct, bl - mean_gray, mean_gray));
            const float gy = gray(r, g, bl);
            v[0] = clip255(fmaf(sat, r - gy, gy)); v[1] = clip255(fmaf(sat, g - gy, gy)); v[2] = clip255(fmaf(sat, bl - gy, gy));
        }
        for (int c = 0; c < C; ++c) {
            float t = v[c] * (1.0f / 255.0f);                                   // ToTensor
            if (p.unif) t += p.unif[((size_t)b * C + c) * HW + i] * (1.0f / 256.0f);   // data.py:62-63
            if (p.gauss && p.sigma > 0.f) t += p.sigma * p.gauss[((size_t)b * (C + p.pad) + c) * HW + i];      // data.py:64-65
            t = p.gamma[c] * (t - p.beta[c]);                                   // data.py:67
            if (p.do_tanh > 0.f) {                                              // data.py:69-71
                t = fminf(fmaxf(t, -(1.f - 1e-7f)), 1.f - 1e-7f);
                t = 0.5f * logf((1.f + t) / (1.f - t));
            }
            v[c] = t;
            ob[(size_t)c * HW + i] = t;
        }
        for (int c = 0; c < p.pad; ++c) {                                       // data.py:73-77: noise channels
            float t = v[c % C] * sqrtf(1.f - p.pad_sigma * p.pad_sigma);
            if (p.gauss) t += p.pad_sigma * p.gauss[((size_t)b * (C + p.pad) + C + c) * HW + i];
            ob[(size_t)(C + c) * HW + i] = t;
        }
    }
}

}  // namespace

extern "C" int ibinn_augment_batch(const unsigned char* images, const int32_t* index, const float* params, const float* unif, const float* gauss,
                                   float* out, int B, int H, int W, int C, int pad_channels, const float* beta, const float* gamma, float sigma,
                                   float pad_sigma, int tanh_augmentation, ibinn_stream_t stream) {
    int e = ibinn_check_device();
    if (e) return e;
    if (!images || !out || !beta || !gamma) return IBINN_ENULL;
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || C > 3 || pad_channels < 0) return IBINN_EINVAL;
    AugP p{};
    p.images = images; p.index = index; p.params = params; p.unif = unif; p.gauss = gauss; p.out = out;
    p.B = B; p.H = H; p.W = W; p.C = C; p.pad = pad_channels; p.sigma = sigma; p.pad_sigma = pad_sigma; p.do_tanh = tanh_augmentation ? 1.f : 0.f;
    for (int c = 0; c < C; ++c) { p.beta[c] = beta[c]; p.gamma[c] = gamma[c]; }      // HOST arrays of C floats
    IBINN_LAUNCH(augment_kernel, dim3(B), dim3(IBINN_THREADS), 0, (cudaStream_t)stream, p);
    return ibinn_launch_status();
}
