// Training-time augmentations on the device, batched over crops.
//
// Replace the per-sample NumPy / imgaug / scipy transforms of src/training/mytransforms.py:13-406 (Flip, Contrast,
// Scaling, Rotate, Blur, Noise, ToTensor) that run in the reference's DataLoader workers: at > 500 crops/s per GPU the
// CPU pipeline cannot keep a B200 fed.  The random DECISIONS and parameters are drawn on the host in the reference's own
// call order (microbeseg_b200/augment.py); the kernels apply them.  uint16 images [n][H][W], float32 / uint8 labels.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/mbseg.h"
#include "common.cuh"

namespace {

// ---- geometric transforms: Flip (8 dihedral maps, exact), Scaling / Rotate (imgaug Affine -> cv2.warpAffine, constant
// border 0): inverse map (sx, sy) = M * (x, y, 1) per sample; mode 0 copy, 1 nearest (order 0), 2 bilinear (order 1)
template <typename T>
__device__ __forceinline__ T finish(double v);
template <>
__device__ __forceinline__ float finish<float>(double v) { return static_cast<float>(v); }
template <>
__device__ __forceinline__ uint16_t finish<uint16_t>(double v) {            // cv2 saturate_cast: round half to even, clamp
    v = rint(v);
    return static_cast<uint16_t>(v < 0.0 ? 0.0 : (v > 65535.0 ? 65535.0 : v));
}
template <>
__device__ __forceinline__ uint8_t finish<uint8_t>(double v) {
    v = rint(v);
    return static_cast<uint8_t>(v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v));
}

template <typename T>
__global__ void __launch_bounds__(256)
aug_warp_kernel(const T *__restrict__ src, T *__restrict__ dst, int H, int W, const double *__restrict__ mats, const int *__restrict__ modes) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t base = static_cast<size_t>(b) * H * W;
    const int mode = modes[b];
    if (mode == 0) {
        dst[base + static_cast<size_t>(y) * W + x] = src[base + static_cast<size_t>(y) * W + x];
        return;
    }
    const double *m = mats + 6 * b;
    const double sx = m[0] * x + m[1] * y + m[2], sy = m[3] * x + m[4] * y + m[5];
    auto at = [&](int yy, int xx) -> double {
        return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? static_cast<double>(src[base + static_cast<size_t>(yy) * W + xx]) : 0.0;
    };
    if (mode == 1) {
        const int ix = static_cast<int>(floor(sx + 0.5)), iy = static_cast<int>(floor(sy + 0.5));
        dst[base + static_cast<size_t>(y) * W + x] = finish<T>(at(iy, ix));
        return;
    }
    const double fx = floor(sx), fy = floor(sy);
    const int ix = static_cast<int>(fx), iy = static_cast<int>(fy);
    const double ax = sx - fx, ay = sy - fy;
    const double v = (1.0 - ay) * ((1.0 - ax) * at(iy, ix) + ax * at(iy, ix + 1)) + ay * ((1.0 - ax) * at(iy + 1, ix) + ax * at(iy + 1, ix + 1));
    dst[base + static_cast<size_t>(y) * W + x] = finish<T>(v);
}

// ---- per-sample 65536-bin histogram of a uint16 image (percentiles, mean, min, max are exact functions of it)
__global__ void __launch_bounds__(256) aug_hist_kernel(const uint16_t *__restrict__ img, int HW, unsigned int *__restrict__ hist) {
    const int b = blockIdx.y;
    const uint16_t *p = img + static_cast<size_t>(b) * HW;
    unsigned int *h = hist + static_cast<size_t>(b) * 65536;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) atomicAdd(&h[p[i]], 1u);
}

// derived[b] = {p0, p1, mean, min, max, 0, 0, 0}: np.percentile (linear interpolation between order statistics, numpy's
// _lerp with its t >= 0.5 form), mean / min / max of the integer image.  One CTA per sample.
__device__ double order_stat(const unsigned int *s_cum, const unsigned int *h, long long k) {
    // smallest value v with cumulative count > k; s_cum holds the cumulative counts of 256-bin groups
    int g = 0;
    while (g < 255 && static_cast<long long>(s_cum[g]) <= k) ++g;
    long long c = g ? s_cum[g - 1] : 0;
    int v = g * 256;
    for (; v < g * 256 + 255; ++v) {
        c += h[v];
        if (c > k) break;
    }
    return static_cast<double>(v);
}
__global__ void __launch_bounds__(256) aug_derive_kernel(const unsigned int *__restrict__ hist, int HW, const double *__restrict__ params,
                                                          double *__restrict__ derived) {
    __shared__ unsigned int s_cum[256];
    __shared__ double s_sum[256];
    __shared__ int s_min[256], s_max[256];
    const int b = blockIdx.x, t = threadIdx.x;
    const unsigned int *h = hist + static_cast<size_t>(b) * 65536;
    unsigned int cnt = 0;
    double sum = 0.0;
    int mn = 65536, mx = -1;
    for (int v = t * 256; v < t * 256 + 256; ++v) {
        const unsigned int c = h[v];
        cnt += c;
        sum += static_cast<double>(c) * v;
        if (c) { mn = mn < v ? mn : v; mx = v; }
    }
    s_cum[t] = cnt; s_sum[t] = sum; s_min[t] = mn; s_max[t] = mx;
    __syncthreads();
    if (t == 0) {
        double tot = 0.0;
        int lo = 65536, hi = -1;
        unsigned int run = 0;
        for (int g = 0; g < 256; ++g) {
            run += s_cum[g];
            s_cum[g] = run;
            tot += s_sum[g];
            lo = s_min[g] < lo ? s_min[g] : lo;
            hi = s_max[g] > hi ? s_max[g] : hi;
        }
        double *d = derived + 8 * b;
        for (int k = 0; k < 2; ++k) {
            const double q = params[4 * b + k];
            const double pos = q / 100.0 * static_cast<double>(HW - 1);
            const double fl = floor(pos);
            const double tt = pos - fl;
            const long long i0 = static_cast<long long>(fl), i1 = i0 + 1 < HW ? i0 + 1 : i0;
            const double a = order_stat(s_cum, h, i0), bb = order_stat(s_cum, h, i1);
            const double diff = bb - a;
            d[k] = tt >= 0.5 ? bb - diff * (1.0 - tt) : a + diff * tt;
        }
        d[2] = tot / static_cast<double>(HW);
        d[3] = static_cast<double>(lo);
        d[4] = static_cast<double>(hi);
    }
}

// Contrast (mytransforms.py:64-126) per sample: mode 1 = percentile stretch (np.percentile + skimage rescale_intensity to
// the dtype range, float64, truncating cast), mode 2 = contrast + gamma adjustment in float32 as the reference writes it
__global__ void __launch_bounds__(256) aug_contrast_kernel(uint16_t *__restrict__ img, int HW, const int *__restrict__ modes,
                                                            const double *__restrict__ params, const double *__restrict__ derived) {
    const int b = blockIdx.y;
    const int mode = modes[b];
    if (mode == 0) return;
    uint16_t *p = img + static_cast<size_t>(b) * HW;
    const double *d = derived + 8 * b;
    if (mode == 1) {
        const double p0 = d[0], p1 = d[1];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
            double v = static_cast<double>(p[i]);
            v = v < p0 ? p0 : (v > p1 ? p1 : v);
            if (p0 != p1) v = __ddiv_rn(__dsub_rn(v, p0), __dsub_rn(p1, p0));
            p[i] = static_cast<uint16_t>(__dadd_rn(__dmul_rn(v, 65535.0), 0.0));
        }
        return;
    }
    // mode 2: img = f32(img) / 65535; (img - mean) * factor + mean; ((img - min) / (rnge + 1e-7)) ** gamma * rnge + min;
    // clip(0, 1) * 65535 -> uint16 (truncation).  mean / min / max of the transformed image follow from the integer ones.
    const float factor = static_cast<float>(params[4 * b + 2]), gamma = static_cast<float>(params[4 * b + 3]);
    const float mean0 = static_cast<float>(d[2] / 65535.0);
    const float lo = __fadd_rn(__fmul_rn(__fsub_rn(__fdiv_rn(static_cast<float>(d[3]), 65535.0f), mean0), factor), mean0);
    const float hi = __fadd_rn(__fmul_rn(__fsub_rn(__fdiv_rn(static_cast<float>(d[4]), 65535.0f), mean0), factor), mean0);
    const float mn = factor >= 0.0f ? lo : hi, mx = factor >= 0.0f ? hi : lo;
    const float rnge = __fsub_rn(mx, mn);
    const float den = __fadd_rn(rnge, 1e-7f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        float v = __fdiv_rn(static_cast<float>(p[i]), 65535.0f);
        v = __fadd_rn(__fmul_rn(__fsub_rn(v, mean0), factor), mean0);
        v = __fadd_rn(__fmul_rn(powf(__fdiv_rn(__fsub_rn(v, mn), den), gamma), rnge), mn);
        v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
        p[i] = static_cast<uint16_t>(__fmul_rn(v, 65535.0f));
    }
}

// Blur (mytransforms.py:38-61): scipy.ndimage.gaussian_filter(img (H,W,1) uint16, sigma): correlate1d along H, then W,
// then the singleton axis; every pass accumulates in float64 in NI_Correlate1D's symmetric order
// (c*w[r] + sum_k (x[-k] + x[+k]) * w[r-k], outermost pair first), 'reflect' borders, and TRUNCATES into the uint16 output.
// weights[b]: [17] float64 from the host (numpy's own exp / normalisation), radius[b] <= 8; radius 0 = no blur.
constexpr int AUG_MAXR = 8;
__global__ void __launch_bounds__(256) aug_blur_pass_kernel(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int H, int W, int axis,
                                                             const double *__restrict__ weights, const int *__restrict__ radius, int last) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t base = static_cast<size_t>(b) * H * W;
    const int r = radius[b];
    const size_t o = base + static_cast<size_t>(y) * W + x;
    if (r == 0) { dst[o] = src[o]; return; }
    const double *w = weights + 17 * b;          // w[0..2r], centre w[r]
    const int n = axis == 0 ? H : W, pos = axis == 0 ? y : x;
    auto at = [&](int i) -> double {
        // scipy 'reflect': (d c b a | a b c d | d c b a)
        while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
        return static_cast<double>(axis == 0 ? src[base + static_cast<size_t>(i) * W + x] : src[base + static_cast<size_t>(y) * W + i]);
    };
    double tmp = __dmul_rn(at(pos), w[r]);
    for (int k = r; k >= 1; --k) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(at(pos - k), at(pos + k)), w[r - k]));
    uint16_t v = static_cast<uint16_t>(tmp);
    if (last) {
        // third pass over the singleton channel axis: every tap sees the same value (reflect)
        const double c = static_cast<double>(v);
        double t3 = __dmul_rn(c, w[r]);
        for (int k = r; k >= 1; --k) t3 = __dadd_rn(t3, __dmul_rn(__dadd_rn(c, c), w[r - k]));
        v = static_cast<uint16_t>(t3);
    }
    dst[o] = v;
}

__global__ void __launch_bounds__(256) aug_max_kernel(const uint16_t *__restrict__ img, int HW, unsigned int *__restrict__ mx) {
    const int b = blockIdx.y;
    const uint16_t *p = img + static_cast<size_t>(b) * HW;
    unsigned int m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) m = p[i] > m ? p[i] : m;
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned int t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(&mx[b], m);
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {       // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Noise (mytransforms.py:236-268: imgaug AdditiveGaussianNoise(scale = k/100 * max(img)), samples rounded, result clipped
// to the dtype) followed by ToTensor's normalisation 2 * (clip(img) - lo) / (hi - lo) - 1 in float32 (utils.py:50-74).
// Counter-based generator (seed, sample, pixel) -> Box-Muller: reproducible, independent of launch geometry.
__global__ void __launch_bounds__(256) aug_noise_normalize_kernel(const uint16_t *__restrict__ img, int HW, const float *__restrict__ noise_frac,
                                                                   const unsigned int *__restrict__ mx, unsigned long long seed, float lo, float hi,
                                                                   uint16_t *__restrict__ img_out, float *__restrict__ out) {
    const int b = blockIdx.y;
    const float frac = noise_frac[b];
    const double sigma = static_cast<double>(frac) * static_cast<double>(mx[b]);
    const uint16_t *p = img + static_cast<size_t>(b) * HW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        double v = static_cast<double>(p[i]);
        if (frac > 0.0f) {
            const unsigned long long r = mix64(seed + 0x9E3779B97F4A7C15ull * (static_cast<unsigned long long>(b) * HW + i + 1));
            const double u1 = (static_cast<double>(r >> 40) + 0.5) * (1.0 / 16777216.0);            // (0, 1)
            const double u2 = (static_cast<double>((r >> 16) & 0xFFFFFFull) + 0.5) * (1.0 / 16777216.0);
            const double g = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
            v += rint(g * sigma);
            v = v < 0.0 ? 0.0 : (v > 65535.0 ? 65535.0 : v);
        }
        if (img_out) img_out[static_cast<size_t>(b) * HW + i] = static_cast<uint16_t>(v);
        if (out) {
            float f = static_cast<float>(v);
            f = f < lo ? lo : (f > hi ? hi : f);
            out[static_cast<size_t>(b) * HW + i] = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(f, lo)), __fsub_rn(hi, lo)), 1.0f);
        }
    }
}

}  // namespace

extern "C" int mbs_aug_warp(const void *src, void *dst, int dtype, int n, int H, int W, const double *mats_dev, const int *modes_dev,
                            void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(src && dst && src != dst && n > 0 && H > 0 && W > 0 && mats_dev && modes_dev, "aug_warp: bad arguments");
    const dim3 grid(mbs::cdiv(W, 32), mbs::cdiv(H, 8), n);
    switch (dtype) {
        case MBS_IN_U8: aug_warp_kernel<uint8_t><<<grid, 256, 0, stream>>>(static_cast<const uint8_t *>(src), static_cast<uint8_t *>(dst), H, W, mats_dev, modes_dev); break;
        case MBS_IN_U16: aug_warp_kernel<uint16_t><<<grid, 256, 0, stream>>>(static_cast<const uint16_t *>(src), static_cast<uint16_t *>(dst), H, W, mats_dev, modes_dev); break;
        case MBS_IN_F32: aug_warp_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(src), static_cast<float *>(dst), H, W, mats_dev, modes_dev); break;
        default: MBS_REQUIRE(false, "aug_warp: unknown dtype %d", dtype);
    }
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t mbs_aug_workspace_bytes(int n) { return static_cast<size_t>(n) * (65536 * 4 + 8 * 8 + 4) + 256; }

extern "C" int mbs_aug_contrast(uint16_t *img, int n, int H, int W, const int *modes_dev, const double *params_dev, void *workspace,
                                size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(img && n > 0 && H > 0 && W > 0 && modes_dev && params_dev && workspace, "aug_contrast: bad arguments");
    MBS_REQUIRE(workspace_bytes >= mbs_aug_workspace_bytes(n), "aug_contrast: workspace too small");
    const int HW = H * W;
    unsigned int *hist = static_cast<unsigned int *>(workspace);
    double *derived = reinterpret_cast<double *>(hist + static_cast<size_t>(n) * 65536);
    MBS_CHECK_CUDA(cudaMemsetAsync(hist, 0, static_cast<size_t>(n) * 65536 * 4, stream));
    const dim3 grid(mbs::cdiv(HW, 256 * 8) < 64 ? mbs::cdiv(HW, 256 * 8) : 64, n);
    aug_hist_kernel<<<grid, 256, 0, stream>>>(img, HW, hist);
    MBS_CHECK_LAUNCH();
    aug_derive_kernel<<<n, 256, 0, stream>>>(hist, HW, params_dev, derived);
    MBS_CHECK_LAUNCH();
    aug_contrast_kernel<<<grid, 256, 0, stream>>>(img, HW, modes_dev, params_dev, derived);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_aug_blur(uint16_t *img, uint16_t *tmp, int n, int H, int W, const double *weights_dev, const int *radius_dev,
                            void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(img && tmp && img != tmp && n > 0 && H > 0 && W > 0 && weights_dev && radius_dev, "aug_blur: bad arguments");
    const dim3 grid(mbs::cdiv(W, 32), mbs::cdiv(H, 8), n);
    aug_blur_pass_kernel<<<grid, 256, 0, stream>>>(img, tmp, H, W, 0, weights_dev, radius_dev, 0);
    MBS_CHECK_LAUNCH();
    aug_blur_pass_kernel<<<grid, 256, 0, stream>>>(tmp, img, H, W, 1, weights_dev, radius_dev, 1);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_aug_noise_normalize(const uint16_t *img, int n, int H, int W, const float *noise_frac_dev, unsigned long long seed,
                                       float min_value, float max_value, uint16_t *img_out, float *out, void *workspace,
                                       size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(img && n > 0 && H > 0 && W > 0 && noise_frac_dev && (img_out || out) && workspace, "aug_noise_normalize: bad arguments");
    MBS_REQUIRE(workspace_bytes >= mbs_aug_workspace_bytes(n), "aug_noise_normalize: workspace too small");
    const int HW = H * W;
    unsigned int *mx = static_cast<unsigned int *>(workspace);
    MBS_CHECK_CUDA(cudaMemsetAsync(mx, 0, static_cast<size_t>(n) * 4, stream));
    const dim3 grid(mbs::cdiv(HW, 256 * 8) < 64 ? mbs::cdiv(HW, 256 * 8) : 64, n);
    aug_max_kernel<<<grid, 256, 0, stream>>>(img, HW, mx);
    MBS_CHECK_LAUNCH();
    aug_noise_normalize_kernel<<<grid, 256, 0, stream>>>(img, HW, noise_frac_dev, mx, seed, min_value, max_value, img_out, out);
    MBS_CHECK_LAUNCH();
    return 0;
}
