// Elementwise / reduction kernels of the training step (HBM bound).
//
// Replace what torch autograd + cuDNN/ATen do around the convolutions in the reference's inner training
// loop (src/training/train.py:460-493): BatchNorm2d in training mode (batch statistics, unets.py:128,153,206,246),
// the activation derivative, the 1x1 heads, SmoothL1Loss (losses.py:30-32, beta = 1, 'mean') and their backward
// passes.  Activations and activation gradients are NHWC bf16, statistics / parameter gradients fp32.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/mbseg.h"
#include "common.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per-channel reductions over an [M][C] bf16 matrix.  Block = 256 threads = 8 "pixel lanes" x 32 "channel lanes";
// a thread owns channels c0 + 32*j (j < C/32 handled in chunks of 8 channel groups to bound registers).
// Generic helper: F(pixel, channel, value) -> up to 2 accumulators.
template <int NACC, typename F>
__device__ __forceinline__ void channel_reduce(long long M, int C, float *out, F f) {
    __shared__ float s_red[8][32][2];
    const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
    for (int cbase = 0; cbase < C; cbase += 32) {
        const int c = cbase + cl;
        float acc[2] = {0.0f, 0.0f};
        if (c < C) {
            for (long long p = static_cast<long long>(blockIdx.x) * 8 + pl; p < M; p += static_cast<long long>(gridDim.x) * 8) {
                float v[2];
                f(p, c, v);
                acc[0] += v[0];
                if (NACC > 1) acc[1] += v[1];
            }
        }
        s_red[pl][cl][0] = acc[0];
        s_red[pl][cl][1] = acc[1];
        __syncthreads();
        if (pl == 0 && c < C) {
            float a0 = 0.0f, a1 = 0.0f;
            for (int k = 0; k < 8; ++k) {
                a0 += s_red[k][cl][0];
                a1 += s_red[k][cl][1];
            }
            atomicAdd(&out[c], a0);
            if (NACC > 1) atomicAdd(&out[C + c], a1);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) bn_stats_kernel(const __nv_bfloat16 *__restrict__ a, long long M, int C, float *sums) {
    channel_reduce<2>(M, C, sums, [&](long long p, int c, float *v) {
        const float x = __bfloat162float(a[p * C + c]);
        v[0] = x;
        v[1] = x * x;
    });
}

__global__ void bn_finalize_kernel(const float *__restrict__ sums, long long M, int C, float eps, float *mean, float *invstd,
                                   float *var_unbiased) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = static_cast<double>(sums[c]) / static_cast<double>(M);
    double var = static_cast<double>(sums[C + c]) / static_cast<double>(M) - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = static_cast<float>(m);
    invstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (var_unbiased) var_unbiased[c] = static_cast<float>(M > 1 ? var * static_cast<double>(M) / static_cast<double>(M - 1) : var);
}

__global__ void bn_apply_kernel(const __nv_bfloat16 *__restrict__ a, long long total, int C, const float *__restrict__ mean,
                                const float *__restrict__ invstd, const float *__restrict__ gamma,
                                const float *__restrict__ beta, __nv_bfloat16 *__restrict__ y) {
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
    if (i >= total) return;
    const int c = static_cast<int>(i % C);
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162 *>(a + i);
    const float x0 = (__bfloat162float(v.x) - mean[c]) * invstd[c] * gamma[c] + beta[c];
    const float x1 = (__bfloat162float(v.y) - mean[c + 1]) * invstd[c + 1] * gamma[c + 1] + beta[c + 1];
    *reinterpret_cast<__nv_bfloat162 *>(y + i) = __floats2bfloat162_rn(x0, x1);
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const __nv_bfloat16 *__restrict__ dy, const __nv_bfloat16 *__restrict__ a, long long M, int C,
                     const float *__restrict__ mean, const float *__restrict__ invstd, float *dgamma_dbeta) {
    channel_reduce<2>(M, C, dgamma_dbeta, [&](long long p, int c, float *v) {
        const float g = __bfloat162float(dy[p * C + c]);
        const float xh = (__bfloat162float(a[p * C + c]) - mean[c]) * invstd[c];
        v[0] = g * xh;
        v[1] = g;
    });
}

// dz = act'(a) * gamma * invstd * (dy - dbeta/M - xhat * dgamma/M);  dbias[c] += sum_p dz
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const __nv_bfloat16 *__restrict__ dy, const __nv_bfloat16 *__restrict__ a, long long M, int C,
                    const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ gamma,
                    const float *__restrict__ dgamma_dbeta, int act, __nv_bfloat16 *__restrict__ dz, float *dbias) {
    const float invM = 1.0f / static_cast<float>(M);
    channel_reduce<1>(M, C, dbias, [&](long long p, int c, float *v) {
        const float av = __bfloat162float(a[p * C + c]);
        const float xh = (av - mean[c]) * invstd[c];
        float g = gamma[c] * invstd[c] * (__bfloat162float(dy[p * C + c]) - dgamma_dbeta[C + c] * invM - xh * dgamma_dbeta[c] * invM);
        if (act == MBS_ACT_RELU && !(av > 0.0f)) g = 0.0f;
        const __nv_bfloat16 gb = __float2bfloat16_rn(g);
        dz[p * C + c] = gb;
        v[0] = __bfloat162float(gb);
    });
}

// 1x1 head forward: pred[p] = sum_c y[p][c] * w[c] + b
__global__ void head_fwd_kernel(const __nv_bfloat16 *__restrict__ y, long long M, int C, const float *__restrict__ w, float b,
                                float *__restrict__ pred) {
    const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= M) return;
    float acc = 0.0f;
    const __nv_bfloat162 *row = reinterpret_cast<const __nv_bfloat162 *>(y + p * C);
    for (int c = 0; c < C / 2; ++c) {
        const __nv_bfloat162 v = row[c];
        acc = fmaf(__bfloat162float(v.x), w[2 * c], acc);
        acc = fmaf(__bfloat162float(v.y), w[2 * c + 1], acc);
    }
    pred[p] = acc + b;
}

// SmoothL1Loss(beta = 1, reduction = 'mean'): loss += sum / M;  g = clamp(pred - target, -1, 1) / M
__global__ void __launch_bounds__(256)
smoothl1_kernel(const float *__restrict__ pred, const float *__restrict__ target, long long M, float *loss, float *__restrict__ g) {
    __shared__ float s_part[8];
    float acc = 0.0f;
    const float invM = 1.0f / static_cast<float>(M);
    for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < M;
         p += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float d = pred[p] - target[p];
        const float ad = fabsf(d);
        acc += ad < 1.0f ? 0.5f * d * d : ad - 0.5f;
        g[p] = fminf(fmaxf(d, -1.0f), 1.0f) * invM;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int k = 0; k < 8; ++k) t += s_part[k];
        atomicAdd(loss, t * invM);
    }
}

// head backward: dy[p][c] = g[p] * w[c];  dw[c] += sum_p g[p] * y[p][c];  dw[C] (= db) += sum_p g[p]
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float *__restrict__ g, const __nv_bfloat16 *__restrict__ y, long long M, int C, const float *__restrict__ w,
                __nv_bfloat16 *__restrict__ dy, float *dw_db) {
    channel_reduce<1>(M, C, dw_db, [&](long long p, int c, float *v) {
        const float gp = g[p];
        dy[p * C + c] = __float2bfloat16_rn(gp * w[c]);
        v[0] = gp * __bfloat162float(y[p * C + c]);
    });
    // bias gradient: one extra pass over g by block 0 .. gridDim (cheap: M floats)
    __shared__ float s_b[8];
    float acc = 0.0f;
    for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < M;
         p += static_cast<long long>(gridDim.x) * blockDim.x)
        acc += g[p];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_b[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int k = 0; k < 8; ++k) t += s_b[k];
        atomicAdd(&dw_db[C], t);
    }
}

// [N][H][W][C] -> channel-major [N][C][H][pitch] (bf16, pitch >= W, padding columns zero), 32x32 tiles via smem.
// One block row = 32 pixels of ONE image row (so that the padded pitch is simple).
__global__ void nhwc_to_chw_kernel(const __nv_bfloat16 *__restrict__ src, int H, int W, int C, int pitch, int shift,
                                   int step, __nv_bfloat16 *__restrict__ dst) {
    __shared__ __nv_bfloat16 tile[32][33];
    const int n = blockIdx.z / H, y = blockIdx.z % H;
    const int x0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const __nv_bfloat16 *s = src + (static_cast<size_t>(n) * H + y) * W * C;
    __nv_bfloat16 *d = dst + static_cast<size_t>(n) * C * H * pitch + static_cast<size_t>(y) * pitch;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int x = (x0 + r) * step + shift, c = c0 + threadIdx.x;   // destination column x0 + r reads source column x
        tile[r][threadIdx.x] = (x >= 0 && x < W && c < C) ? s[static_cast<size_t>(x) * C + c] : __float2bfloat16(0.0f);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, x = x0 + threadIdx.x;
        if (c < C && x < pitch) d[static_cast<size_t>(c) * H * pitch + x] = tile[threadIdx.x][r];
    }
}

// U[n][2y][2x][c] = src[n][y][x][c], zeros elsewhere (input of the stride-2 conv's data gradient)
__global__ void zero_insert_kernel(const __nv_bfloat16 *__restrict__ src, int N, int H, int W, int C, __nv_bfloat16 *__restrict__ dst) {
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;   // 8 channels (16 B) per thread
    const long long total = static_cast<long long>(N) * 2 * H * 2 * W * C;
    if (i >= total) return;
    const int c = static_cast<int>(i % C);
    long long t = i / C;
    const int X = static_cast<int>(t % (2 * W));
    t /= 2 * W;
    const int Y = static_cast<int>(t % (2 * H));
    const int n = static_cast<int>(t / (2 * H));
    uint4 v = make_uint4(0, 0, 0, 0);
    if (!(X & 1) && !(Y & 1)) v = *reinterpret_cast<const uint4 *>(src + ((static_cast<size_t>(n) * H + Y / 2) * W + X / 2) * C + c);
    *reinterpret_cast<uint4 *>(dst + i) = v;
}

__global__ void add3_kernel(const __nv_bfloat16 *__restrict__ a, const __nv_bfloat16 *__restrict__ b,
                            const __nv_bfloat16 *__restrict__ c, long long n, __nv_bfloat16 *__restrict__ out) {
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
    if (i >= n) return;
    const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162 *>(a + i);
    const __nv_bfloat162 y = *reinterpret_cast<const __nv_bfloat162 *>(b + i);
    float s0 = __bfloat162float(x.x) + __bfloat162float(y.x), s1 = __bfloat162float(x.y) + __bfloat162float(y.y);
    if (c) {
        const __nv_bfloat162 z = *reinterpret_cast<const __nv_bfloat162 *>(c + i);
        s0 += __bfloat162float(z.x);
        s1 += __bfloat162float(z.y);
    }
    *reinterpret_cast<__nv_bfloat162 *>(out + i) = __floats2bfloat162_rn(s0, s1);
}

// first conv (Cin = 1) weight gradient: dW[co][t] = sum_{n,y,x} dz[n][y][x][co] * x[n][y+ky-1][x+kx-1]
__global__ void __launch_bounds__(256)
first_conv_wgrad_kernel(const float *__restrict__ x, const __nv_bfloat16 *__restrict__ dz, int N, int H, int W, int C, float *dw) {
    // thread = (channel c = threadIdx.x % C, pixel lane); each block walks a strided set of pixels
    extern __shared__ float s_acc[];       // [C][9]
    for (int i = threadIdx.x; i < C * 9; i += blockDim.x) s_acc[i] = 0.0f;
    __syncthreads();
    const int c = threadIdx.x % C, pl = threadIdx.x / C, lanes = blockDim.x / C;
    float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long P = static_cast<long long>(N) * H * W;
    for (long long p = static_cast<long long>(blockIdx.x) * lanes + pl; p < P; p += static_cast<long long>(gridDim.x) * lanes) {
        const int xx = static_cast<int>(p % W);
        const int yy = static_cast<int>((p / W) % H);
        const long long n = p / (static_cast<long long>(W) * H);
        const float g = __bfloat162float(dz[p * C + c]);
        if (g == 0.0f) continue;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int y2 = yy + t / 3 - 1, x2 = xx + t % 3 - 1;
            if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) acc[t] = fmaf(g, x[(n * H + y2) * W + x2], acc[t]);
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) atomicAdd(&s_acc[c * 9 + t], acc[t]);
    __syncthreads();
    for (int i = threadIdx.x; i < C * 9; i += blockDim.x) atomicAdd(&dw[i], s_acc[i]);
}

inline int grid_for(long long work, int per_block, int cap) {
    long long b = (work + per_block - 1) / per_block;
    if (b < 1) b = 1;
    return static_cast<int>(b > cap ? cap : b);
}

}  // namespace

extern "C" int mbs_bn_train_fwd(const void *a, long long M, int C, const float *gamma, const float *beta, float eps, void *y,
                                float *sums_scratch, float *mean, float *invstd, float *var_unbiased, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(M > 0 && C > 0 && C % 2 == 0, "bn_train_fwd: bad shape");
    MBS_CHECK_CUDA(cudaMemsetAsync(sums_scratch, 0, 2 * C * sizeof(float), stream));
    bn_stats_kernel<<<grid_for(M, 8 * 16, 148 * 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(a), M, C, sums_scratch);
    MBS_CHECK_LAUNCH();
    bn_finalize_kernel<<<mbs::cdiv(C, 128), 128, 0, stream>>>(sums_scratch, M, C, eps, mean, invstd, var_unbiased);
    MBS_CHECK_LAUNCH();
    const long long total = M * C;
    bn_apply_kernel<<<static_cast<int>((total / 2 + 255) / 256), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16 *>(a), total, C, mean, invstd, gamma, beta, static_cast<__nv_bfloat16 *>(y));
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_bn_train_bwd(const void *dy, const void *a, long long M, int C, const float *mean, const float *invstd,
                                const float *gamma, int act, void *dz, float *dgamma_dbeta, float *dbias, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(M > 0 && C > 0, "bn_train_bwd: bad shape");
    MBS_REQUIRE(act == MBS_ACT_NONE || act == MBS_ACT_RELU, "bn_train_bwd: only relu / none are supported in training");
    MBS_CHECK_CUDA(cudaMemsetAsync(dgamma_dbeta, 0, 2 * C * sizeof(float), stream));
    MBS_CHECK_CUDA(cudaMemsetAsync(dbias, 0, C * sizeof(float), stream));
    const int grid = grid_for(M, 8 * 16, 148 * 8);
    bn_bwd_reduce_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(dy), static_cast<const __nv_bfloat16 *>(a), M,
                                                   C, mean, invstd, dgamma_dbeta);
    MBS_CHECK_LAUNCH();
    bn_bwd_apply_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(dy), static_cast<const __nv_bfloat16 *>(a), M,
                                                  C, mean, invstd, gamma, dgamma_dbeta, act, static_cast<__nv_bfloat16 *>(dz),
                                                  dbias);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_head_fwd(const void *y, long long M, int C, const float *w, float b, float *pred, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(M > 0 && C > 0 && C % 2 == 0, "head_fwd: bad shape");
    head_fwd_kernel<<<static_cast<int>((M + 255) / 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(y), M, C, w, b, pred);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_smoothl1(const float *pred, const float *target, long long M, float *loss_accum, float *grad, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(M > 0, "smoothl1: bad shape");
    smoothl1_kernel<<<grid_for(M, 256 * 8, 148 * 8), 256, 0, stream>>>(pred, target, M, loss_accum, grad);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_head_bwd(const float *g, const void *y, long long M, int C, const float *w, void *dy, float *dw_db,
                            void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(M > 0 && C > 0, "head_bwd: bad shape");
    MBS_CHECK_CUDA(cudaMemsetAsync(dw_db, 0, (C + 1) * sizeof(float), stream));
    head_bwd_kernel<<<grid_for(M, 8 * 16, 148 * 8), 256, 0, stream>>>(g, static_cast<const __nv_bfloat16 *>(y), M, C, w,
                                                                      static_cast<__nv_bfloat16 *>(dy), dw_db);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_nhwc_to_chw(const void *src, int N, int H, int W, int C, int pitch, int shift_x, int step_x, void *dst,
                               void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && (step_x == 1 || step_x == 2) && pitch * step_x >= W && pitch % 8 == 0,
                "nhwc_to_chw: bad shape / pitch");
    MBS_REQUIRE(static_cast<long long>(N) * H <= 65535, "nhwc_to_chw: N*H too large for one launch");
    dim3 grid(mbs::cdiv(pitch, 32), mbs::cdiv(C, 32), N * H), block(32, 8);
    nhwc_to_chw_kernel<<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16 *>(src), H, W, C, pitch, shift_x, step_x,
                                                   static_cast<__nv_bfloat16 *>(dst));
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_zero_insert_up2(const void *src, int N, int H, int W, int C, void *dst, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "zero_insert_up2: bad shape");
    const long long total = static_cast<long long>(N) * 4 * H * W * C;
    zero_insert_kernel<<<static_cast<int>((total / 8 + 255) / 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(src), N, H, W, C,
                                                                                     static_cast<__nv_bfloat16 *>(dst));
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_add3_bf16(const void *a, const void *b, const void *c, long long n, void *out, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(n > 0 && n % 2 == 0, "add3: bad size");
    add3_kernel<<<static_cast<int>((n / 2 + 255) / 256), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16 *>(a), static_cast<const __nv_bfloat16 *>(b), static_cast<const __nv_bfloat16 *>(c), n,
        static_cast<__nv_bfloat16 *>(out));
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_first_conv_wgrad(const float *x, const void *dz, int N, int H, int W, int C, float *dw, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(C > 0 && 256 % C == 0, "first_conv_wgrad: C must divide 256");
    MBS_CHECK_CUDA(cudaMemsetAsync(dw, 0, static_cast<size_t>(C) * 9 * sizeof(float), stream));
    first_conv_wgrad_kernel<<<148 * 4, 256, C * 9 * sizeof(float), stream>>>(x, static_cast<const __nv_bfloat16 *>(dz), N, H, W, C, dw);
    MBS_CHECK_LAUNCH();
    return 0;
}
