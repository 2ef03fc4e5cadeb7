// Elementwise / reduction kernels of the training step (HBM bound).
//
// Replace what torch autograd + cuDNN/ATen do around the convolutions in the reference's inner training
// loop (src/training/train.py:460-493): BatchNorm2d in training mode (batch statistics, unets.py:128,153,206,246),
// the activation derivative, the 1x1 heads, SmoothL1Loss (losses.py:30-32, beta = 1, 'mean') and their backward
// passes.  Activations and activation gradients are NHWC bf16, statistics / parameter gradients fp32.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "../../include/mbseg.h"
#include "common.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Mish (unets.py:81-89: x * tanh(softplus(x)), softplus threshold 20) with ONE exponential:
//   tanh(log(1 + e)) = ((1+e)^2 - 1) / ((1+e)^2 + 1) = n / (n + 2),  n = e * (e + 2),  e = exp(x)
// (these passes are HBM bound only if the special-function unit is not asked for exp + log + tanh per element).
__device__ __forceinline__ float mish_f(float x) {
    if (x > 20.0f) return x;
    const float e = __expf(x);
    const float n = e * (e + 2.0f);
    return x * __fdividef(n, n + 2.0f);
}
// d/dx mish = tanh(sp) + x * (1 - tanh(sp)^2) * sigmoid(x)
__device__ __forceinline__ float mish_grad_f(float x) {
    if (x > 20.0f) return 1.0f;
    const float e = __expf(x);
    const float n = e * (e + 2.0f);
    const float t = __fdividef(n, n + 2.0f);
    const float sg = __fdividef(e, e + 1.0f);
    return fmaf(x * (1.0f - t * t), sg, t);
}
// Training keeps ONE activation-sized tensor per layer.  act == MBS_ACT_MISH: it is the pre-activation z (needed by
// the derivative) and a = mish(z) is recomputed on load; otherwise it is the post-activation a itself.
template <bool MISH>
__device__ __forceinline__ void act_on_load(float (&f)[8]) {
    if (MISH) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = mish_f(f[j]);
    }
}

// Per-channel reductions over an [M][C] bf16 matrix with fully coalesced 16-byte accesses: C/8 consecutive threads
// cover one pixel row (8 channels each), a 256-thread block covers 2048/C pixels per iteration, every thread keeps
// its 8 channels' partial sums in registers; partials are combined through shared memory and one atomicAdd per
// (block, channel).  F(p, c0, vals_in..., acc) is supplied per kernel.  Requires C % 8 == 0 and C <= 2048.
struct Bf16x8 {
    uint4 raw;
    __device__ __forceinline__ void unpack(float (&f)[8]) const {
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __bfloat162float(h[i].x);
            f[2 * i + 1] = __bfloat162float(h[i].y);
        }
    }
    __device__ __forceinline__ static uint4 pack(const float (&f)[8]) {
        uint4 r;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        return r;
    }
};

__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

// Walk the pixels p = p0, p0 + stride, ... < M four at a time: the four 16-byte loads (per operand) are issued
// before any of them is consumed, so each thread keeps 64-128 bytes in flight (these kernels are HBM bound).
template <typename F>
__device__ __forceinline__ void for_rows4(const __nv_bfloat16 *a, long long p0, long long stride, long long M, int C, int c0, F f) {
    long long p = p0;
    for (; p + 3 * stride < M; p += 4 * stride) {
        const uint4 r0 = ldg16(a + p * C + c0), r1 = ldg16(a + (p + stride) * C + c0);
        const uint4 r2 = ldg16(a + (p + 2 * stride) * C + c0), r3 = ldg16(a + (p + 3 * stride) * C + c0);
        f(p, r0);
        f(p + stride, r1);
        f(p + 2 * stride, r2);
        f(p + 3 * stride, r3);
    }
    for (; p < M; p += stride) f(p, ldg16(a + p * C + c0));
}
template <typename F>
__device__ __forceinline__ void for_rows4x2(const __nv_bfloat16 *a, const __nv_bfloat16 *b, long long p0, long long stride,
                                            long long M, int C, int c0, F f) {
    long long p = p0;
    for (; p + 3 * stride < M; p += 4 * stride) {
        const uint4 a0 = ldg16(a + p * C + c0), a1 = ldg16(a + (p + stride) * C + c0);
        const uint4 a2 = ldg16(a + (p + 2 * stride) * C + c0), a3 = ldg16(a + (p + 3 * stride) * C + c0);
        const uint4 b0 = ldg16(b + p * C + c0), b1 = ldg16(b + (p + stride) * C + c0);
        const uint4 b2 = ldg16(b + (p + 2 * stride) * C + c0), b3 = ldg16(b + (p + 3 * stride) * C + c0);
        f(p, a0, b0);
        f(p + stride, a1, b1);
        f(p + 2 * stride, a2, b2);
        f(p + 3 * stride, a3, b3);
    }
    for (; p < M; p += stride) f(p, ldg16(a + p * C + c0), ldg16(b + p * C + c0));
}

// PARTIAL = true: the block's sums are stored to out[blockIdx.x][NACC][C] (summed later by reduce_partials_kernel:
// deterministic, and no same-address atomics from a thousand blocks -- those serialise in L2 and cost 10-25 us per
// BatchNorm layer); PARTIAL = false: atomicAdd into out[NACC][C].
template <int NACC, bool PARTIAL = false>
__device__ __forceinline__ void block_channel_atomic(float (&acc)[NACC][8], int C, int c0, float *out) {
    // threads with the same c0 (same threadIdx.x % (C/8)) hold partials of the same channels.  Every thread parks its
    // partials in its own shared-memory row ([pixel slot][channel], 256 * 8 floats per accumulator) and the block sums
    // the slots in a fixed order: no shared-memory atomics, so the result does not depend on warp scheduling.
    __shared__ float s_acc[NACC][2048];
    const int tpp = C / 8;                        // threads per pixel
    const int ppb = 256 / tpp;                    // pixel slots per block
    const int slot = threadIdx.x / tpp;
    __syncthreads();                              // callers may reuse this function's shared memory back to back
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (slot < ppb) s_acc[a][slot * C + (c0 < C ? c0 : 0) + j] = c0 < C ? acc[a][j] : 0.0f;
    __syncthreads();
    for (int i = threadIdx.x; i < NACC * C; i += blockDim.x) {
        const int a = i / C, c = i % C;
        float t = 0.0f;
        for (int sl = 0; sl < ppb; ++sl) t += s_acc[a][sl * C + c];
        if (PARTIAL)
            out[static_cast<size_t>(blockIdx.x) * NACC * C + i] = t;
        else
            atomicAdd(&out[a * C + c], t);
    }
}

template <bool MISH>
__global__ void __launch_bounds__(256) bn_stats_kernel(const __nv_bfloat16 *__restrict__ a, long long M, int C, float *sums) {
    const int tpp = C / 8, c0 = (threadIdx.x % tpp) * 8, ppb = 256 / tpp;
    float acc[2][8] = {};
    if (threadIdx.x < ppb * tpp) {
        for_rows4(a, static_cast<long long>(blockIdx.x) * ppb + threadIdx.x / tpp, static_cast<long long>(gridDim.x) * ppb, M, C, c0,
                  [&](long long, const uint4 &raw) {
                      Bf16x8 v;
                      v.raw = raw;
                      float f[8];
                      v.unpack(f);
                      act_on_load<MISH>(f);
#pragma unroll
                      for (int j = 0; j < 8; ++j) {
                          acc[0][j] += f[j];
                          acc[1][j] = fmaf(f[j], f[j], acc[1][j]);
                      }
                  });
    }
    block_channel_atomic<2, true>(acc, C, threadIdx.x < ppb * tpp ? c0 : C, sums);
}

// out[i] = sum over blocks of part[b][i].  32 consecutive i per CTA (128-byte rows), 32 thread rows split the blocks
// (<= 19 independent loads per thread: these launches are pure latency) and combine through shared memory in a fixed
// order (deterministic).
constexpr int RP_ROWS = 32;
__device__ __forceinline__ float column_partial(const float *__restrict__ part, int nblocks, int n, int i, int ty) {
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    int b = ty;
    for (; b + 3 * RP_ROWS < nblocks; b += 4 * RP_ROWS) {
        s0 += part[static_cast<size_t>(b) * n + i];
        s1 += part[static_cast<size_t>(b + RP_ROWS) * n + i];
        s2 += part[static_cast<size_t>(b + 2 * RP_ROWS) * n + i];
        s3 += part[static_cast<size_t>(b + 3 * RP_ROWS) * n + i];
    }
    for (; b < nblocks; b += RP_ROWS) s0 += part[static_cast<size_t>(b) * n + i];
    return (s0 + s1) + (s2 + s3);
}
__global__ void __launch_bounds__(32 * RP_ROWS) reduce_partials_kernel(const float *__restrict__ part, int nblocks, int n, float *__restrict__ out) {
    __shared__ float s_part[RP_ROWS][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx;
    s_part[ty][tx] = i < n ? column_partial(part, nblocks, n, i, ty) : 0.0f;
    __syncthreads();
    if (ty == 0 && i < n) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < RP_ROWS; ++k) t += s_part[k][tx];
        out[i] = t;
    }
}

// reduce_partials_kernel (both accumulators of 32 channels per CTA, same summation order) + bn_finalize_kernel in one
// launch: part[b][2][C] per-block sums -> sums[c] = gamma*invstd, sums[C+c] = beta - mean*gamma*invstd, mean, invstd and
// the running statistics.
__global__ void __launch_bounds__(32 * RP_ROWS)
bn_reduce_finalize_kernel(const float *__restrict__ part, int nblocks, float *sums, long long M, int C, float eps,
                          const float *__restrict__ gamma, const float *__restrict__ beta, float *mean, float *invstd, float momentum,
                          float *running_mean, float *running_var, long long *num_batches_tracked) {
    __shared__ float s_part[2][RP_ROWS][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const int n = 2 * C;
    s_part[0][ty][tx] = c < C ? column_partial(part, nblocks, n, c, ty) : 0.0f;
    s_part[1][ty][tx] = c < C ? column_partial(part, nblocks, n, C + c, ty) : 0.0f;
    __syncthreads();
    if (ty != 0 || c >= C) return;
    float t0 = 0.0f, t1 = 0.0f;
#pragma unroll
    for (int k = 0; k < RP_ROWS; ++k) { t0 += s_part[0][k][tx]; t1 += s_part[1][k][tx]; }
    const double m = static_cast<double>(t0) / static_cast<double>(M);
    double var = static_cast<double>(t1) / static_cast<double>(M) - m * m;
    if (var < 0.0) var = 0.0;
    const float mf = static_cast<float>(m), isf = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    mean[c] = mf;
    invstd[c] = isf;
    if (running_mean) {
        const float var_u = static_cast<float>(M > 1 ? var * static_cast<double>(M) / static_cast<double>(M - 1) : var);
        running_mean[c] = running_mean[c] * (1.0f - momentum) + mf * momentum;
        running_var[c] = running_var[c] * (1.0f - momentum) + var_u * momentum;
        if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
    }
    const float k = gamma[c] * isf;
    sums[c] = k;
    sums[C + c] = fmaf(-mf, k, beta[c]);
}

// GroupNorm(groups, C) / InstanceNorm2d(C) of ONE sample in eval mode (unets.py:129-132): statistics over the sample's own
// M = H*W pixels and the C/groups channels of a group (groups == C: instance norm), biased variance, eps inside the
// sqrt; gamma / beta optional (InstanceNorm2d's default has none).  Turns the per-channel sums into the affine of the
// apply pass: sums[c] = gamma*invstd, sums[C+c] = beta - mean*gamma*invstd.
__global__ void group_norm_finalize_kernel(float *sums, long long M, int C, int groups, float eps, const float *__restrict__ gamma,
                                           const float *__restrict__ beta) {
    extern __shared__ float s_grp[];          // [2][groups]
    const int cpg = C / groups;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int j = 0; j < cpg; ++j) {
            s1 += static_cast<double>(sums[g * cpg + j]);
            s2 += static_cast<double>(sums[C + g * cpg + j]);
        }
        const double cnt = static_cast<double>(M) * cpg;
        const double m = s1 / cnt;
        double var = s2 / cnt - m * m;
        if (var < 0.0) var = 0.0;
        s_grp[g] = static_cast<float>(m);
        s_grp[groups + g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float k = (gamma ? gamma[c] : 1.0f) * s_grp[groups + g];
        sums[c] = k;
        sums[C + c] = fmaf(-s_grp[g], k, beta ? beta[c] : 0.0f);
    }
}

// y = a * scale + shift with the row-vectorised mapping (8 channels per thread, affine in registers)
template <bool MISH>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const __nv_bfloat16 *__restrict__ a, long long M, int C, const float *__restrict__ scale_shift,
                __nv_bfloat16 *__restrict__ y) {
    const int tpp = C / 8, c0 = (threadIdx.x % tpp) * 8, ppb = 256 / tpp;
    if (threadIdx.x >= ppb * tpp) return;
    float k[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        k[j] = scale_shift[c0 + j];
        b[j] = scale_shift[C + c0 + j];
    }
    for_rows4(a, static_cast<long long>(blockIdx.x) * ppb + threadIdx.x / tpp, static_cast<long long>(gridDim.x) * ppb, M, C, c0,
              [&](long long p, const uint4 &raw) {
                  Bf16x8 v;
                  v.raw = raw;
                  float f[8];
                  v.unpack(f);
                  act_on_load<MISH>(f);
#pragma unroll
                  for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], k[j], b[j]);
                  *reinterpret_cast<uint4 *>(y + p * C + c0) = Bf16x8::pack(f);
              });
}

template <bool MISH>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const __nv_bfloat16 *__restrict__ dy, const __nv_bfloat16 *__restrict__ a, long long M, int C,
                     const float *__restrict__ mean, const float *__restrict__ invstd, float *dgamma_dbeta) {
    const int tpp = C / 8, c0 = (threadIdx.x % tpp) * 8, ppb = 256 / tpp;
    float acc[2][8] = {};
    if (threadIdx.x < ppb * tpp) {
        float mu[8], is[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { mu[j] = mean[c0 + j]; is[j] = invstd[c0 + j]; }
        for_rows4x2(dy, a, static_cast<long long>(blockIdx.x) * ppb + threadIdx.x / tpp, static_cast<long long>(gridDim.x) * ppb, M, C,
                    c0, [&](long long, const uint4 &rg, const uint4 &ra) {
                        Bf16x8 vg, va;
                        vg.raw = rg;
                        va.raw = ra;
                        float g[8], x[8];
                        vg.unpack(g);
                        va.unpack(x);
                        act_on_load<MISH>(x);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            acc[0][j] = fmaf(g[j], (x[j] - mu[j]) * is[j], acc[0][j]);
                            acc[1][j] += g[j];
                        }
                    });
    }
    block_channel_atomic<2, true>(acc, C, threadIdx.x < ppb * tpp ? c0 : C, dgamma_dbeta);
}

// dz = act'(a) * gamma * invstd * (dy - dbeta/M - xhat * dgamma/M);  dbias[c] += sum_p dz
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const __nv_bfloat16 *__restrict__ dy, const __nv_bfloat16 *__restrict__ a, long long M, int C,
                    const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ gamma,
                    const float *__restrict__ dgamma_dbeta, int act, __nv_bfloat16 *__restrict__ dz, float *dbias) {
    const int tpp = C / 8, c0 = (threadIdx.x % tpp) * 8, ppb = 256 / tpp;
    const float invM = 1.0f / static_cast<float>(M);
    float acc[1][8] = {};
    if (threadIdx.x < ppb * tpp) {
        float mu[8], is[8], k1[8], k2[8], k3[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mu[j] = mean[c0 + j];
            is[j] = invstd[c0 + j];
            k1[j] = gamma[c0 + j] * is[j];
            k2[j] = dgamma_dbeta[C + c0 + j] * invM;
            k3[j] = dgamma_dbeta[c0 + j] * invM;
        }
        for_rows4x2(dy, a, static_cast<long long>(blockIdx.x) * ppb + threadIdx.x / tpp, static_cast<long long>(gridDim.x) * ppb, M, C,
                    c0, [&](long long p, const uint4 &rg, const uint4 &ra) {
                        Bf16x8 vg, va;
                        vg.raw = rg;
                        va.raw = ra;
                        float g[8], x[8], o[8];
                        vg.unpack(g);
                        va.unpack(x);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float xa = act == MBS_ACT_MISH ? mish_f(x[j]) : x[j];      // post-activation value
                            float v = k1[j] * (g[j] - k2[j] - (xa - mu[j]) * is[j] * k3[j]);
                            if (act == MBS_ACT_RELU && !(x[j] > 0.0f)) v = 0.0f;
                            if (act == MBS_ACT_MISH) v *= mish_grad_f(x[j]);
                            o[j] = v;
                        }
                        const uint4 packed = Bf16x8::pack(o);
                        *reinterpret_cast<uint4 *>(dz + p * C + c0) = packed;
                        Bf16x8 back;
                        back.raw = packed;
                        float r[8];
                        back.unpack(r);
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[0][j] += r[j];
                    });
    }
    block_channel_atomic<1, true>(acc, C, threadIdx.x < ppb * tpp ? c0 : C, dbias);
}

// 1x1 head forward: pred[p] = sum_c y[p][c] * w[c] + b.  C/8 consecutive lanes share a pixel (16-byte loads, a warp
// reads 512 contiguous bytes) and combine with shuffles; C/8 must be a power of two <= 32.
__global__ void __launch_bounds__(256) head_fwd_kernel(const __nv_bfloat16 *__restrict__ y, long long M, int C, const float *__restrict__ w,
                                                       const float *__restrict__ b, float *__restrict__ pred) {
    const int tpp = C / 8, c0 = (threadIdx.x % tpp) * 8, ppb = 256 / tpp;
    float wv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) wv[j] = w[c0 + j];
    const float bias = b[0];
    for (long long p = static_cast<long long>(blockIdx.x) * ppb + threadIdx.x / tpp; p < M; p += static_cast<long long>(gridDim.x) * ppb) {
        Bf16x8 v;
        v.raw = ldg16(y + p * C + c0);
        float f[8];
        v.unpack(f);
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(f[j], wv[j], acc);
        for (int o = tpp >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (c0 == 0) pred[p] = acc + bias;
    }
}

// The distance-method criteria of losses.py:24-35 with reduction 'mean': loss += sum / M and g = dloss/dpred.
//   kind 0 SmoothL1Loss (beta = 1): g = clamp(d, -1, 1) / M;  1 L1Loss: g = sign(d) / M;  2 MSELoss: g = 2 d / M
__global__ void __launch_bounds__(256)
smoothl1_kernel(const float *__restrict__ pred, const float *__restrict__ target, long long M, float *loss, float *__restrict__ g,
                int kind) {
    __shared__ float s_part[8];
    float acc = 0.0f;
    const float invM = 1.0f / static_cast<float>(M);
    for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < M;
         p += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float d = pred[p] - target[p];
        const float ad = fabsf(d);
        if (kind == 0) {
            acc += ad < 1.0f ? 0.5f * d * d : ad - 0.5f;
            g[p] = fminf(fmaxf(d, -1.0f), 1.0f) * invM;
        } else if (kind == 1) {
            acc += ad;
            g[p] = (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) * invM;       // torch: sign(0) = 0
        } else {
            acc += d * d;
            g[p] = 2.0f * d * invM;
        }
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

// Criteria of the boundary method (losses.py:16-21, 71-96), 3 classes, logits planar [3][M], labels uint8 [M]:
//   'ce'      : nn.CrossEntropyLoss()  = mean over pixels of -log softmax(z)[y]
//   'ce_dice' : ce + 0.5 * sum_{c=1,2} c * dice_c,  dice_c = 1 - (2 sum(g_c p_c) + 1) / (sum(g_c^2) + sum(p_c^2) + 1)
//               with p = softmax(z), g = one_hot(y), sums over the WHOLE batch (losses.py:58-63, 90-94).
// Pass 1 accumulates the seven sums, pass 2 writes dloss/dz and the loss value.
// sums: [0] ce, [1..2] sum g_c p_c, [3..4] sum p_c^2, [5..6] sum g_c   (c = 1, 2)
__device__ __forceinline__ void softmax3(float z0, float z1, float z2, float &p0, float &p1, float &p2, float &lse) {
    const float m = fmaxf(z0, fmaxf(z1, z2));
    const float e0 = expf(z0 - m), e1 = expf(z1 - m), e2 = expf(z2 - m);
    const float s = e0 + e1 + e2;
    p0 = e0 / s; p1 = e1 / s; p2 = e2 / s;
    lse = m + logf(s);
}

__global__ void __launch_bounds__(256)
ce_dice_sums_kernel(const float *__restrict__ z, const uint8_t *__restrict__ y, long long M, double *sums) {
    float a[7] = {0, 0, 0, 0, 0, 0, 0};
    for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < M; p += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float z0 = z[p], z1 = z[M + p], z2 = z[2 * M + p];
        float p0, p1, p2, lse;
        softmax3(z0, z1, z2, p0, p1, p2, lse);
        const int c = y[p];
        a[0] += lse - (c == 0 ? z0 : (c == 1 ? z1 : z2));
        a[1] += c == 1 ? p1 : 0.0f;
        a[2] += c == 2 ? p2 : 0.0f;
        a[3] += p1 * p1;
        a[4] += p2 * p2;
        a[5] += c == 1 ? 1.0f : 0.0f;
        a[6] += c == 2 ? 1.0f : 0.0f;
    }
    __shared__ float s_part[8][7];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const float v = warp_sum(a[k]);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += static_cast<double>(s_part[w][threadIdx.x]);
        atomicAdd(&sums[threadIdx.x], t);
    }
}

__global__ void __launch_bounds__(256)
ce_dice_grad_kernel(const float *__restrict__ z, const uint8_t *__restrict__ y, long long M, int with_dice, const double *__restrict__ sums,
                    float *loss, float *__restrict__ g) {
    const float invM = 1.0f / static_cast<float>(M);
    // dice_c = 1 - N_c / D_c;  d dice_c / d p_c(pixel) = -(2 g_c D_c - 2 p_c N_c) / D_c^2;  weight 0.5 * c
    float N1 = 0, N2 = 0, D1 = 1, D2 = 1;
    if (with_dice) {
        N1 = static_cast<float>(2.0 * sums[1] + 1.0);
        N2 = static_cast<float>(2.0 * sums[2] + 1.0);
        D1 = static_cast<float>(sums[5] + sums[3] + 1.0);
        D2 = static_cast<float>(sums[6] + sums[4] + 1.0);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float l = static_cast<float>(sums[0]) * invM;
        if (with_dice) l += 0.5f * (1.0f * (1.0f - N1 / D1) + 2.0f * (1.0f - N2 / D2));
        atomicAdd(loss, l);
    }
    for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < M; p += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float z0 = z[p], z1 = z[M + p], z2 = z[2 * M + p];
        float p0, p1, p2, lse;
        softmax3(z0, z1, z2, p0, p1, p2, lse);
        const int c = y[p];
        float g0 = (p0 - (c == 0 ? 1.0f : 0.0f)) * invM, g1 = (p1 - (c == 1 ? 1.0f : 0.0f)) * invM, g2 = (p2 - (c == 2 ? 1.0f : 0.0f)) * invM;
        if (with_dice) {
            const float dp1 = -0.5f * 1.0f * (2.0f * (c == 1 ? 1.0f : 0.0f) * D1 - 2.0f * p1 * N1) / (D1 * D1);
            const float dp2 = -0.5f * 2.0f * (2.0f * (c == 2 ? 1.0f : 0.0f) * D2 - 2.0f * p2 * N2) / (D2 * D2);
            // softmax Jacobian: dz_j = sum_c dp_c * p_c * (delta_cj - p_j)
            const float t = dp1 * p1 + dp2 * p2;
            g0 += -p0 * t;
            g1 += dp1 * p1 - p1 * t;
            g2 += dp2 * p2 - p2 * t;
        }
        g[p] = g0;
        g[M + p] = g1;
        g[2 * M + p] = g2;
    }
}

// head backward: dy[p][c] = g[p] * w[c];  dw[c] += sum_p g[p] * y[p][c];  dw[C] (= db) += sum_p g[p]
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float *__restrict__ g, const __nv_bfloat16 *__restrict__ y, long long M, int C, const float *__restrict__ w,
                __nv_bfloat16 *__restrict__ dy, float *part_dw, float *part_db) {
    __shared__ float s_g[8];
    const int tpp = C / 8, c0 = (threadIdx.x % tpp) * 8, ppb = 256 / tpp;
    float acc[1][8] = {};
    float gsum = 0.0f;
    if (threadIdx.x < ppb * tpp) {
        float wv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) wv[j] = w[c0 + j];
        for (long long p = static_cast<long long>(blockIdx.x) * ppb + threadIdx.x / tpp; p < M; p += static_cast<long long>(gridDim.x) * ppb) {
            const float gp = g[p];
            Bf16x8 vy;
            vy.raw = *reinterpret_cast<const uint4 *>(y + p * C + c0);
            float yv[8], o[8];
            vy.unpack(yv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                o[j] = gp * wv[j];
                acc[0][j] = fmaf(gp, yv[j], acc[0][j]);
            }
            *reinterpret_cast<uint4 *>(dy + p * C + c0) = Bf16x8::pack(o);
            if (c0 == 0) gsum += gp;
        }
    }
    // per-block partial sums, combined in a fixed order by reduce_partials_kernel (no atomics: bitwise reproducible)
    block_channel_atomic<1, true>(acc, C, threadIdx.x < ppb * tpp ? c0 : C, part_dw);
    gsum = warp_sum(gsum);
    if ((threadIdx.x & 31) == 0) s_g[threadIdx.x >> 5] = gsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += s_g[k];
        part_db[blockIdx.x] = t;
    }
}

// Data-gradient filter of a 3x3 conv in GEMM-packed form: out[ci][t][co] = bf16(w[co][ci][8 - t]) (spatially flipped,
// in/out channels swapped) straight from the reference-layout fp32 weight [Cout][Cin][3][3].  32x32 (co, ci) tiles
// through shared memory: reads are 288-float runs, writes 64-byte runs.
__global__ void __launch_bounds__(256) pack_dgrad3x3_kernel(const float *__restrict__ w, int Cout, int Cin, __nv_bfloat16 *__restrict__ out) {
    __shared__ float tile[32][32 * 9 + 1];
    const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
    for (int i = threadIdx.x; i < 32 * 288; i += 256) {
        const int r = i / 288, k = i - r * 288;
        tile[r][k] = (co0 + r < Cout && ci0 + k / 9 < Cin) ? w[(static_cast<size_t>(co0 + r) * Cin + ci0) * 9 + k] : 0.0f;
    }
    __syncthreads();
    // 8 consecutive output channels (16 bytes) per thread
    for (int i = threadIdx.x; i < 32 * 9 * 4; i += 256) {
        const int c8 = (i & 3) * 8, t = (i >> 2) % 9, ci = i / 36;
        if (ci0 + ci >= Cin) continue;
        __nv_bfloat16 *dst = out + (static_cast<size_t>(ci0 + ci) * 9 + t) * Cout + co0 + c8;
        if (co0 + c8 + 8 <= Cout && (Cout & 7) == 0) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = tile[c8 + j][ci * 9 + (8 - t)];
            *reinterpret_cast<uint4 *>(dst) = Bf16x8::pack(f);
        } else {
            for (int j = 0; j < 8 && co0 + c8 + j < Cout; ++j) dst[j] = __float2bfloat16_rn(tile[c8 + j][ci * 9 + (8 - t)]);
        }
    }
}

// Every GEMM-packed bf16 weight of one training step in ONE launch (the per-layer pack kernels cost ~90 launches and
// 1.2 ms per step, mostly strided 2-byte writes).  A job is one conv layer; a CTA owns a 32 x 32 (co, ci) tile of it,
// reads the fp32 reference-layout weight once (runs of 32*taps floats) and writes both packed forms in 64-byte runs:
//   conv3x3  w[co][ci][3][3]:  fwd[co][t][ci] = w[co][ci][t]        dgrad[ci][t][co] = w[co][ci][8 - t]
//   convT2x2 w[ci][co][2][2]:  fwd[q*Cout + co][ci] = w[ci][co][q]  dgrad[ci][q][co] = w[ci][co][q]
__global__ void __launch_bounds__(256) pack_train_weights_kernel(const mbs_pack_job *__restrict__ jobs, int n_jobs) {
    __shared__ float tile[32][32 * 9 + 1];
    int j = 0;
    while (j + 1 < n_jobs && static_cast<int>(blockIdx.x) >= jobs[j + 1].tile0) ++j;
    const mbs_pack_job job = jobs[j];
    const int local = static_cast<int>(blockIdx.x) - job.tile0;
    const float *__restrict__ w = static_cast<const float *>(job.w);
    __nv_bfloat16 *fwd = static_cast<__nv_bfloat16 *>(job.fwd), *dg = static_cast<__nv_bfloat16 *>(job.dgrad);
    if (job.kind == 0) {
        const int Cout = job.cout, Cin = job.cin;
        const int tci = Cin / 32;
        const int co0 = (local / tci) * 32, ci0 = (local % tci) * 32;
        for (int i = threadIdx.x; i < 32 * 288; i += 256) {
            const int r = i / 288, k = i - r * 288;
            tile[r][k] = w[(static_cast<size_t>(co0 + r) * Cin + ci0) * 9 + k];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * 9 * 4; i += 256) {
            const int c8 = (i & 3) * 8, t = (i >> 2) % 9, r = i / 36;
            float f[8];
            if (fwd) {                       // row (co = r, tap t): 8 consecutive input channels
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] = tile[r][(c8 + q) * 9 + t];
                *reinterpret_cast<uint4 *>(fwd + (static_cast<size_t>(co0 + r) * 9 + t) * Cin + ci0 + c8) = Bf16x8::pack(f);
            }
            if (dg) {                        // row (ci = r, tap t): 8 consecutive output channels, filter flipped
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] = tile[c8 + q][r * 9 + (8 - t)];
                *reinterpret_cast<uint4 *>(dg + (static_cast<size_t>(ci0 + r) * 9 + t) * Cout + co0 + c8) = Bf16x8::pack(f);
            }
        }
    } else {
        const int Cout = job.cout, Cin = job.cin;          // w[Cin][Cout][2][2]
        const int tco = Cout / 32;
        const int ci0 = (local / tco) * 32, co0 = (local % tco) * 32;
        for (int i = threadIdx.x; i < 32 * 128; i += 256) {
            const int r = i >> 7, k = i & 127;
            tile[r][k] = w[(static_cast<size_t>(ci0 + r) * Cout + co0) * 4 + k];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * 4 * 4; i += 256) {
            const int c8 = (i & 3) * 8, q = (i >> 2) & 3, r = i >> 4;
            float f[8];
            if (fwd) {                       // row (q, co = r): 8 consecutive input channels
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = tile[c8 + k][r * 4 + q];
                *reinterpret_cast<uint4 *>(fwd + (static_cast<size_t>(q) * Cout + co0 + r) * Cin + ci0 + c8) = Bf16x8::pack(f);
            }
            if (dg) {                        // row (ci = r, q): 8 consecutive output channels
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = tile[r][(c8 + k) * 4 + q];
                *reinterpret_cast<uint4 *>(dg + (static_cast<size_t>(ci0 + r) * 4 + q) * Cout + co0 + c8) = Bf16x8::pack(f);
            }
        }
    }
}

// Weight gradient from the GEMM layout g[co][t][ci] to the reference layout out[co][ci][3][3] (fp32)
__global__ void __launch_bounds__(256) unpack_grad3x3_kernel(const float *__restrict__ g, int Cin, float *__restrict__ out) {
    __shared__ float tile[9][256 + 1];
    const int co = blockIdx.x, ci0 = blockIdx.y * 256;
    const int n = min(256, Cin - ci0);
    for (int i = threadIdx.x; i < 9 * 256; i += 256) {
        const int t = i >> 8, c = i & 255;
        if (c < n) tile[t][c] = g[(static_cast<size_t>(co) * 9 + t) * Cin + ci0 + c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * 9; i += 256) out[(static_cast<size_t>(co) * Cin + ci0) * 9 + i] = tile[i % 9][i / 9];
}

// Deterministic end of the weight gradient: out = sum over the K splits (in split order) of the partial tiles written by
// mbs_conv_wgrad(partial = 1), and the layout change to the reference's parameter layout in the same pass.
//   part_s: [splits_s][Cm][TAPS][cn_s] for the (up to two) sources of a concatenated input
//   LAYOUT 0: out[co][coff_s + ci][TAPS] (Conv2d weight [Cout][Cin][3][3]);  LAYOUT 1: out[ci][co][TAPS] (ConvTranspose2d
//   weight [Cin][Cout][2][2]).  One CTA = one output channel co and CI input channels; SG thread groups share the splits
//   and are combined through shared memory in a fixed order.
template <int TAPS, int CI, int SG, int LAYOUT>
__global__ void __launch_bounds__(CI * SG)
wgrad_reduce_kernel(const float *__restrict__ part0, int splits0, int cn0, const float *__restrict__ part1, int splits1, int cn1,
                    int Cm, float *__restrict__ out) {
    __shared__ float s_sum[SG][TAPS][CI + 1];          // SG = 16, TAPS = 9: 37 KB
    const int co = blockIdx.x;
    const int ci0 = blockIdx.y * CI;                       // channel offset in the concatenated input
    const bool second = ci0 >= cn0;
    const float *__restrict__ part = second ? part1 : part0;
    const int splits = second ? splits1 : splits0, cn = second ? cn1 : cn0;
    const int cl0 = second ? ci0 - cn0 : ci0;              // channel offset inside this source
    const int tx = threadIdx.x % CI, ty = threadIdx.x / CI;
    const bool live = cl0 + tx < cn;
    const size_t split_stride = static_cast<size_t>(Cm) * TAPS * cn;
    float acc[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) acc[t] = 0.0f;
    if (live) {
        const float *src = part + static_cast<size_t>(co) * TAPS * cn + cl0 + tx;
        for (int k = ty; k < splits; k += SG) {
#pragma unroll
            for (int t = 0; t < TAPS; ++t) acc[t] += src[k * split_stride + static_cast<size_t>(t) * cn];
        }
    }
#pragma unroll
    for (int t = 0; t < TAPS; ++t) s_sum[ty][t][tx] = acc[t];
    __syncthreads();
    const int cin_total = cn0 + cn1;
    for (int i = threadIdx.x; i < CI * TAPS; i += CI * SG) {
        const int c = i / TAPS, t = i - c * TAPS;
        if (cl0 + c >= cn) continue;
        float v = 0.0f;
#pragma unroll
        for (int g = 0; g < SG; ++g) v += s_sum[g][t][c];
        if (LAYOUT == 0)
            out[(static_cast<size_t>(co) * cin_total + ci0) * TAPS + i] = v;       // CI * TAPS contiguous floats per CTA
        else
            out[(static_cast<size_t>(ci0 + c) * Cm + co) * TAPS + t] = v;
    }
}

// nn.MaxPool2d(kernel_size=2, stride=2) on NHWC bf16 (pool_method = 'max', unets.py:306-307,363-364): 8 channels per thread
__global__ void maxpool2x2_kernel(const __nv_bfloat16 *__restrict__ src, int N, int H, int W, int C, __nv_bfloat16 *__restrict__ dst) {
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
    const int Ho = H / 2, Wo = W / 2;
    const long long total = static_cast<long long>(N) * Ho * Wo * C;
    if (i >= total) return;
    const int c = static_cast<int>(i % C);
    long long t = i / C;
    const int x = static_cast<int>(t % Wo);
    t /= Wo;
    const int y = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    const __nv_bfloat16 *p = src + ((static_cast<size_t>(n) * H + 2 * y) * W + 2 * x) * C + c;
    const uint4 a = *reinterpret_cast<const uint4 *>(p), b = *reinterpret_cast<const uint4 *>(p + C);
    const uint4 d = *reinterpret_cast<const uint4 *>(p + static_cast<size_t>(W) * C), e = *reinterpret_cast<const uint4 *>(p + static_cast<size_t>(W) * C + C);
    uint4 o;
    const __nv_bfloat162 *pa = reinterpret_cast<const __nv_bfloat162 *>(&a), *pb = reinterpret_cast<const __nv_bfloat162 *>(&b);
    const __nv_bfloat162 *pd = reinterpret_cast<const __nv_bfloat162 *>(&d), *pe = reinterpret_cast<const __nv_bfloat162 *>(&e);
    __nv_bfloat162 *po = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) po[j] = __hmax2_nan(__hmax2_nan(pa[j], pb[j]), __hmax2_nan(pd[j], pe[j]));     // NaN propagates like torch
    *reinterpret_cast<uint4 *>(dst + i) = o;
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

// eight elements (16 bytes) per thread and step, grid-stride (the two-element kernel below is the fallback for odd sizes /
// unaligned views); the same float additions in the same order, so results are identical
__global__ void __launch_bounds__(256)
add3_vec8_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b, const uint4 *__restrict__ c, long long n8,
                 uint4 *__restrict__ out) {
    auto add2 = [](unsigned x, unsigned y, unsigned z, bool has_c) {
        const __nv_bfloat162 xv = *reinterpret_cast<const __nv_bfloat162 *>(&x), yv = *reinterpret_cast<const __nv_bfloat162 *>(&y);
        float s0 = __bfloat162float(xv.x) + __bfloat162float(yv.x), s1 = __bfloat162float(xv.y) + __bfloat162float(yv.y);
        if (has_c) {
            const __nv_bfloat162 zv = *reinterpret_cast<const __nv_bfloat162 *>(&z);
            s0 += __bfloat162float(zv.x);
            s1 += __bfloat162float(zv.y);
        }
        const __nv_bfloat162 r = __floats2bfloat162_rn(s0, s1);
        return *reinterpret_cast<const unsigned *>(&r);
    };
    const bool has_c = c != nullptr;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 x = a[i], y = b[i];
        const uint4 z = has_c ? c[i] : make_uint4(0, 0, 0, 0);
        out[i] = make_uint4(add2(x.x, y.x, z.x, has_c), add2(x.y, y.y, z.y, has_c), add2(x.z, y.z, z.z, has_c),
                            add2(x.w, y.w, z.w, has_c));
    }
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

// first conv (Cin = 1) weight gradient: dW[co][t] = sum_{n,y,x} dz[n][y][x][co] * x[n][y+ky-1][x+kx-1].
// C/8 consecutive threads share a pixel (16-byte dz loads, the nine x taps are warp-broadcast loads), every thread keeps
// 9 x 8 partial sums in registers; the block combines them through shared memory in a fixed order and writes its partial
// to part[blockIdx.x][C*9] (summed by reduce_partials_kernel: deterministic, no atomics).
__global__ void __launch_bounds__(256)
first_conv_wgrad_kernel(const float *__restrict__ x, const __nv_bfloat16 *__restrict__ dz, int N, int H, int W, int C, float *part) {
    __shared__ float s_red[2048];                   // [pixel slot][C]: ppb * C = 2048 floats
    const int tpp = C / 8, c0 = (threadIdx.x % tpp) * 8, ppb = 256 / tpp;
    float acc[9][8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = 0.0f;
    const long long M = static_cast<long long>(N) * H * W;
    for (long long p = static_cast<long long>(blockIdx.x) * ppb + threadIdx.x / tpp; p < M; p += static_cast<long long>(gridDim.x) * ppb) {
        Bf16x8 v;
        v.raw = ldg16(dz + p * C + c0);
        float g[8];
        v.unpack(g);
        const int xx = static_cast<int>(p % W);
        const int yy = static_cast<int>((p / W) % H);
        const float *xc = x + p;                        // x is [N][H][W]: same flat pixel index
        const bool up = yy > 0, down = yy + 1 < H, left = xx > 0, right = xx + 1 < W;
        float xv[9];
        xv[0] = (up && left) ? __ldg(xc - W - 1) : 0.0f;
        xv[1] = up ? __ldg(xc - W) : 0.0f;
        xv[2] = (up && right) ? __ldg(xc - W + 1) : 0.0f;
        xv[3] = left ? __ldg(xc - 1) : 0.0f;
        xv[4] = __ldg(xc);
        xv[5] = right ? __ldg(xc + 1) : 0.0f;
        xv[6] = (down && left) ? __ldg(xc + W - 1) : 0.0f;
        xv[7] = down ? __ldg(xc + W) : 0.0f;
        xv[8] = (down && right) ? __ldg(xc + W + 1) : 0.0f;
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(g[j], xv[t], acc[t][j]);
    }
    const int slot = threadIdx.x / tpp;
    float *out = part + static_cast<size_t>(blockIdx.x) * C * 9;
#pragma unroll 1
    for (int t = 0; t < 9; ++t) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) s_red[slot * C + c0 + j] = acc[t][j];
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += 256) {
            float sum = 0.0f;
            for (int sl = 0; sl < ppb; ++sl) sum += s_red[sl * C + c];
            out[c * 9 + t] = sum;
        }
    }
}

constexpr int BN_MAX_BLOCKS = 148 * 4;
inline int grid_rows(long long M, int C) {          // blocks for the row-vectorised reductions: >= 8 rows per thread
    const long long ppb = 256 / (C / 8);
    long long b = (M + ppb * 8 - 1) / (ppb * 8);
    if (b < 1) b = 1;
    static int cap = 0;
    if (cap == 0) {
        const char *e = getenv("MBS_BN_BLOCKS");       // A/B knob, <= BN_MAX_BLOCKS
        cap = e ? atoi(e) : 2 * 148;                   // one resident wave at 2 blocks / SM (measured: 296 beats 148 and 592)
        if (cap < 1 || cap > BN_MAX_BLOCKS) cap = BN_MAX_BLOCKS;
    }
    return static_cast<int>(b > cap ? cap : b);
}
inline int grid_for(long long work, int per_block, int cap) {
    long long b = (work + per_block - 1) / per_block;
    if (b < 1) b = 1;
    return static_cast<int>(b > cap ? cap : b);
}

}  // namespace

extern "C" int mbs_bn_train_fwd(const void *a, long long M, int C, const float *gamma, const float *beta, float eps, void *y,
                                float *sums_scratch, float *mean, float *invstd, float momentum, float *running_mean,
                                float *running_var, long long *num_batches_tracked, int act, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(M > 0 && C >= 8 && C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, "bn_train_fwd: bad shape (C must be 8*2^k <= 2048)");
    const int grid = grid_rows(M, C);
    float *part = sums_scratch + 2 * C;          // [grid][2][C] per-block partial sums
    MBS_REQUIRE(act == MBS_ACT_NONE || act == MBS_ACT_MISH, "bn_train_fwd: act must be NONE (tensor is post-activation) or MISH (pre-activation)");
    if (act == MBS_ACT_MISH)
        bn_stats_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(a), M, C, part);
    else
        bn_stats_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(a), M, C, part);
    MBS_CHECK_LAUNCH();
    bn_reduce_finalize_kernel<<<mbs::cdiv(C, 32), 32 * RP_ROWS, 0, stream>>>(part, grid, sums_scratch, M, C, eps, gamma, beta, mean, invstd,
                                                                    momentum, running_mean, running_var, num_batches_tracked);
    MBS_CHECK_LAUNCH();
    if (act == MBS_ACT_MISH)
        bn_apply_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(a), M, C, sums_scratch,
                                                        static_cast<__nv_bfloat16 *>(y));
    else
        bn_apply_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(a), M, C, sums_scratch,
                                                         static_cast<__nv_bfloat16 *>(y));
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_sample_group_norm(const void *a, long long M, int C, int groups, const float *gamma, const float *beta, float eps,
                                     void *y, float *scratch, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(a && y && scratch && M > 0 && C >= 8 && C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0 && groups > 0 && C % groups == 0,
                "sample_group_norm: bad shape (C must be 8*2^k <= 2048 and divisible by groups)");
    const int grid = grid_rows(M, C);
    float *part = scratch + 2 * C;
    bn_stats_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(a), M, C, part);
    MBS_CHECK_LAUNCH();
    reduce_partials_kernel<<<mbs::cdiv(2 * C, 32), 32 * RP_ROWS, 0, stream>>>(part, grid, 2 * C, scratch);
    MBS_CHECK_LAUNCH();
    group_norm_finalize_kernel<<<1, 256, 2 * groups * sizeof(float), stream>>>(scratch, M, C, groups, eps, gamma, beta);
    MBS_CHECK_LAUNCH();
    bn_apply_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(a), M, C, scratch, static_cast<__nv_bfloat16 *>(y));
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t mbs_bn_scratch_floats(int C) { return static_cast<size_t>(C) * (2 + 2 * BN_MAX_BLOCKS); }

extern "C" int mbs_bn_train_bwd(const void *dy, const void *a, long long M, int C, const float *mean, const float *invstd,
                                const float *gamma, int act, void *dz, float *dgamma_dbeta, float *dbias, float *scratch,
                                void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(M > 0 && C >= 8 && C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, "bn_train_bwd: bad shape");
    MBS_REQUIRE(act == MBS_ACT_NONE || act == MBS_ACT_RELU || act == MBS_ACT_MISH,
                "bn_train_bwd: training supports relu (tensor = post-activation), mish (tensor = pre-activation) or none");
    const int grid = grid_rows(M, C);
    float *part = scratch + 2 * C;               // [grid][2][C], then reused as [grid][C]
    if (act == MBS_ACT_MISH)
        bn_bwd_reduce_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(dy),
                                                             static_cast<const __nv_bfloat16 *>(a), M, C, mean, invstd, part);
    else
        bn_bwd_reduce_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(dy),
                                                              static_cast<const __nv_bfloat16 *>(a), M, C, mean, invstd, part);
    MBS_CHECK_LAUNCH();
    reduce_partials_kernel<<<mbs::cdiv(2 * C, 32), 32 * RP_ROWS, 0, stream>>>(part, grid, 2 * C, dgamma_dbeta);
    MBS_CHECK_LAUNCH();
    bn_bwd_apply_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(dy), static_cast<const __nv_bfloat16 *>(a), M,
                                                  C, mean, invstd, gamma, dgamma_dbeta, act, static_cast<__nv_bfloat16 *>(dz),
                                                  part);
    MBS_CHECK_LAUNCH();
    reduce_partials_kernel<<<mbs::cdiv(C, 32), 32 * RP_ROWS, 0, stream>>>(part, grid, C, dbias);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_head_fwd(const void *y, long long M, int C, const float *w, const float *b_dev, float *pred, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(M > 0 && C >= 8 && C <= 256 && (C & (C - 1)) == 0, "head_fwd: C must be a power of two in [8, 256]");
    head_fwd_kernel<<<grid_rows(M, C), 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(y), M, C, w, b_dev, pred);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_smoothl1(const float *pred, const float *target, long long M, float *loss_accum, float *grad, void *stream_) {
    return mbs_regression_loss(pred, target, M, 0, loss_accum, grad, stream_);
}

extern "C" int mbs_regression_loss(const float *pred, const float *target, long long M, int kind, float *loss_accum, float *grad,
                                   void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(M > 0 && kind >= 0 && kind <= 2, "regression_loss: bad shape / kind");
    smoothl1_kernel<<<grid_for(M, 256 * 8, 148 * 8), 256, 0, stream>>>(pred, target, M, loss_accum, grad, kind);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_ce_dice_loss(const float *logits, const uint8_t *labels, long long M, int with_dice, float *loss_accum, float *grad,
                                double *sums7, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(M > 0 && logits && labels && loss_accum && grad && sums7, "ce_dice_loss: bad arguments");
    MBS_CHECK_CUDA(cudaMemsetAsync(sums7, 0, 7 * sizeof(double), stream));
    const int grid = grid_for(M, 256 * 8, 148 * 8);
    ce_dice_sums_kernel<<<grid, 256, 0, stream>>>(logits, labels, M, sums7);
    MBS_CHECK_LAUNCH();
    ce_dice_grad_kernel<<<grid, 256, 0, stream>>>(logits, labels, M, with_dice, sums7, loss_accum, grad);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_head_bwd(const float *g, const void *y, long long M, int C, const float *w, void *dy, float *dw_db,
                            float *scratch, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(M > 0 && C >= 8 && C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0 && scratch, "head_bwd: bad shape");
    const int grid = grid_rows(M, C);
    float *part_dw = scratch, *part_db = scratch + static_cast<size_t>(grid) * C;       // (C + 1) * grid floats
    head_bwd_kernel<<<grid, 256, 0, stream>>>(g, static_cast<const __nv_bfloat16 *>(y), M, C, w, static_cast<__nv_bfloat16 *>(dy), part_dw,
                                              part_db);
    MBS_CHECK_LAUNCH();
    reduce_partials_kernel<<<mbs::cdiv(C, 32), 32 * RP_ROWS, 0, stream>>>(part_dw, grid, C, dw_db);
    MBS_CHECK_LAUNCH();
    reduce_partials_kernel<<<1, 32 * RP_ROWS, 0, stream>>>(part_db, grid, 1, dw_db + C);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_pack_conv3x3_dgrad(const float *w, int Cout, int Cin, void *packed, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(w && packed && Cout > 0 && Cin > 0, "pack_conv3x3_dgrad: bad arguments");
    pack_dgrad3x3_kernel<<<dim3(mbs::cdiv(Cout, 32), mbs::cdiv(Cin, 32)), 256, 0, stream>>>(w, Cout, Cin,
                                                                                          static_cast<__nv_bfloat16 *>(packed));
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_pack_train_weights(const mbs_pack_job *jobs_dev, int n_jobs, int total_tiles, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(jobs_dev && n_jobs > 0 && total_tiles > 0, "pack_train_weights: bad arguments");
    pack_train_weights_kernel<<<total_tiles, 256, 0, stream>>>(jobs_dev, n_jobs);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_unpack_conv3x3_grad(const float *g, int Cout, int Cin, float *out, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(g && out && Cout > 0 && Cin > 0, "unpack_conv3x3_grad: bad arguments");
    unpack_grad3x3_kernel<<<dim3(Cout, mbs::cdiv(Cin, 256)), 256, 0, stream>>>(g, Cin, out);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_wgrad_reduce(const float *part0, int splits0, int cn0, const float *part1, int splits1, int cn1, int Cm, int layout,
                                float *out, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(part0 && out && splits0 > 0 && cn0 > 0 && cn0 % 64 == 0 && Cm > 0 && (layout == 0 || layout == 1),
                "wgrad_reduce: bad arguments");
    MBS_REQUIRE(part1 ? (splits1 > 0 && cn1 > 0 && cn1 % 64 == 0 && layout == 0) : (splits1 == 0 && cn1 == 0),
                "wgrad_reduce: bad second source");
    const dim3 grid(Cm, (cn0 + cn1) / 64);
    const bool many = splits0 >= 8 || splits1 >= 8;
    if (layout == 0) {
        if (many) wgrad_reduce_kernel<9, 64, 16, 0><<<grid, 1024, 0, stream>>>(part0, splits0, cn0, part1, splits1, cn1, Cm, out);
        else wgrad_reduce_kernel<9, 64, 1, 0><<<grid, 64, 0, stream>>>(part0, splits0, cn0, part1, splits1, cn1, Cm, out);
    } else {
        if (many) wgrad_reduce_kernel<4, 64, 16, 1><<<grid, 1024, 0, stream>>>(part0, splits0, cn0, nullptr, 0, 0, Cm, out);
        else wgrad_reduce_kernel<4, 64, 1, 1><<<grid, 64, 0, stream>>>(part0, splits0, cn0, nullptr, 0, 0, Cm, out);
    }
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_maxpool2x2(const void *src, int N, int H, int W, int C, void *dst, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(src && dst && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0, "maxpool2x2: bad shape");
    const long long total = static_cast<long long>(N) * (H / 2) * (W / 2) * C;
    maxpool2x2_kernel<<<static_cast<int>((total / 8 + 255) / 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(src), N, H, W, C,
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
    const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                         reinterpret_cast<uintptr_t>(out);
    if (n % 8 == 0 && (al & 15) == 0) {
        const long long n8 = n / 8;
        const long long want = (n8 + 255) / 256;
        const int grid = static_cast<int>(want < 148 * 16 ? want : 148 * 16);
        add3_vec8_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint4 *>(a), static_cast<const uint4 *>(b),
                                                   static_cast<const uint4 *>(c), n8, static_cast<uint4 *>(out));
        MBS_CHECK_LAUNCH();
        return 0;
    }
    add3_kernel<<<static_cast<int>((n / 2 + 255) / 256), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16 *>(a), static_cast<const __nv_bfloat16 *>(b), static_cast<const __nv_bfloat16 *>(c), n,
        static_cast<__nv_bfloat16 *>(out));
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_first_conv_wgrad(const float *x, const void *dz, int N, int H, int W, int C, float *dw, float *scratch,
                                    void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(x && dz && dw && scratch && C >= 8 && C <= 256 && (C & (C - 1)) == 0, "first_conv_wgrad: C must be a power of two in [8, 256]");
    const long long M = static_cast<long long>(N) * H * W;
    const int grid = grid_rows(M, C);          // scratch: grid * C * 9 floats (<= mbs_bn_scratch_floats(2048))
    first_conv_wgrad_kernel<<<grid, 256, 0, stream>>>(x, static_cast<const __nv_bfloat16 *>(dz), N, H, W, C, scratch);
    MBS_CHECK_LAUNCH();
    reduce_partials_kernel<<<mbs::cdiv(C * 9, 32), 32 * RP_ROWS, 0, stream>>>(scratch, grid, C * 9, dw);
    MBS_CHECK_LAUNCH();
    return 0;
}
