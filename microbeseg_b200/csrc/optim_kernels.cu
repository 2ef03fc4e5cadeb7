// Fused Ranger step (RAdam + gradient centralisation + lookahead) over ALL parameter tensors in one launch.
//
// Replaces the per-tensor Python loop of the reference's default optimizer
// (/root/reference/src/training/ranger2020.py:101-208, selected by train_script.py's --optimizer ranger):
//   grad -= mean(grad, dims 1..)                      (gradient centralisation, tensors with dim > 1, :30-40,153)
//   exp_avg_sq = exp_avg_sq*beta2 + (1-beta2)*grad^2  (:158)      exp_avg = exp_avg*beta1 + (1-beta1)*grad   (:161)
//   G = N_sma > threshold ? exp_avg / (sqrt(exp_avg_sq) + eps) : exp_avg                    (:187-191)
//   G += weight_decay * p                                                                    (:193-194)
//   p -= step_size*lr * G                                                                    (:199)
//   every k steps: slow += alpha*(p - slow); p = slow                                        (:204-210)
// The scalar schedule (N_sma, step_size; :165-180) is evaluated on the host in Python floats exactly as the
// reference does and passed in.  One CTA = one "row": an output-channel slice of a dim>1 tensor (the unit the
// centralisation mean runs over) or a 1024-element chunk of a 1-D tensor.  HBM bound: 5 fp32 streams in, 4-5 out.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/mbseg.h"
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
ranger_step_kernel(const mbs_ranger_tensor *__restrict__ T, int n_tensors, float beta1, float beta2, float eps,
                   float weight_decay, float step_lr, int use_denom, int lookahead, float alpha) {
    // tensor of this row: binary search over the row prefix
    int lo = 0, hi = n_tensors - 1;
    const int row = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (T[mid].row_start <= row) lo = mid; else hi = mid - 1;
    }
    const mbs_ranger_tensor t = T[lo];
    const long long base = static_cast<long long>(row - t.row_start) * t.row_len;
    const long long rem = t.numel - base;
    const int len = static_cast<int>(rem < t.row_len ? rem : t.row_len);
    float *p = t.p + base, *g = t.g + base, *m = t.exp_avg + base, *s = t.exp_avg_sq + base;
    float *slow = t.slow + base;
    float mean = 0.0f;
    if (t.gc) {
        __shared__ float warp_part[8];
        float acc = 0.0f;
        for (int i = threadIdx.x; i < len; i += 256) acc += g[i];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        float tot = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += warp_part[w];
        mean = tot / static_cast<float>(len);
    }
    const float one_m_b1 = 1.0f - beta1, one_m_b2 = 1.0f - beta2;
    for (int i = threadIdx.x; i < len; i += 256) {
        float gi = g[i];
        if (t.gc) {
            gi = gi + (-mean);
            g[i] = gi;                      // the reference centralises p.grad in place (fp32 .float() is a view)
        }
        const float si = s[i] * beta2 + one_m_b2 * gi * gi;
        const float mi = m[i] * beta1 + one_m_b1 * gi;
        s[i] = si;
        float G = use_denom ? mi / (sqrtf(si) + eps) : mi;
        float pi = p[i];
        if (weight_decay != 0.0f) G += weight_decay * pi;
        // reference quirk (:191-194): below the N_sma threshold G_grad aliases exp_avg, so the in-place weight-decay
        // add_ also modifies the stored first moment
        m[i] = (!use_denom && weight_decay != 0.0f) ? G : mi;
        pi = pi + (-step_lr) * G;
        if (lookahead) {
            const float sl = slow[i] + alpha * (pi - slow[i]);
            slow[i] = sl;
            pi = sl;
        }
        p[i] = pi;
    }
}

// Fused Adam / AMSGrad step over ALL parameter tensors in one launch: the reference's Adam recipe
// (/root/reference/src/training/train.py:380-385: torch.optim.Adam(lr=8e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0,
// amsgrad=True)), same arithmetic as torch's single-tensor path (torch/optim/adam.py::_single_tensor_adam):
//   g += weight_decay * p;  m = m + (1-beta1)*(g - m);  v = v*beta2 + (1-beta2)*g*g;  vmax = max(vmax, v)
//   p += -step_size * m / (sqrt(vmax or v) / sqrt(bias_correction2) + eps),   step_size = lr / bias_correction1
// The table is the Ranger one (`slow` = max_exp_avg_sq); one CTA = a 1024-element chunk.  HBM bound: 5 streams in, 4 out.
__global__ void __launch_bounds__(256)
adam_step_kernel(const mbs_ranger_tensor *__restrict__ T, int n_tensors, float beta1, float beta2, float eps, float weight_decay,
                 float step_size, float bc2_sqrt, int amsgrad) {
    int lo = 0, hi = n_tensors - 1;
    const int row = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (T[mid].row_start <= row) lo = mid; else hi = mid - 1;
    }
    const mbs_ranger_tensor t = T[lo];
    const long long base = static_cast<long long>(row - t.row_start) * t.row_len;
    const long long rem = t.numel - base;
    const int len = static_cast<int>(rem < t.row_len ? rem : t.row_len);
    float *p = t.p + base, *m = t.exp_avg + base, *v = t.exp_avg_sq + base, *vm = t.slow + base;
    const float *g = t.g + base;
    const float w1 = 1.0f - beta1, w2 = 1.0f - beta2;
    for (int i = threadIdx.x; i < len; i += 256) {
        float gi = g[i];
        float pi = p[i];
        if (weight_decay != 0.0f) gi = gi + weight_decay * pi;
        const float mi = m[i] + w1 * (gi - m[i]);                  // Tensor.lerp_(grad, 1 - beta1)
        const float vi = v[i] * beta2 + w2 * gi * gi;              // mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        m[i] = mi;
        v[i] = vi;
        float d = vi;
        if (amsgrad) {
            d = fmaxf(vm[i], vi);
            vm[i] = d;
        }
        const float denom = sqrtf(d) / bc2_sqrt + eps;
        p[i] = pi + (-step_size) * (mi / denom);                   // addcdiv_(exp_avg, denom, value=-step_size)
    }
}

}  // namespace

extern "C" int mbs_adam_step(const mbs_ranger_tensor *tensors_dev, int n_tensors, int total_rows, float beta1, float beta2,
                             float eps, float weight_decay, float step_size, float bias_correction2_sqrt, int amsgrad,
                             void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(tensors_dev != nullptr && n_tensors > 0 && total_rows > 0, "adam_step: empty parameter table");
    adam_step_kernel<<<total_rows, 256, 0, stream>>>(tensors_dev, n_tensors, beta1, beta2, eps, weight_decay, step_size,
                                                     bias_correction2_sqrt, amsgrad);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_ranger_step(const mbs_ranger_tensor *tensors_dev, int n_tensors, int total_rows, float beta1, float beta2,
                               float eps, float weight_decay, float step_lr, int use_denom, int lookahead, float alpha,
                               void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(tensors_dev != nullptr && n_tensors > 0 && total_rows > 0, "ranger_step: empty parameter table");
    ranger_step_kernel<<<total_rows, 256, 0, stream>>>(tensors_dev, n_tensors, beta1, beta2, eps, weight_decay, step_lr,
                                                       use_denom, lookahead, alpha);
    MBS_CHECK_LAUNCH();
    return 0;
}
