// fp32 "check mode" of the network ("fp32/TF32-free check mode"): every convolution of the U-Net as a
// plain CUDA-core fp32 direct convolution (FMA accumulation, no tensor cores, no bf16 anywhere).  It runs the SAME layer
// plan as the product engine (layer wiring, concat order, pooling, transposed convs, folded eval BatchNorm, 1x1 heads),
// so comparing it with the fp32 oracle at full frame size checks the plan to ~1e-6, while the bf16 tensor-core kernels
// are compared with it (and with torch on identical operands, layer by layer) to isolate the precision policy.
// Diagnostic path only: ~1 s per 2048^2 frame; never used by the frame loop.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/mbseg.h"
#include "common.cuh"

namespace {

__device__ __forceinline__ float act_f32(float v, int act) {
    switch (act) {
        case MBS_ACT_RELU: return fmaxf(v, 0.0f);
        case MBS_ACT_LEAKYRELU: return v > 0.0f ? v : 0.01f * v;
        case MBS_ACT_ELU: return v > 0.0f ? v : expm1f(v);
        case MBS_ACT_MISH: {
            const float sp = v > 20.0f ? v : log1pf(expf(v));
            return v * tanhf(sp);
        }
        default: return v;
    }
}

// one thread per (output pixel, output channel); consecutive threads = consecutive output channels, so the input
// values are broadcasts and the weight reads (layout [tap][Cin][Cout]) are coalesced
__global__ void __launch_bounds__(256)
conv_ref_f32_kernel(const mbs_convref_desc d, long long total) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    const int co = static_cast<int>(idx % d.Cout);
    long long p = idx / d.Cout;
    const int Ho = d.mode == 1 ? d.H / 2 : (d.mode == 2 ? 2 * d.H : d.H);
    const int Wo = d.mode == 1 ? d.W / 2 : (d.mode == 2 ? 2 * d.W : d.W);
    const int xo = static_cast<int>(p % Wo);
    p /= Wo;
    const int yo = static_cast<int>(p % Ho);
    const int n = static_cast<int>(p / Ho);
    const int Cin = d.C0 + d.C1;
    float acc = d.bias ? d.bias[co] : 0.0f;
    auto in_at = [&](int yi, int xi, int ci) -> float {
        const size_t px = (static_cast<size_t>(n) * d.H + yi) * d.W + xi;
        return ci < d.C0 ? d.src0[px * d.C0 + ci] : d.src1[px * d.C1 + (ci - d.C0)];
    };
    if (d.mode == 2) {                                    // ConvTranspose2d(k=2, s=2): one tap per output pixel
        const int yi = yo >> 1, xi = xo >> 1, t = (yo & 1) * 2 + (xo & 1);
        const float *w = d.weight + static_cast<size_t>(t) * Cin * d.Cout + co;
        for (int ci = 0; ci < Cin; ++ci) acc = fmaf(in_at(yi, xi, ci), w[static_cast<size_t>(ci) * d.Cout], acc);
    } else if (d.mode == 4) {                             // 1x1 head
        const float *w = d.weight + co;
        for (int ci = 0; ci < Cin; ++ci) acc = fmaf(in_at(yo, xo, ci), w[static_cast<size_t>(ci) * d.Cout], acc);
    } else {                                              // 3x3, padding 1, stride 1 (mode 0) or 2 (mode 1)
        const int s = d.mode == 1 ? 2 : 1;
        for (int ky = 0; ky < 3; ++ky) {
            const int yi = yo * s + ky - 1;
            if (yi < 0 || yi >= d.H) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int xi = xo * s + kx - 1;
                if (xi < 0 || xi >= d.W) continue;
                const float *w = d.weight + static_cast<size_t>(ky * 3 + kx) * Cin * d.Cout + co;
                for (int ci = 0; ci < Cin; ++ci) acc = fmaf(in_at(yi, xi, ci), w[static_cast<size_t>(ci) * d.Cout], acc);
            }
        }
    }
    acc = act_f32(acc, d.act);
    if (d.scale) acc = fmaf(acc, d.scale[co], d.shift[co]);
    d.dst[idx] = acc;
}

}  // namespace

extern "C" int mbs_conv_ref_f32(const mbs_convref_desc *d, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(d && d->src0 && d->weight && d->dst && d->N > 0 && d->H > 0 && d->W > 0 && d->C0 > 0 && d->Cout > 0 &&
                    (d->mode == 0 || d->mode == 1 || d->mode == 2 || d->mode == 4) && (d->C1 == 0 || d->src1),
                "conv_ref_f32: bad descriptor");
    MBS_REQUIRE(d->mode != 1 || (d->H % 2 == 0 && d->W % 2 == 0), "conv_ref_f32: stride 2 needs even sizes");
    const long long Ho = d->mode == 1 ? d->H / 2 : (d->mode == 2 ? 2ll * d->H : d->H);
    const long long Wo = d->mode == 1 ? d->W / 2 : (d->mode == 2 ? 2ll * d->W : d->W);
    const long long total = static_cast<long long>(d->N) * Ho * Wo * d->Cout;
    const long long blocks = (total + 255) / 256;
    MBS_REQUIRE(blocks < (1ll << 31), "conv_ref_f32: too many outputs for one launch");
    conv_ref_f32_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(*d, total);
    MBS_CHECK_LAUNCH();
    return 0;
}
