// U-Net building blocks for sm_100a: implicit-GEMM convolutions on tcgen05 tensor cores.
//
// Replaces the cuDNN calls behind src/utils/unets.py (ConvBlock :92-173, ConvPool :176-226,
// TranspConvBlock :229-264, DUNet.forward :463-506 incl. torch.cat and the 1x1 heads).
//
// Design (one CTA = one 128-pixel x BN-channel output tile):
//   * activations are NHWC bf16; an output tile is an 8x16 pixel patch, so the A operand of tap
//     (ky,kx) is the same patch shifted by (ky-1,kx-1): one 4-D TMA box load per (tap, 64-channel
//     chunk), out-of-bounds rows/cols zero-filled by TMA = the conv's zero padding, for free;
//   * stride-2 convs use the TMA traversal stride (elementStrides=2) on the same tensor;
//   * ConvTranspose2d(2,2) is a plain GEMM [pixels,Cin]x[Cin,4*Cout] with a pixel-shuffle scatter
//     in the epilogue; torch.cat([up,skip]) is a second K source (second tensor map);
//   * smem tiles are 128B-swizzled K-major, consumed by tcgen05.mma (cta_group::1, M=128,
//     N=BN, K=16) with fp32 accumulators in TMEM; 4 epilogue warps read TMEM with tcgen05.ld and
//     apply bias -> activation -> BatchNorm(eval) affine -> bf16 (and the fused 1x1 head);
//   * warp roles: warp0 = TMA producer, warp1 = TMEM alloc + MMA issuer, warps2-5 = epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/mbseg.h"
#include "common.cuh"

namespace {

constexpr int TILE_W = 16;
constexpr int TILE_H = 8;
constexpr int BM = 128;  // TILE_W * TILE_H
constexpr int BK = 64;   // bf16 elements per K chunk = 128 bytes = one swizzle row
constexpr int A_BYTES = BM * BK * 2;
constexpr int EPI_WARPS = 8;                     // 2 per TMEM lane quadrant (column halves)
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;  // warp0 TMA, warp1 MMA, then the epilogue warps
__host__ __device__ constexpr int threads_for_groups(int ng) { return 64 + 128 * ng; }   // NG epilogue groups of 4 warps (one per TMEM lane quadrant)

struct ConvKParams {
    int mode, act;
    int tiles_x, tiles_y;     // tiles over the GEMM-M pixel grid
    int Hm, Wm;               // GEMM-M pixel grid dims (conv: output dims; convT: input dims)
    int chunks0, chunks1;     // 64-channel chunks of source 0 / source 1
    int taps;                 // 9 (conv) or 1 (convT)
    int n_tiles;              // GEMM-N tiles
    int num_tiles;            // total tiles (images x tiles_y x tiles_x x n_tiles)
    int Cout;
    const float *bias, *scale, *shift;
    __nv_bfloat16 *dst;
    int ldd, coffd, Hd, Wd;   // destination view
    const float *head_w;      // [head_n][Cout]
    float head_b[4];
    int head_n;               // number of fused 1x1 head outputs (<= 4)
    float *head_out;          // [N][head_n][H][W]
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Set when a barrier wait exceeds its cycle budget: the kernel then runs to completion with
// garbage instead of hanging the GPU; hosts/tests read it through mbs_debug_flags().
__device__ int g_mbar_timeout = 0;

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {  // ~2 s: something is wrong, never hang the device
            atomicExch(&g_mbar_timeout, 1);
            return;
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint64_t pack_f32x2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float &a, float &b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {   // sm_100 FFMA2
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3,
                                             int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Warp-uniform variants: called by ALL lanes of the MMA warp in converged code, one elected lane issues.  Inside an
// `if (lane == 0)` region the compiler wraps every UTCHMMA (its operands live in uniform registers) in an ELECT +
// BRA.U.ANY uniformisation loop -- ~10 SASS instructions and a branch per MMA, more than the 32 cycles an N = 64 MMA takes.
__device__ __forceinline__ void umma_f16_w(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "elect.sync _|q, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// descriptors as (low word, high word): the high words are loop constants and the low words (start address field) advance
// by compile-time constants, so an MMA costs two 32-bit adds, the pack and the issue
template <bool ACC>
__device__ __forceinline__ void umma_f16_w2(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "elect.sync _|q, 0xffffffff;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint32_t bar) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "elect.sync _|q, 0xffffffff;\n"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(bar)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (sm_100 UMMA).
//   [0,14)  start address >> 4      [16,30) leading byte offset >> 4 (unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups)    [46,48) version = 1
//   [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BN.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case MBS_ACT_RELU: return fmaxf(v, 0.0f);
        case MBS_ACT_LEAKYRELU: return v > 0.0f ? v : 0.01f * v;
        case MBS_ACT_ELU: return v > 0.0f ? v : expm1f(v);
        case MBS_ACT_MISH: {
            float sp = v > 20.0f ? v : log1pf(expf(v));
            return v * tanhf(sp);
        }
        default: return v;
    }
}

// Epilogue of one tile, executed by the 8 epilogue warps: TMEM -> registers -> bias / activation /
// BatchNorm affine -> bf16 -> XOR-swizzled smem transpose -> full-sector global stores (+ fused 1x1 head).
template <int BN, int TW>
__device__ __forceinline__ void epilogue_tile(const ConvKParams &p, const float *s_par, uint32_t s_epi, float *s_head,
                                              uint32_t tmem_base, uint32_t tfull, uint32_t tempty, int lt, int quad,
                                              int half, int lane, int x0, int y0, int img, int n0, bool has_head,
                                              int acc_col, bool first_sub, bool last_sub) {
    const int row = quad * 32 + lane;
    if (first_sub) mbar_wait(tfull, (lt >> 1) & 1);
    tcgen05_fence_after();
    float head_acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 1
    for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc_col + c), r);
        const int col = n0 + c;
        int q = 0, co = col;
        if (p.mode == MBS_CONVT2X2_S2) {
            q = col / p.Cout;
            co = col - q * p.Cout;
        }
        const float4 *pb = reinterpret_cast<const float4 *>(s_par + co);
        const float4 *psc = reinterpret_cast<const float4 *>(s_par + p.Cout + co);
        const float4 *psh = reinterpret_cast<const float4 *>(s_par + 2 * p.Cout + co);
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b4 = pb[j];
            v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b4.x;
            v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
            v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z;
            v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
        }
        // one uniform branch per chunk (a per-element switch costs an indirect branch each)
        switch (p.act) {
            case MBS_ACT_RELU:
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                break;
            case MBS_ACT_LEAKYRELU:
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.0f ? v[j] : 0.01f * v[j];
                break;
            case MBS_ACT_ELU:
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.0f ? v[j] : expm1f(v[j]);
                break;
            case MBS_ACT_MISH:
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], MBS_ACT_MISH);
                break;
            default: break;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 s4 = psc[j], t4 = psh[j];
            v[4 * j + 0] = fmaf(v[4 * j + 0], s4.x, t4.x);
            v[4 * j + 1] = fmaf(v[4 * j + 1], s4.y, t4.y);
            v[4 * j + 2] = fmaf(v[4 * j + 2], s4.z, t4.z);
            v[4 * j + 3] = fmaf(v[4 * j + 3], s4.w, t4.w);
        }
        if (BN == 64 && has_head) {   // heads need Cout == 64, i.e. only the BN == 64 instantiations carry this code
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                if (h < p.head_n) {
                    const float4 *phw = reinterpret_cast<const float4 *>(s_par + (3 + h) * p.Cout + co);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 w4 = phw[j];
                        head_acc[h] = fmaf(v[4 * j + 0], w4.x, head_acc[h]);
                        head_acc[h] = fmaf(v[4 * j + 1], w4.y, head_acc[h]);
                        head_acc[h] = fmaf(v[4 * j + 2], w4.z, head_acc[h]);
                        head_acc[h] = fmaf(v[4 * j + 3], w4.w, head_acc[h]);
                    }
                }
            }
        }
        if (p.dst) {
            uint32_t packed[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                packed[j] = *reinterpret_cast<uint32_t *>(&h2);
            }
            // two passes of 16 columns: stage 32 rows x 32 B (16-byte chunks XOR-swizzled, conflict free),
            // then 2 lanes write one pixel's 32 contiguous bytes (a full sector), 16 pixels per instruction
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    const uint32_t a = s_epi + lane * 32u + static_cast<uint32_t>((ch ^ ((lane >> 2) & 1)) * 16);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a),
                                 "r"(packed[8 * hh + 4 * ch]), "r"(packed[8 * hh + 4 * ch + 1]),
                                 "r"(packed[8 * hh + 4 * ch + 2]), "r"(packed[8 * hh + 4 * ch + 3])
                                 : "memory");
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int rr = i * 16 + (lane >> 1);          // row inside this warp's 32
                    const int ch = lane & 1;
                    const int trow = quad * 32 + rr;
                    const int qy = y0 + trow / TW, qx = x0 + trow % TW;
                    uint32_t v0, v1, v2, v3;
                    const uint32_t a = s_epi + rr * 32u + static_cast<uint32_t>((ch ^ ((rr >> 2) & 1)) * 16);
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                                 : "r"(a)
                                 : "memory");
                    if (qy < p.Hm && qx < p.Wm) {
                        size_t off;
                        if (p.mode == MBS_CONVT2X2_S2) {
                            const int oy = 2 * qy + (q >> 1), ox = 2 * qx + (q & 1);
                            off = ((static_cast<size_t>(img) * p.Hd + oy) * p.Wd + ox) * p.ldd + p.coffd + co;
                        } else {
                            off = ((static_cast<size_t>(img) * p.Hd + qy) * p.Wd + qx) * p.ldd + p.coffd + col;
                        }
                        *reinterpret_cast<uint4 *>(p.dst + off + hh * 16 + ch * 8) = make_uint4(v0, v1, v2, v3);
                    }
                }
                __syncwarp();
            }
        }
    }
    // all TMEM reads of this tile are complete (tcgen05.wait::ld inside tmem_ld32)
    tcgen05_fence_before();
    __syncwarp();
    if (last_sub && lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty) : "memory");
    }
    if (BN == 64 && has_head) {
        // the two column-half warps of a quadrant combine their partial dot products through smem
        float *slot = s_head + (lt & 1) * 512 + row;
        if (half == 1) {
#pragma unroll
            for (int h = 0; h < 4; ++h)
                if (h < p.head_n) slot[h * 128] = head_acc[h];
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        if (half == 0) {
            const int py = y0 + row / TW, px = x0 + row % TW;
            if (py < p.Hm && px < p.Wm) {
#pragma unroll
                for (int h = 0; h < 4; ++h)
                    if (h < p.head_n)
                        p.head_out[((static_cast<size_t>(img) * p.head_n + h) * p.Hm + py) * p.Wm + px] =
                            (head_acc[h] + slot[h * 128]) + p.head_b[h];
            }
        }
    }
}

// TMA-store epilogue.  The epilogue warps form NG groups of 4 warps (one warp per TMEM lane quadrant); a single
// warp converts ~1 column per 55 cycles (dependent TMEM load -> math -> smem store chains), so short-K layers
// (transposed convs, K = 576 full-resolution convs) need 4 groups to keep up with the tensor pipe.
// Work items are (tile, 64-column chunk) pairs dealt round-robin to the groups; a group converts its
// chunk (TMEM -> fp32 math -> bf16), writes the 128-row x 128-byte tile into its own 16 KiB staging buffer in
// the 128B-swizzled layout, and one elected thread hands it to the TMA engine (cp.async.bulk.tensor store),
// which writes full lines, clips at the tensor boundary and keeps the LSU free.  The transposed conv uses a
// 5-D view (C, dx, x, dy, y) of the 2x up-sampled destination, so its pixel-shuffle scatter is one box too.
constexpr int STAGE_TILE_BYTES = 128 * 128;
template <int BN, int TW, int NG, bool PAIR = false, int DEPTH = 1>
__device__ __forceinline__ void epilogue_tile_tma(const ConvKParams &p, const CUtensorMap *tmD, const float *s_par,
                                                  uint32_t s_stage, uint32_t tmem_base, uint32_t tfull0,
                                                  uint32_t tempty0, int lt, int group, int quad, int lane, bool leader,
                                                  int x0, int y0, int img, int n0, bool has_head, bool &pending) {
    constexpr int CHUNKS = BN / 64;
    // BN == 64: whole tiles are dealt round-robin to the groups, each group owning one TMEM accumulator
    // (a group may run ahead of the others, so it must not share an mbarrier phase sequence with them)
    // DEPTH accumulators per group: the MMAs of tile i only wait for the epilogue of tile i - DEPTH * NG.  With one
    // accumulator per group a tile costs (MMA + epilogue) / NG -- MMA and epilogue of the same buffer serialise -- and the
    // epilogue of a 64-column tile (~3600 cycles of dependent TMEM load -> math -> smem chains) dwarfs its 1152 MMA cycles.
    constexpr int NBUF = CHUNKS == 1 ? NG * DEPTH : 2;
    if (CHUNKS == 1 && (lt % NG) != group) return;
    const int buf = lt % NBUF;
    const uint32_t tfull = tfull0 + 8u * buf, tempty = tempty0 + 8u * buf;
    const int row = quad * 32 + lane;
    mbar_wait(tfull, (lt / NBUF) & 1);
    tcgen05_fence_after();
    float head_acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 1
    for (int chn = (CHUNKS == 1 ? 0 : group); chn < CHUNKS; chn += (CHUNKS == 1 ? 1 : NG)) {
        const int col0 = n0 + chn * 64;
        int q = 0, co0 = col0;
        if (p.mode == MBS_CONVT2X2_S2) {
            q = col0 / p.Cout;
            co0 = col0 - q * p.Cout;
        }
        if (p.dst && pending) {      // the previous store out of this staging buffer must have been read
            if (leader) tma_store_wait_read();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
            pending = false;
        }
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
            uint32_t r[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(buf * BN + chn * 64 + k * 32), r);
            const int co = co0 + k * 32;
            const float4 *pb = reinterpret_cast<const float4 *>(s_par + co);
            const float4 *psc = reinterpret_cast<const float4 *>(s_par + p.Cout + co);
            const float4 *psh = reinterpret_cast<const float4 *>(s_par + 2 * p.Cout + co);
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b4 = pb[j];
                v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b4.x;
                v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
                v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z;
                v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
            }
            switch (p.act) {      // one uniform branch per 32 columns
                case MBS_ACT_RELU:
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                    break;
                case MBS_ACT_LEAKYRELU:
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.0f ? v[j] : 0.01f * v[j];
                    break;
                case MBS_ACT_ELU:
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.0f ? v[j] : expm1f(v[j]);
                    break;
                case MBS_ACT_MISH:
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], MBS_ACT_MISH);
                    break;
                default: break;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 s4 = psc[j], t4 = psh[j];
                v[4 * j + 0] = fmaf(v[4 * j + 0], s4.x, t4.x);
                v[4 * j + 1] = fmaf(v[4 * j + 1], s4.y, t4.y);
                v[4 * j + 2] = fmaf(v[4 * j + 2], s4.z, t4.z);
                v[4 * j + 3] = fmaf(v[4 * j + 3], s4.w, t4.w);
            }
            if (BN == 64 && has_head) {
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    if (h < p.head_n) {
                        const float4 *phw = reinterpret_cast<const float4 *>(s_par + (3 + h) * p.Cout + co);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 w4 = phw[j];
                            head_acc[h] = fmaf(v[4 * j + 0], w4.x, head_acc[h]);
                            head_acc[h] = fmaf(v[4 * j + 1], w4.y, head_acc[h]);
                            head_acc[h] = fmaf(v[4 * j + 2], w4.z, head_acc[h]);
                            head_acc[h] = fmaf(v[4 * j + 3], w4.w, head_acc[h]);
                        }
                    }
                }
            }
            if (p.dst) {
                // row r of the staging tile = pixel r of the patch; 16-byte chunk c lives at c ^ (r & 7) (SWIZZLE_128B)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]);
                    __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
                    __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
                    const uint32_t a = s_stage + row * 128u + static_cast<uint32_t>(((k * 4 + j) ^ (row & 7)) * 16);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a),
                                 "r"(*reinterpret_cast<uint32_t *>(&h0)), "r"(*reinterpret_cast<uint32_t *>(&h1)),
                                 "r"(*reinterpret_cast<uint32_t *>(&h2)), "r"(*reinterpret_cast<uint32_t *>(&h3))
                                 : "memory");
                }
            }
        }
        if (p.dst) {
            fence_proxy_async();                         // generic-proxy smem writes -> visible to the TMA engine
            asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
            if (leader) {
                if (p.mode == MBS_CONVT2X2_S2)
                    tma_store_5d(tmD, s_stage, co0, q & 1, x0, q >> 1, img * p.Hm + y0);
                else
                    tma_store_4d(tmD, s_stage, col0, x0, y0, img);
                tma_store_commit();
            }
            pending = true;
        }
    }
    // all TMEM reads of this tile are complete (tcgen05.wait::ld inside tmem_ld32)
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) {
        if (PAIR)     // 2-CTA pair: the MMA issuer lives in the even CTA; clearing the peer bit addresses ITS barrier
            asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(tempty & 0xFEFFFFFFu) : "memory");
        else
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty) : "memory");
    }
    if (BN == 64 && has_head) {
        const int py = y0 + row / TW, px = x0 + row % TW;
        if (py < p.Hm && px < p.Wm) {
#pragma unroll
            for (int h = 0; h < 4; ++h)
                if (h < p.head_n)
                    p.head_out[((static_cast<size_t>(img) * p.head_n + h) * p.Hm + py) * p.Wm + px] = head_acc[h] + p.head_b[h];
        }
    }
}

// MT = pixel tiles (8 rows x 16 columns each, stacked in y) per work item: MT = 2 halves the weight
// traffic L2 -> smem per MMA (one B tile feeds two accumulators), which is what bounds the Cout = 128 layers.
template <int BN, int STAGES, bool TMA_EPI, int MT = 1, int NG = 2>
struct SmemPlan {
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int A_STAGE = MT * A_BYTES;
    // TMA epilogue: two 16 KiB staging tiles (one per epilogue group); direct epilogue: per warp 32 rows x 16 bf16
    // plus the head partial sums
    static constexpr int EPI_BYTES = TMA_EPI ? NG * STAGE_TILE_BYTES : EPI_WARPS * 1024 + 2 * 4 * 128 * 4;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = STAGES * A_STAGE;
    static constexpr int OFF_EPI = OFF_B + STAGES * B_BYTES;
    static constexpr int OFF_BAR = OFF_EPI + EPI_BYTES;           // full[S], empty[S], tmem_full[2], tmem_empty[2]
    static constexpr int OFF_TMEM = OFF_BAR + 8 * (2 * STAGES + 4);
    static constexpr int OFF_PAR = (OFF_TMEM + 8 + 15) / 16 * 16;  // bias/scale/shift(/head_w) floats (float4 reads)
    static constexpr int FIXED = OFF_PAR + 1024;                  // + slack for the manual 1024-B alignment
    static int dyn_bytes(int cout) { return FIXED + (3 * cout + (cout == 64 ? 4 * 64 : 0)) * 4; }   // heads only with Cout == 64
};

// Persistent, warp-specialised implicit-GEMM kernel.  Each CTA walks tiles t = blockIdx.x, +gridDim.x, ...
// The smem ring (TMA <-> MMA) runs continuously across tiles; the accumulator is double buffered in
// TMEM so that the epilogue of tile i overlaps the MMAs of tile i+1.
template <int BN, int STAGES, int MIN_CTAS, bool TMA_EPI, int MT = 1, int NG = 2>
__global__ void __launch_bounds__(threads_for_groups(NG), MIN_CTAS)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD, const ConvKParams p) {
    using Plan = SmemPlan<BN, STAGES, TMA_EPI, MT, NG>;
    static_assert(NG == 2 || (TMA_EPI && BN >= 128), "extra epilogue groups only with the TMA-store epilogue of wide tiles");
    static_assert(MT == 1 || !TMA_EPI, "the TMA-store epilogue handles one pixel tile per work item");
    static_assert(2 * MT * BN <= 512, "accumulators exceed TMEM");
    constexpr int A_STAGE = Plan::A_STAGE;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *gbase = smem_raw + (base - raw);
    const uint32_t sA = base + Plan::OFF_A;
    const uint32_t sB = base + Plan::OFF_B;
    const uint32_t sBar = base + Plan::OFF_BAR;
    auto full_bar = [&](int s) { return sBar + 8u * s; };
    auto empty_bar = [&](int s) { return sBar + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return sBar + 8u * (2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return sBar + 8u * (2 * STAGES + 2 + b); };
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(gbase + Plan::OFF_TMEM);
    float *s_par = reinterpret_cast<float *>(gbase + Plan::OFF_PAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int chunks = p.chunks0 + p.chunks1;
    const int num_k_iters = p.taps * chunks;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA0);
        if (p.chunks1 > 0) prefetch_tmap(&tmA1);
        prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), (TMA_EPI && BN == 64) ? 4 : 4 * NG);   // one arrive per participating epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_ptr)), 2 * MT * BN);
    if (warp >= 2) {
        for (int j = threadIdx.x - 64; j < p.Cout; j += threads_for_groups(NG) - 64) {
            s_par[j] = p.bias[j];
            s_par[p.Cout + j] = p.scale[j];
            s_par[2 * p.Cout + j] = p.shift[j];
            for (int h = 0; h < p.head_n; ++h) s_par[(3 + h) * p.Cout + j] = p.head_w[h * p.Cout + j];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int it_g = 0;  // ring position, continues across tiles
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.n_tiles;
                int m_tile = tile / p.n_tiles;
                const int tx = m_tile % p.tiles_x;
                const int ty = (m_tile / p.tiles_x) % p.tiles_y;
                const int img = m_tile / tiles_per_img;
                const int x0 = tx * TILE_W, y0 = ty * (TILE_H * MT), n0 = n_tile * BN;
                for (int it = 0; it < num_k_iters; ++it, ++it_g) {
                    const int s = it_g % STAGES;
                    const uint32_t ph = (it_g / STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_expect_tx(full_bar(s), A_STAGE + Plan::B_BYTES);
                    const int tap = it / chunks;
                    const int cc = it - tap * chunks;
                    int cx, cy;
                    if (p.mode == MBS_CONV3X3_S1) {
                        cx = x0 + (tap % 3) - 1;
                        cy = y0 + (tap / 3) - 1;
                    } else if (p.mode == MBS_CONV3X3_S2) {
                        cx = 2 * x0 + (tap % 3) - 1;
                        cy = 2 * y0 + (tap / 3) - 1;
                    } else if (p.mode == MBS_CONV2X2_S2) {
                        cx = 2 * x0 + (tap & 1);
                        cy = 2 * y0 + (tap >> 1);
                    } else {
                        cx = x0;
                        cy = y0;
                    }
                    if (cc < p.chunks0)
                        tma_load_4d(sA + s * A_STAGE, &tmA0, full_bar(s), cc * BK, cx, cy, img);
                    else
                        tma_load_4d(sA + s * A_STAGE, &tmA1, full_bar(s), (cc - p.chunks0) * BK, cx, cy, img);
                    tma_load_2d(sB + s * Plan::B_BYTES, &tmB, full_bar(s), it * BK, n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (single thread) =====
        {       // all lanes run the issue loop in converged code, one elected lane issues (umma_f16_w)
            constexpr uint32_t idesc = make_idesc(BM, BN);
            int it_g = 0, lt = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
                const int buf = lt & 1;
                mbar_wait(tempty_bar(buf), ((lt >> 1) & 1) ^ 1u);   // epilogue drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * MT * BN);
                for (int it = 0; it < num_k_iters; ++it, ++it_g) {
                    const int s = it_g % STAGES;
                    const uint32_t ph = (it_g / STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tcgen05_fence_after();
                    const uint64_t adesc = make_sw128_desc(sA + s * A_STAGE);
                    const uint64_t bdesc = make_sw128_desc(sB + s * Plan::B_BYTES);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // advance 32 bytes (16 bf16) inside the 128-byte swizzle row: +2 in the >>4 encoding
                            umma_f16_w(tmem_d + static_cast<uint32_t>(mt * BN), adesc + (mt * (A_BYTES >> 4)) + 2u * k,
                                     bdesc + 2u * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit_w(empty_bar(s));  // frees the smem slot when these MMAs retire
                }
                umma_commit_w(tfull_bar(buf));     // accumulators of this work item complete
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: TMEM -> registers -> bias/act/BN -> bf16 -> smem transpose -> coalesced global =====
        const int e = warp - 2;
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
        const int half = e >> 2;                   // epilogue group / which half of the BN columns this warp converts
        const bool has_head = p.head_out != nullptr;
        const uint32_t s_epi = base + Plan::OFF_EPI + static_cast<uint32_t>(e) * 1024u;
        const uint32_t s_stage = base + Plan::OFF_EPI + static_cast<uint32_t>(half) * STAGE_TILE_BYTES;
        float *s_head = reinterpret_cast<float *>(gbase + Plan::OFF_EPI + EPI_WARPS * 1024);
        const bool leader = (e & 3) == 0 && lane == 0;
        bool pending = false;
        int lt = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
            const int n_tile = tile % p.n_tiles;
            int m_tile = tile / p.n_tiles;
            const int tx = m_tile % p.tiles_x;
            const int ty = (m_tile / p.tiles_x) % p.tiles_y;
            const int img = m_tile / tiles_per_img;
            const int x0 = tx * TILE_W, y0 = ty * (TILE_H * MT), n0 = n_tile * BN;
            if (TMA_EPI) {
                epilogue_tile_tma<BN, TILE_W, NG>(p, &tmD, s_par, s_stage, tmem_base, tfull_bar(0), tempty_bar(0), lt,
                                                  half, quad, lane, leader, x0, y0, img, n0, has_head, pending);
            } else {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                    epilogue_tile<BN, TILE_W>(p, s_par, s_epi, s_head, tmem_base, tfull_bar(lt & 1), tempty_bar(lt & 1), lt,
                                              quad, half, lane, x0, y0 + mt * TILE_H, img, n0, has_head,
                                              (lt & 1) * MT * BN + mt * BN, mt == 0, mt == MT - 1);
            }
        }
        if (TMA_EPI && leader && pending) tma_store_wait_all();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 2 * MT * BN);
    }
}

// ------------------------------------------------------------------------------------------
// Full-resolution variant for Cout = 64, Cin = 64 (+64): halo tiles + smem-resident weights.
//
// At full resolution (M = H*W pixels, N = 64) the generic kernel is bound by L2 -> smem traffic:
// every tap re-reads the A patch (9x) and every tile re-reads all weights.  Here
//   * a tile is 8 wide x 16 tall, so UMMA row group g (8 rows) is image row g of the tile;
//   * per 64-channel source chunk ONE TMA box load brings the (16+2)x(8+2) pixel halo patch
//     (128 B per pixel, SWIZZLE_128B); the A operand of tap (ky,kx) is the same smem with the
//     descriptor start address advanced by (ky*10+kx)*128 B and stride-byte-offset 10*128 B
//     (the swizzle is a function of the absolute smem address, so shifted starts stay consistent
//     with what TMA wrote);
//   * all 9 x chunks weight tiles (8 KiB each) are loaded once per CTA and stay in smem.
// L2 -> smem traffic per tile drops from 9*16 KiB + 72 KiB to 22.5 KiB per source chunk.
// ------------------------------------------------------------------------------------------
constexpr int HT_W = 8, HT_H = 16;
constexpr int HALO_W = HT_W + 2, HALO_H = HT_H + 2;
constexpr int HALO_BYTES = HALO_W * HALO_H * 128;              // 23040
constexpr int HALO_SLOT = (HALO_BYTES + 1023) / 1024 * 1024;   // 23552
constexpr int W_TILE_BYTES = 64 * 128;                         // one (tap, chunk) weight tile, N = 64

template <int CHUNKS, int STAGES, int NG>
struct HaloPlan {
    static constexpr int OFF_W = 0;
    static constexpr int OFF_A = 9 * CHUNKS * W_TILE_BYTES;
    static constexpr int OFF_EPI = OFF_A + STAGES * HALO_SLOT;
    static constexpr int EPI_BYTES = NG * STAGE_TILE_BYTES;    // TMA-store staging, one tile per epilogue group
    static constexpr int NBUF = 2 * NG;                        // TMEM accumulators (two per epilogue group)
    static constexpr int OFF_BAR = OFF_EPI + EPI_BYTES;        // full[S], empty[S], tfull[NBUF], tempty[NBUF], wbar
    static constexpr int OFF_TMEM = OFF_BAR + 8 * (2 * STAGES + 2 * NBUF + 1);
    static constexpr int OFF_PAR = (OFF_TMEM + 8 + 15) / 16 * 16;
    static constexpr int DYN_BYTES = OFF_PAR + 7 * 64 * 4 + 1024;
};

__device__ __forceinline__ uint64_t make_sw128_desc_sbo(uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// FIRST: the network's first layer (normalise + pad + Conv2d(1,64,3) + act + BN affine, see first_conv_kernel) is computed by
// four extra producer warps straight into the swizzled halo stages -- the A operand never exists in global memory (saves
// one 64-channel full-resolution write + read per frame).  Same FFMA2 chains as first_conv_kernel, so the bf16 values in
// shared memory are bit-identical to what TMA would have loaded from first_conv_kernel's output.
struct FirstParams {
    const void *img;
    int H, W, pad_y, pad_x;
    float lo, hi;
    const float *lohi_dev;
    const float *weight, *bias, *scale, *shift;
    int act;
};
constexpr int FIRST_IN_W = HALO_W + 2, FIRST_IN_H = HALO_H + 2;     // input patch of one halo patch (12 x 20)
constexpr int FIRST_THREADS = 256;

template <int CHUNKS, int STAGES, int NG, bool FIRST = false, typename T = uint8_t, bool RELU1 = false>
__global__ void __launch_bounds__(threads_for_groups(NG) + (FIRST ? FIRST_THREADS : 0), 1)
conv_halo64_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                   const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
                   const ConvKParams p, const FirstParams f) {
    using Plan = HaloPlan<CHUNKS, STAGES, NG>;
    constexpr int BN = 64;
    static_assert(NG == 2 || NG == 4, "TMEM allocations are powers of two (2 * NG * 64 columns)");
    static_assert(!FIRST || CHUNKS == 1, "the fused first layer is the only source");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *gbase = smem_raw + (base - raw);
    const uint32_t sW = base + Plan::OFF_W;
    const uint32_t sA = base + Plan::OFF_A;
    const uint32_t sBar = base + Plan::OFF_BAR;
    auto full_bar = [&](int s) { return sBar + 8u * s; };
    auto empty_bar = [&](int s) { return sBar + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return sBar + 8u * (2 * STAGES + b); };
    constexpr int NBUF = Plan::NBUF;
    auto tempty_bar = [&](int b) { return sBar + 8u * (2 * STAGES + NBUF + b); };
    const uint32_t w_bar = sBar + 8u * (2 * STAGES + 2 * NBUF);
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(gbase + Plan::OFF_TMEM);
    float *s_par = reinterpret_cast<float *>(gbase + Plan::OFF_PAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA0);
        if (CHUNKS > 1) prefetch_tmap(&tmA1);
        prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 4);     // BN == 64: tiles are dealt round-robin to the epilogue groups
        }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_ptr)), 2 * NG * BN);
    if (warp >= 2) {
        for (int j = threadIdx.x - 64; j < 64; j += threads_for_groups(NG) - 64) {
            s_par[j] = p.bias[j];
            s_par[64 + j] = p.scale[j];
            s_par[128 + j] = p.shift[j];
            for (int h = 0; h < p.head_n; ++h) s_par[(3 + h) * 64 + j] = p.head_w[h * 64 + j];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: resident weights once, then one halo patch per (tile, source chunk) =====
            mbar_expect_tx(w_bar, 9 * CHUNKS * W_TILE_BYTES);
            for (int t = 0; t < 9 * CHUNKS; ++t) tma_load_2d(sW + t * W_TILE_BYTES, &tmB, w_bar, t * BK, 0);
            int it_g = 0;
            for (int tile = blockIdx.x; !FIRST && tile < p.num_tiles; tile += gridDim.x) {
                const int tx = tile % p.tiles_x;
                const int ty = (tile / p.tiles_x) % p.tiles_y;
                const int img = tile / tiles_per_img;
                const int x0 = tx * HT_W, y0 = ty * HT_H;
#pragma unroll
                for (int cc = 0; cc < CHUNKS; ++cc, ++it_g) {
                    const int s = it_g % STAGES;
                    const uint32_t ph = (it_g / STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_expect_tx(full_bar(s), HALO_BYTES);
                    tma_load_4d(sA + s * HALO_SLOT, cc == 0 ? &tmA0 : &tmA1, full_bar(s), 0, x0 - 1, y0 - 1, img);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop in converged code, one elected lane issues (see umma_f16_w) =====
        {
            constexpr uint32_t idesc = make_idesc(BM, BN);
            mbar_wait(w_bar, 0);
            int it_g = 0, lt = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
                const int buf = lt % NBUF;
                mbar_wait(tempty_bar(buf), ((lt / NBUF) & 1) ^ 1u);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * BN);
#pragma unroll
                for (int cc = 0; cc < CHUNKS; ++cc, ++it_g) {
                    const int s = it_g % STAGES;
                    const uint32_t ph = (it_g / STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tcgen05_fence_after();
                    const uint32_t a_slot = sA + s * HALO_SLOT;
                    const uint64_t adesc0 = make_sw128_desc_sbo(a_slot, HALO_W * 128);
                    const uint64_t bdesc0 = make_sw128_desc(sW + cc * W_TILE_BYTES);
                    const uint32_t alo = static_cast<uint32_t>(adesc0), ahi = static_cast<uint32_t>(adesc0 >> 32);
                    const uint32_t blo = static_cast<uint32_t>(bdesc0), bhi = static_cast<uint32_t>(bdesc0 >> 32);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // start-address fields (>> 4): tap shift inside the halo patch, weight tile of the tap, 32-byte K step
                            const uint32_t da = static_cast<uint32_t>((((tap / 3) * HALO_W + (tap % 3)) * 128 + k * 32) >> 4);
                            const uint32_t db = static_cast<uint32_t>((tap * CHUNKS * W_TILE_BYTES + k * 32) >> 4);
                            if (cc == 0 && tap == 0 && k == 0)
                                umma_f16_w2<false>(tmem_d, alo + da, ahi, blo + db, bhi, idesc);
                            else
                                umma_f16_w2<true>(tmem_d, alo + da, ahi, blo + db, bhi, idesc);
                        }
                    }
                    umma_commit_w(empty_bar(s));
                }
                umma_commit_w(tfull_bar(buf));
            }
        }
        __syncwarp();
    } else if (FIRST && warp >= 2 + 4 * NG) {
        // ===== first-layer producer warps: compute the halo patch of every tile into its swizzled stage =====
        // thread = (4-channel group g, position lane pl); two halo positions per iteration (independent FFMA2 chains)
        const int tp = threadIdx.x - threads_for_groups(NG);
        const int g = tp & 15;
        const int pl = tp >> 4;
        constexpr int LANES = FIRST_THREADS / 16;
        float *s_in = s_par + 7 * 64;         // [FIRST_IN_H][FIRST_IN_W] normalised input patch
        const T *img_base = static_cast<const T *>(f.img);
        float lo = f.lo, hi = f.hi;
        if (f.lohi_dev) {
            lo = f.lohi_dev[0];
            hi = f.lohi_dev[1];
        }
        const float range = hi - lo;
        const int Hp = f.H + f.pad_y, Wp = f.W + f.pad_x;
        // channels 0,1 of the group as FFMA2 chains (fma.rn.f32x2: 4 cycles of the fmaheavy pipe per warp instruction),
        // channels 2,3 as scalar FFMA chains (either FMA pipe): measured, an all-FFMA2 producer is bound by fmaheavy alone
        uint64_t w2[9], b2, sc2, sh2;
        float w1[2][9], b1[2], sc1[2], sh1[2];
        {
            const int ca = g * 4;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                w2[t] = pack_f32x2(f.weight[ca * 9 + t], f.weight[(ca + 1) * 9 + t]);
                w1[0][t] = f.weight[(ca + 2) * 9 + t];
                w1[1][t] = f.weight[(ca + 3) * 9 + t];
            }
            b2 = pack_f32x2(f.bias[ca], f.bias[ca + 1]);
            sc2 = pack_f32x2(f.scale[ca], f.scale[ca + 1]);
            sh2 = pack_f32x2(f.shift[ca], f.shift[ca + 1]);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                b1[u] = f.bias[ca + 2 + u];
                sc1[u] = f.scale[ca + 2 + u];
                sh1[u] = f.shift[ca + 2 + u];
            }
        }
        constexpr int IN_N = FIRST_IN_H * FIRST_IN_W;                       // 240 <= FIRST_THREADS: one value per thread
        static_assert(IN_N <= FIRST_THREADS, "one input value per producer thread");
        const int in_r = tp / FIRST_IN_W, in_c = tp - in_r * FIRST_IN_W;
        // raw input value of a tile's patch: the load is issued one tile ahead, the value is converted / normalised only
        // after the FMA work of the current tile (finish), so the global latency hides behind it
        int kind = 0;      // 0: conv zero padding, 1: pad value (frame min), 2: pixel
        auto load_raw = [&](int tx, int ty, int img) -> T {
            const int yy = ty * HT_H - 2 + in_r, xx = tx * HT_W - 2 + in_c;
            kind = 0;
            T raw = T(0);
            if (tp < IN_N && yy >= 0 && yy < Hp && xx >= 0 && xx < Wp) {
                kind = 1;
                if (yy >= f.pad_y && xx >= f.pad_x) {
                    kind = 2;
                    raw = img_base[(static_cast<size_t>(img) * f.H + (yy - f.pad_y)) * f.W + (xx - f.pad_x)];
                }
            }
            return raw;
        };
        auto finish = [&](T rawv) -> float {
            if (kind == 0) return 0.0f;
            const float raw = kind == 2 ? static_cast<float>(rawv) : lo;
            return hi < lo ? raw : __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(raw, lo)), range), 1.0f);
        };
        // one halo position, 4 channels, first_conv_kernel's summation order per channel
        auto item = [&](const float (&in)[9], uint32_t (&out)[2]) {
            uint64_t a = b2;
            float c0 = b1[0], c1 = b1[1];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                a = fma_f32x2(pack_f32x2(in[t], in[t]), w2[t], a);
                c0 = __fmaf_rn(in[t], w1[0][t], c0);
                c1 = __fmaf_rn(in[t], w1[1][t], c1);
            }
            float o0, o1;
            unpack_f32x2(a, o0, o1);
            if (RELU1) {
                o0 = fmaxf(o0, 0.0f);
                o1 = fmaxf(o1, 0.0f);
                c0 = fmaxf(c0, 0.0f);
                c1 = fmaxf(c1, 0.0f);
            } else if (f.act != MBS_ACT_NONE) {
                o0 = apply_act(o0, f.act);
                o1 = apply_act(o1, f.act);
                c0 = apply_act(c0, f.act);
                c1 = apply_act(c1, f.act);
            }
            unpack_f32x2(fma_f32x2(pack_f32x2(o0, o1), sc2, sh2), o0, o1);
            c0 = __fmaf_rn(c0, sc1[0], sh1[0]);
            c1 = __fmaf_rn(c1, sc1[1], sh1[1]);
            __nv_bfloat162 h0 = __floats2bfloat162_rn(o0, o1), h1 = __floats2bfloat162_rn(c0, c1);
            out[0] = *reinterpret_cast<uint32_t *>(&h0);
            out[1] = *reinterpret_cast<uint32_t *>(&h1);
        };
        // tile coordinates are carried from the prefetch to the next iteration (one set of divisions per tile)
        int tile = blockIdx.x;
        int tx = 0, ty = 0, img = 0;
        if (tile < p.num_tiles) {
            tx = tile % p.tiles_x;
            ty = (tile / p.tiles_x) % p.tiles_y;
            img = tile / tiles_per_img;
        }
        T rawv = tile < p.num_tiles ? load_raw(tx, ty, img) : T(0);
        for (int it_g = 0; tile < p.num_tiles; ++it_g) {
            const int x0 = tx * HT_W - 1, y0 = ty * HT_H - 1;       // image position of halo pixel (0, 0)
            if (tp < IN_N) s_in[tp] = finish(rawv);
            named_bar_sync(8, FIRST_THREADS);
            tile += gridDim.x;
            if (tile < p.num_tiles) {
                tx = tile % p.tiles_x;
                ty = (tile / p.tiles_x) % p.tiles_y;
                img = tile / tiles_per_img;
                rawv = load_raw(tx, ty, img);
            }
            const int s = it_g % STAGES;
            mbar_wait(empty_bar(s), ((it_g / STAGES) & 1) ^ 1u);
            const uint32_t slot = sA + s * HALO_SLOT;
#pragma unroll 1
            for (int pos0 = pl; pos0 < HALO_W * HALO_H; pos0 += 2 * LANES) {
                float in[2][9];
                bool live[2];
                int pos[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    pos[q] = pos0 + q * LANES;
                    const int pc = pos[q] < HALO_W * HALO_H ? pos[q] : pl;       // clamp: the value is not stored
                    const int hy = pc / HALO_W, hx = pc - hy * HALO_W;
                    live[q] = x0 + hx >= 0 && x0 + hx < Wp && y0 + hy >= 0 && y0 + hy < Hp;
                    const float *ip = s_in + hy * FIRST_IN_W + hx;
#pragma unroll
                    for (int t = 0; t < 9; ++t) in[q][t] = ip[(t / 3) * FIRST_IN_W + (t % 3)];
                }
                uint32_t out[2][2];
                item(in[0], out[0]);
                item(in[1], out[1]);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (pos[q] >= HALO_W * HALO_H) continue;
                    // outside the layer's output = the conv's zero padding (what TMA zero fill gave);
                    // SWIZZLE_128B: 16-byte chunk index XOR (128-byte row index mod 8); the slot is 1024-byte aligned
                    st_shared_v2(slot + pos[q] * 128 + (((g >> 1) ^ (pos[q] & 7)) << 4) + ((g & 1) << 3),
                                 live[q] ? out[q][0] : 0u, live[q] ? out[q][1] : 0u);
                }
            }
            fence_proxy_async();               // generic-proxy stores -> visible to the tensor core's async-proxy reads
            named_bar_sync(8, FIRST_THREADS);  // also: everyone is done reading s_in
            if (tp == 0) mbar_arrive(full_bar(s));
        }
    } else {
        const int e = warp - 2;
        const int quad = warp & 3;
        const int group = e >> 2;
        const uint32_t s_stage = base + Plan::OFF_EPI + static_cast<uint32_t>(group) * STAGE_TILE_BYTES;
        const bool has_head = p.head_out != nullptr;
        const bool leader = (e & 3) == 0 && lane == 0;
        bool pending = false;
        int lt = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
            const int tx = tile % p.tiles_x;
            const int ty = (tile / p.tiles_x) % p.tiles_y;
            const int img = tile / tiles_per_img;
            epilogue_tile_tma<BN, HT_W, NG, false, 2>(p, &tmD, s_par, s_stage, tmem_base, tfull_bar(0), tempty_bar(0), lt, group,
                                                      quad, lane, leader, tx * HT_W, ty * HT_H, img, 0, has_head, pending);
        }
        if (leader && pending) tma_store_wait_all();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 2 * NG * BN);
    }
}

// ------------------------------------------------------------------------------------------
// conv_halo64_kernel as a 2-CTA PAIR (cta_group::2, cluster of two CTAs on neighbouring SMs).
//
// M = 128, N = 64 MMAs read 4 KiB (A) + 2 KiB (B) of shared memory per 32 cycles = 192 B/clk against 128 B/clk of
// shared-memory bandwidth: the full-resolution Cout = 64 layers sit at 46-51 % of the tensor pipe.  With cta_group::2 one
// instruction drives both SMs' tensor cores on M = 256 (each CTA's own 128-pixel tile) and each CTA supplies HALF of the
// weight tile (32 of the 64 output channels): 4 + 1 KiB per 32 cycles = 160 B/clk.  The even CTA issues every MMA; both
// CTAs load their own halo patches (the odd CTA's TMA signals the even CTA's mbarrier: .cta_group::2 loads with the peer bit
// of the barrier address cleared), stage / accumulator hand-backs are multicast commits to both CTAs, the odd CTA's
// epilogue warps arrive on the even CTA's "accumulator drained" barrier.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
__device__ __forceinline__ void umma_f16_2sm_w(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;       // all lanes call, one elected lane issues (see umma_f16_w)
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "elect.sync _|q, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_w(uint32_t bar) {
    const unsigned short mask = 3;
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "elect.sync _|q, 0xffffffff;\n"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
        "}\n" ::"r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {      // arrives on the barrier at this offset in BOTH CTAs
    const unsigned short mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

constexpr int W_HALF_BYTES = 32 * 128;                         // one (tap, chunk) weight tile, this CTA's 32 output channels

template <int CHUNKS, int STAGES, int NG>
struct HaloPairPlan {
    static constexpr int OFF_W = 0;
    static constexpr int OFF_A = 9 * CHUNKS * W_HALF_BYTES;
    static constexpr int OFF_EPI = OFF_A + STAGES * HALO_SLOT;
    static constexpr int EPI_BYTES = NG * STAGE_TILE_BYTES;
    static constexpr int OFF_BAR = OFF_EPI + EPI_BYTES;        // full[S], empty[S], tfull[NG], tempty[NG], wbar
    static constexpr int OFF_TMEM = OFF_BAR + 8 * (2 * STAGES + 2 * NG + 1);
    static constexpr int OFF_PAR = (OFF_TMEM + 8 + 15) / 16 * 16;
    static constexpr int DYN_BYTES = OFF_PAR + 7 * 64 * 4 + 1024;
};

template <int CHUNKS, int STAGES, int NG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(threads_for_groups(NG), 1)
conv_halo64_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                        const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD, const ConvKParams p) {
    using Plan = HaloPairPlan<CHUNKS, STAGES, NG>;
    constexpr int BN = 64;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *gbase = smem_raw + (base - raw);
    const uint32_t sW = base + Plan::OFF_W;
    const uint32_t sA = base + Plan::OFF_A;
    const uint32_t sBar = base + Plan::OFF_BAR;
    auto full_bar = [&](int s) { return sBar + 8u * s; };
    auto empty_bar = [&](int s) { return sBar + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return sBar + 8u * (2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return sBar + 8u * (2 * STAGES + NG + b); };
    const uint32_t w_bar = sBar + 8u * (2 * STAGES + 2 * NG);
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(gbase + Plan::OFF_TMEM);
    float *s_par = reinterpret_cast<float *>(gbase + Plan::OFF_PAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const bool lead = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int num_pairs = (p.num_tiles + 1) >> 1;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA0);
        if (CHUNKS > 1) prefetch_tmap(&tmA1);
        prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);       // used in the even CTA only: its producer's arrive + both CTAs' TMA bytes
            mbar_init(empty_bar(s), 1);      // multicast commit from the even CTA's MMA thread
        }
        for (int b = 0; b < NG; ++b) {
            mbar_init(tfull_bar(b), 1);      // multicast commit
            mbar_init(tempty_bar(b), 8);     // used in the even CTA only: 4 epilogue warps of each CTA
        }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(smem_u32(const_cast<uint32_t *>(tmem_ptr)), NG * BN);
    if (warp >= 2) {
        for (int j = threadIdx.x - 64; j < 64; j += threads_for_groups(NG) - 64) {
            s_par[j] = p.bias[j];
            s_par[64 + j] = p.scale[j];
            s_par[128 + j] = p.shift[j];
            for (int h = 0; h < p.head_n; ++h) s_par[(3 + h) * 64 + j] = p.head_w[h * 64 + j];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                      // both CTAs' barriers are initialised before any remote arrive / multicast commit
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    auto my_tile = [&](int pair) {
        const int t = 2 * pair + rank;
        return t < p.num_tiles ? t : p.num_tiles - 1;      // odd tile count: the odd CTA repeats the last tile (identical stores)
    };

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer (both CTAs): this CTA's half of the weights once, then one halo patch per (pair, chunk) =====
            if (lead) mbar_expect_tx(w_bar, 2 * 9 * CHUNKS * W_HALF_BYTES);
            for (int t = 0; t < 9 * CHUNKS; ++t) tma_load_2d_2sm(sW + t * W_HALF_BYTES, &tmB, w_bar, t * BK, 32 * rank);
            int it_g = 0;
            for (int pair = cluster_id; pair < num_pairs; pair += num_clusters) {
                const int tile = my_tile(pair);
                const int tx = tile % p.tiles_x;
                const int ty = (tile / p.tiles_x) % p.tiles_y;
                const int img = tile / tiles_per_img;
                const int x0 = tx * HT_W, y0 = ty * HT_H;
#pragma unroll
                for (int cc = 0; cc < CHUNKS; ++cc, ++it_g) {
                    const int s = it_g % STAGES;
                    const uint32_t ph = (it_g / STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    if (lead) mbar_expect_tx(full_bar(s), 2 * HALO_BYTES);
                    tma_load_4d_2sm(sA + s * HALO_SLOT, cc == 0 ? &tmA0 : &tmA1, full_bar(s), 0, x0 - 1, y0 - 1, img);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the EVEN CTA drives both tensor cores =====
        if (lead) {
            constexpr uint32_t idesc = make_idesc(256, BN);
            mbar_wait(w_bar, 0);
            int it_g = 0, lt = 0;
            for (int pair = cluster_id; pair < num_pairs; pair += num_clusters, ++lt) {
                const int buf = lt % NG;
                mbar_wait(tempty_bar(buf), ((lt / NG) & 1) ^ 1u);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * BN);
#pragma unroll
                for (int cc = 0; cc < CHUNKS; ++cc, ++it_g) {
                    const int s = it_g % STAGES;
                    const uint32_t ph = (it_g / STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tcgen05_fence_after();
                    const uint32_t a_slot = sA + s * HALO_SLOT;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t a_tap = a_slot + static_cast<uint32_t>(((tap / 3) * HALO_W + (tap % 3)) * 128);
                        const uint64_t adesc = make_sw128_desc_sbo(a_tap, HALO_W * 128);
                        const uint64_t bdesc = make_sw128_desc(sW + (tap * CHUNKS + cc) * W_HALF_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16_2sm_w(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (cc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit_2sm_w(empty_bar(s));
                }
                umma_commit_2sm_w(tfull_bar(buf));
            }
        }
        __syncwarp();
    } else {
        const int e = warp - 2;
        const int quad = warp & 3;
        const int group = e >> 2;
        const uint32_t s_stage = base + Plan::OFF_EPI + static_cast<uint32_t>(group) * STAGE_TILE_BYTES;
        const bool has_head = p.head_out != nullptr;
        const bool leader = (e & 3) == 0 && lane == 0;
        bool pending = false;
        int lt = 0;
        for (int pair = cluster_id; pair < num_pairs; pair += num_clusters, ++lt) {
            const int tile = my_tile(pair);
            const int tx = tile % p.tiles_x;
            const int ty = (tile / p.tiles_x) % p.tiles_y;
            const int img = tile / tiles_per_img;
            epilogue_tile_tma<BN, HT_W, NG, true>(p, &tmD, s_par, s_stage, tmem_base, tfull_bar(0), tempty_bar(0), lt, group, quad, lane,
                                                  leader, tx * HT_W, ty * HT_H, img, 0, has_head, pending);
        }
        if (leader && pending) tma_store_wait_all();
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                      // the peer's tensor core reads this CTA's weight half until its last MMA retires
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_2sm(tmem_base, NG * BN);
    }
}

// ------------------------------------------------------------------------------------------
// Halo tiles with STREAMED weights for the wide stride-1 layers (Cout = 128: two stacked pixel tiles, Cout >= 256).
//
// ncu on the generic kernel (512 -> 512 @256^2): tensor pipe 74 %, L2 -> SM 10.5 TB/s = 44 B/clk/SM, i.e. AT the chip-wide
// L2 delivery cap (~6300 B/clk): every tap re-reads its 16 KiB A patch from L2 (9 x per 64-channel chunk) next to the
// 32 KiB weight tile.  Here ONE TMA box brings the (16*MT + 2) x 10 pixel halo patch of a chunk (23 / 43.5 KiB) and the
// nine taps are nine UMMA descriptors into it (same trick as conv_halo64_kernel); the weight tiles stream through their
// own ring, one per (tap, chunk).  L2 -> SM traffic per chunk: 9 x 16 + 9 x 32 = 432 KiB -> 23 + 288 = 311 KiB (BN = 256).
// K order: chunk-major, tap inside (the generic kernel: tap-major) -- fp32 accumulation order differs, nothing else.
// ------------------------------------------------------------------------------------------
template <int BN, int MT, int SA, int SB>
struct HsPlan {
    static constexpr int HALO_ROWS = MT * HT_H + 2;
    static constexpr int A_BYTES_ = HALO_W * HALO_ROWS * 128;
    static constexpr int A_SLOT = (A_BYTES_ + 1023) / 1024 * 1024;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = SA * A_SLOT;
    static constexpr int OFF_EPI = OFF_B + SB * B_BYTES;
    static constexpr int EPI_BYTES = EPI_WARPS * 2048 + 2 * 4 * 128 * 4;     // 2 KiB of staging per epilogue warp (transposing epilogue)
    static constexpr int OFF_BAR = OFF_EPI + EPI_BYTES;            // fullA[SA], emptyA[SA], fullB[SB], emptyB[SB], tfull[2], tempty[2]
    static constexpr int OFF_TMEM = OFF_BAR + 8 * (2 * SA + 2 * SB + 4);
    static constexpr int OFF_PAR = (OFF_TMEM + 8 + 15) / 16 * 16;
    static int dyn_bytes(int cout) { return OFF_PAR + 1024 + 3 * cout * 4; }
};

template <int BN, int MT, int SA, int SB>
__global__ void __launch_bounds__(threads_for_groups(2), 1)
conv_hstream_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmB, const ConvKParams p) {
    using Plan = HsPlan<BN, MT, SA, SB>;
    static_assert(2 * MT * BN <= 512, "accumulators exceed TMEM");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *gbase = smem_raw + (base - raw);
    const uint32_t sA = base + Plan::OFF_A;
    const uint32_t sB = base + Plan::OFF_B;
    const uint32_t sBar = base + Plan::OFF_BAR;
    auto fullA = [&](int s) { return sBar + 8u * s; };
    auto emptyA = [&](int s) { return sBar + 8u * (SA + s); };
    auto fullB = [&](int s) { return sBar + 8u * (2 * SA + s); };
    auto emptyB = [&](int s) { return sBar + 8u * (2 * SA + SB + s); };
    auto tfull_bar = [&](int b) { return sBar + 8u * (2 * SA + 2 * SB + b); };
    auto tempty_bar = [&](int b) { return sBar + 8u * (2 * SA + 2 * SB + 2 + b); };
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(gbase + Plan::OFF_TMEM);
    float *s_par = reinterpret_cast<float *>(gbase + Plan::OFF_PAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int chunks = p.chunks0 + p.chunks1;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA0);
        if (p.chunks1 > 0) prefetch_tmap(&tmA1);
        prefetch_tmap(&tmB);
        for (int s = 0; s < SA; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_ptr)), 2 * MT * BN);
    if (warp >= 2) {
        for (int j = threadIdx.x - 64; j < p.Cout; j += threads_for_groups(2) - 64) {
            s_par[j] = p.bias[j];
            s_par[p.Cout + j] = p.scale[j];
            s_par[2 * p.Cout + j] = p.shift[j];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: one halo patch per (tile, chunk), one weight tile per (tile, chunk, tap) =====
            int ia = 0, ib = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.n_tiles;
                const int m_tile = tile / p.n_tiles;
                const int tx = m_tile % p.tiles_x;
                const int ty = (m_tile / p.tiles_x) % p.tiles_y;
                const int img = m_tile / tiles_per_img;
                const int x0 = tx * HT_W, y0 = ty * (HT_H * MT), n0 = n_tile * BN;
                for (int cc = 0; cc < chunks; ++cc, ++ia) {
                    const int sa = ia % SA;
                    mbar_wait(emptyA(sa), ((ia / SA) & 1) ^ 1u);
                    mbar_expect_tx(fullA(sa), Plan::A_BYTES_);
                    if (cc < p.chunks0)
                        tma_load_4d(sA + sa * Plan::A_SLOT, &tmA0, fullA(sa), cc * BK, x0 - 1, y0 - 1, img);
                    else
                        tma_load_4d(sA + sa * Plan::A_SLOT, &tmA1, fullA(sa), (cc - p.chunks0) * BK, x0 - 1, y0 - 1, img);
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap, ++ib) {
                        const int sb = ib % SB;
                        mbar_wait(emptyB(sb), ((ib / SB) & 1) ^ 1u);
                        mbar_expect_tx(fullB(sb), Plan::B_BYTES);
                        tma_load_2d(sB + sb * Plan::B_BYTES, &tmB, fullB(sb), (tap * chunks + cc) * BK, n0);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (single thread) =====
        {       // all lanes run the issue loop in converged code, one elected lane issues (umma_f16_w)
            constexpr uint32_t idesc = make_idesc(BM, BN);
            int ia = 0, ib = 0, lt = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
                const int buf = lt & 1;
                mbar_wait(tempty_bar(buf), ((lt >> 1) & 1) ^ 1u);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * MT * BN);
                for (int cc = 0; cc < chunks; ++cc, ++ia) {
                    const int sa = ia % SA;
                    mbar_wait(fullA(sa), (ia / SA) & 1);
                    tcgen05_fence_after();
                    const uint32_t a_slot = sA + sa * Plan::A_SLOT;
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap, ++ib) {
                        const int sb = ib % SB;
                        mbar_wait(fullB(sb), (ib / SB) & 1);
                        tcgen05_fence_after();
                        const uint32_t a_tap = a_slot + static_cast<uint32_t>(((tap / 3) * HALO_W + (tap % 3)) * 128);
                        const uint64_t bdesc = make_sw128_desc(sB + sb * Plan::B_BYTES);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            const uint64_t adesc = make_sw128_desc_sbo(a_tap + static_cast<uint32_t>(mt * HT_H * HALO_W * 128), HALO_W * 128);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma_f16_w(tmem_d + static_cast<uint32_t>(mt * BN), adesc + 2u * k, bdesc + 2u * k, idesc,
                                         (cc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit_w(emptyB(sb));
                    }
                    umma_commit_w(emptyA(sa));
                }
                umma_commit_w(tfull_bar(buf));
            }
        }
        __syncwarp();
    } else {
        const int e = warp - 2;
        const int quad = warp & 3;
        const int half = e >> 2;
        const uint32_t s_epi = base + Plan::OFF_EPI + static_cast<uint32_t>(e) * 1024u;
        float *s_head = reinterpret_cast<float *>(gbase + Plan::OFF_EPI + EPI_WARPS * 1024);
        int lt = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
            const int n_tile = tile % p.n_tiles;
            const int m_tile = tile / p.n_tiles;
            const int tx = m_tile % p.tiles_x;
            const int ty = (m_tile / p.tiles_x) % p.tiles_y;
            const int img = m_tile / tiles_per_img;
            const int x0 = tx * HT_W, y0 = ty * (HT_H * MT), n0 = n_tile * BN;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
                epilogue_tile<BN, HT_W>(p, s_par, s_epi, s_head, tmem_base, tfull_bar(lt & 1), tempty_bar(lt & 1), lt, quad, half, lane,
                                        x0, y0 + mt * HT_H, img, n0, false, (lt & 1) * MT * BN + mt * BN, mt == 0, mt == MT - 1);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 2 * MT * BN);
    }
}

// ------------------------------------------------------------------------------------------
// Cout = 128: the same halo / streamed-weight scheme with the operand roles SWAPPED.
//
// With pixels as the M operand a 128-channel layer issues M = 128, N = 128 MMAs: 4 KiB (A) + 4 KiB (B) of shared-memory
// operand reads per 64 cycles = 128 B/clk, the shared-memory bandwidth itself (ncu: tensor pipe 61-66 %, whatever the
// L2 traffic).  Here D^T[cout][pixel] = W[cout][k] * X[pixel][k]^T: the 128 output channels are the M rows (weight tile =
// A operand, K-major as TMA delivers it), 256 pixels (an 8 x 32 tile of the halo patch, one descriptor with the 10-pixel
// row pitch as stride byte offset) are the N columns: 12 KiB per 128 cycles = 96 B/clk, as for the Cout >= 256 layers.
// The accumulator comes out channel-major (TMEM lane = channel): per-channel bias / affine are per-thread scalars, and
// the epilogue transposes 32 x 32 blocks through shared memory so that global stores are 16-byte NHWC vectors.
// ------------------------------------------------------------------------------------------
template <int TW>
__device__ __forceinline__ void epilogue_tile_T(const ConvKParams &p, const float *s_par, uint32_t s_epi, uint32_t tmem_base, uint32_t tfull,
                                                uint32_t tempty, int lt, int quad, int half, int lane, int x0, int y0, int img, int acc_col) {
    mbar_wait(tfull, (lt >> 1) & 1);
    tcgen05_fence_after();
    const int ch = quad * 32 + lane;
    const float bias = s_par[ch], sc = s_par[p.Cout + ch], sh = s_par[2 * p.Cout + ch];
#pragma unroll 1
    for (int c = half * 128; c < (half + 1) * 128; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc_col + c), r);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bias;
        switch (p.act) {      // one uniform branch per 32 pixels
            case MBS_ACT_RELU:
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                break;
            case MBS_ACT_LEAKYRELU:
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.0f ? v[j] : 0.01f * v[j];
                break;
            case MBS_ACT_ELU:
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.0f ? v[j] : expm1f(v[j]);
                break;
            case MBS_ACT_MISH:
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], MBS_ACT_MISH);
                break;
            default: break;
        }
        // stage [pixel][channel] (64 B per pixel for this warp's 32 channels); lanes = consecutive channels: conflict free
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const __nv_bfloat16 h = __float2bfloat16_rn(fmaf(v[j], sc, sh));
            asm volatile("st.shared.b16 [%0], %1;" ::"r"(s_epi + static_cast<uint32_t>(j * 64 + lane * 2)), "h"(*reinterpret_cast<const unsigned short *>(&h)) : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int unit = i * 32 + lane;
            const int px = unit >> 2, c8 = unit & 3;
            uint32_t v0, v1, v2, v3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(s_epi + static_cast<uint32_t>(px * 64 + c8 * 16)) : "memory");
            const int col = c + px;
            const int qy = y0 + col / TW, qx = x0 + col % TW;
            if (qy < p.Hm && qx < p.Wm) {
                const size_t off = ((static_cast<size_t>(img) * p.Hd + qy) * p.Wd + qx) * p.ldd + p.coffd + quad * 32 + c8 * 8;
                *reinterpret_cast<uint4 *>(p.dst + off) = make_uint4(v0, v1, v2, v3);
            }
        }
        __syncwarp();
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty) : "memory");
}

template <int SA, int SB>
__global__ void __launch_bounds__(threads_for_groups(2), 1)
conv_hstreamT_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                     const __grid_constant__ CUtensorMap tmB, const ConvKParams p) {
    using Plan = HsPlan<128, 2, SA, SB>;          // halo patch of an 8 x 32 pixel tile, 128-row weight tiles
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *gbase = smem_raw + (base - raw);
    const uint32_t sA = base + Plan::OFF_A;
    const uint32_t sB = base + Plan::OFF_B;
    const uint32_t sBar = base + Plan::OFF_BAR;
    auto fullA = [&](int s) { return sBar + 8u * s; };
    auto emptyA = [&](int s) { return sBar + 8u * (SA + s); };
    auto fullB = [&](int s) { return sBar + 8u * (2 * SA + s); };
    auto emptyB = [&](int s) { return sBar + 8u * (2 * SA + SB + s); };
    auto tfull_bar = [&](int b) { return sBar + 8u * (2 * SA + 2 * SB + b); };
    auto tempty_bar = [&](int b) { return sBar + 8u * (2 * SA + 2 * SB + 2 + b); };
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(gbase + Plan::OFF_TMEM);
    float *s_par = reinterpret_cast<float *>(gbase + Plan::OFF_PAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int chunks = p.chunks0 + p.chunks1;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA0);
        if (p.chunks1 > 0) prefetch_tmap(&tmA1);
        prefetch_tmap(&tmB);
        for (int s = 0; s < SA; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_ptr)), 512);
    if (warp >= 2) {
        for (int j = threadIdx.x - 64; j < 128; j += threads_for_groups(2) - 64) {
            s_par[j] = p.bias[j];
            s_par[128 + j] = p.scale[j];
            s_par[256 + j] = p.shift[j];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int ia = 0, ib = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int tx = tile % p.tiles_x;
                const int ty = (tile / p.tiles_x) % p.tiles_y;
                const int img = tile / tiles_per_img;
                const int x0 = tx * HT_W, y0 = ty * (HT_H * 2);
                for (int cc = 0; cc < chunks; ++cc, ++ia) {
                    const int sa = ia % SA;
                    mbar_wait(emptyA(sa), ((ia / SA) & 1) ^ 1u);
                    mbar_expect_tx(fullA(sa), Plan::A_BYTES_);
                    if (cc < p.chunks0)
                        tma_load_4d(sA + sa * Plan::A_SLOT, &tmA0, fullA(sa), cc * BK, x0 - 1, y0 - 1, img);
                    else
                        tma_load_4d(sA + sa * Plan::A_SLOT, &tmA1, fullA(sa), (cc - p.chunks0) * BK, x0 - 1, y0 - 1, img);
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap, ++ib) {
                        const int sb = ib % SB;
                        mbar_wait(emptyB(sb), ((ib / SB) & 1) ^ 1u);
                        mbar_expect_tx(fullB(sb), Plan::B_BYTES);
                        tma_load_2d(sB + sb * Plan::B_BYTES, &tmB, fullB(sb), (tap * chunks + cc) * BK, 0);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        {       // all lanes run the issue loop in converged code, one elected lane issues (umma_f16_w)
            constexpr uint32_t idesc = make_idesc(BM, 256);       // M = 128 output channels, N = 256 pixels
            int ia = 0, ib = 0, lt = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
                const int buf = lt & 1;
                mbar_wait(tempty_bar(buf), ((lt >> 1) & 1) ^ 1u);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * 256);
                for (int cc = 0; cc < chunks; ++cc, ++ia) {
                    const int sa = ia % SA;
                    mbar_wait(fullA(sa), (ia / SA) & 1);
                    tcgen05_fence_after();
                    const uint32_t a_slot = sA + sa * Plan::A_SLOT;
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap, ++ib) {
                        const int sb = ib % SB;
                        mbar_wait(fullB(sb), (ib / SB) & 1);
                        tcgen05_fence_after();
                        const uint64_t wdesc = make_sw128_desc(sB + sb * Plan::B_BYTES);
                        const uint64_t xdesc = make_sw128_desc_sbo(a_slot + static_cast<uint32_t>(((tap / 3) * HALO_W + (tap % 3)) * 128), HALO_W * 128);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16_w(tmem_d, wdesc + 2u * k, xdesc + 2u * k, idesc, (cc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                        umma_commit_w(emptyB(sb));
                    }
                    umma_commit_w(emptyA(sa));
                }
                umma_commit_w(tfull_bar(buf));
            }
        }
        __syncwarp();
    } else {
        const int e = warp - 2;
        const int quad = warp & 3;
        const int half = e >> 2;
        const uint32_t s_epi = base + Plan::OFF_EPI + static_cast<uint32_t>(e) * 2048u;       // 32 pixels x 64 B per warp
        int lt = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
            const int tx = tile % p.tiles_x;
            const int ty = (tile / p.tiles_x) % p.tiles_y;
            const int img = tile / tiles_per_img;
            epilogue_tile_T<HT_W>(p, s_par, s_epi, tmem_base, tfull_bar(lt & 1), tempty_bar(lt & 1), lt, quad, half, lane, tx * HT_W,
                                  ty * (HT_H * 2), img, (lt & 1) * 256);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// Weight gradient straight from the NHWC activations (no channel-major copies).
//
//   G[m][tap][n] += sum over pixels o of  A[sA*o + offA(tap)][m] * B[sB*o + offB(tap)][n]
//
// The contraction index (pixels) is the SLOW dimension of both NHWC operands, i.e. both are "MN-major" UMMA
// operands: one TMA box = an 8x8 (or 4x16 ...) pixel patch x 64 channels = 64 rows of 128 B (SWIZZLE_128B), which is
// exactly the canonical MN-major SW128 layout ((64 ch contiguous, LBO between 64-channel blocks), (8 pixel rows of
// 128 B, SBO = 1024 B between 8-pixel groups)).  A K = 16 MMA step consumes 16 consecutive pixels (2 KiB).
// Tap shifts are pixel coordinates of the box (never the innermost dimension), so there is no alignment constraint;
// out-of-image pixels are zero-filled by TMA = the conv's zero padding (and the ragged patch edges).
//   conv3x3 s1/s2: A = dz (pixel o), B = x (pixel s*o + tap - 1);  convT2x2: A = d(up) (pixel 2o + (dy,dx)), B = x.
//   Cm == 64, stride 1: the 128 accumulator rows hold TWO taps x 64 output channels: the shift is moved onto dz
//   (sum_o dz[o] x[o+t] = sum_p dz[p-t] x[p]) so the two 64-row A blocks use different shifts and B is unshifted.
// One CTA = one (m-tile, n-tile, tap item, K split) and adds its fp32 tile into G with vector reductions.
// ------------------------------------------------------------------------------------------
struct WgradNhwcParams {
    int N, Ho, Wo;
    int pw, ph;                // pixel patch of one K chunk (pw * ph = 64)
    int ptx, pty;              // patches per image
    int sA, sB;                // pixel stride of the two operands
    int kind;                  // 0 conv s1, 1 conv s2, 2 convT
    int pair;                  // 1: rows 0-63 / 64-127 of the tile are two taps of the same 64 output channels
    int taps, tap_items;
    int tpi, cn_tile;          // taps stacked along the N tile (non-pair conv kinds): N tile = tpi x cn_tile input channels
    int Cm, Cn;
    int m_tiles, n_tiles, splits;
    int per;                   // pixel patches per K split (every split is non-empty)
    int partial;               // 1: out is [splits][Cm][taps][Cn], each split stores its tile (no atomics: ordered reduce later)
    float *out;
    int out_ld, out_coff;
};
__device__ __forceinline__ void st_v4(float *addr, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(addr) = make_float4(a, b, c, d);
}

__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;      // between 64-element blocks along M / N
    d |= static_cast<uint64_t>(1024 >> 4) << 32;            // between 8-row groups along K
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int WG_BLOCK = 64 * 128;     // one 64-pixel x 64-channel box

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_nhwc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradNhwcParams p) {
    constexpr int A_TILE = 2 * WG_BLOCK;
    constexpr int B_TILE = (BN / 64) * WG_BLOCK;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *gbase = smem_raw + (base - raw);
    const uint32_t sA = base;
    const uint32_t sB = base + STAGES * A_TILE;
    const uint32_t sBar = sB + STAGES * B_TILE;
    auto full_bar = [&](int s) { return sBar + 8u * s; };
    auto empty_bar = [&](int s) { return sBar + 8u * (STAGES + s); };
    const uint32_t tfull_bar = sBar + 8u * (2 * STAGES);
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(gbase + STAGES * (A_TILE + B_TILE) + 8 * (2 * STAGES + 1));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int t = blockIdx.x;
    const int split = t % p.splits; t /= p.splits;
    const int item = t % p.tap_items; t /= p.tap_items;
    const int n_tile = t % p.n_tiles;
    const int m_tile = t / p.n_tiles;
    const int n0 = n_tile * p.cn_tile;
    // N tile = tpi taps x cn_tile input channels (narrow layers: one dz tile feeds several taps, so the L2 -> smem
    // traffic per MMA is that of a 256-column tile); the last item of a layer may hold fewer taps
    const int tap_base = item * p.tpi;
    const int nvalid = min(p.tpi, p.taps - tap_base);
    const int boxes_per_tap = p.cn_tile / 64;
    // the two 64-row blocks of the A tile: channel offset, tap, validity
    int blk_c[2], blk_tap[2];
    bool blk_ok[2];
    if (p.pair) {
        blk_c[0] = blk_c[1] = 0;
        blk_tap[0] = 2 * item;
        blk_tap[1] = 2 * item + 1;
        blk_ok[0] = true;
        blk_ok[1] = blk_tap[1] < p.taps;
        if (!blk_ok[1]) blk_tap[1] = blk_tap[0];
    } else {
        blk_c[0] = m_tile * 128;
        blk_c[1] = m_tile * 128 + 64;
        blk_tap[0] = blk_tap[1] = tap_base;
        blk_ok[0] = true;
        blk_ok[1] = blk_c[1] < p.Cm;
        if (!blk_ok[1]) blk_c[1] = blk_c[0];
    }
    const int patches = p.N * p.ptx * p.pty;
    const int per = p.per;
    const int k0 = split * per;
    const int k1 = min(patches, k0 + per);
    const int num_k_iters = max(0, k1 - k0);

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_ptr)), BN);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (num_k_iters > 0) {
        if (warp == 0) {
            if (lane == 0) {
                // pixel offsets of the boxes for this item
                int ax[2], ay[2];
                const bool b_shift = !(p.kind == 2 || p.pair);       // conv kinds without pairing: the tap shift sits on x
                const int n_boxes = p.pair || p.kind == 2 ? BN / 64 : nvalid * boxes_per_tap;
                for (int b = 0; b < 2; ++b) {
                    const int tp = blk_tap[b];
                    if (p.kind == 2) {            // convT: A = d(up) at 2o + (dy, dx)
                        ax[b] = tp & 1;
                        ay[b] = tp >> 1;
                    } else if (p.pair) {          // shift moved onto dz
                        ax[b] = -(tp % 3 - 1);
                        ay[b] = -(tp / 3 - 1);
                    } else {
                        ax[b] = ay[b] = 0;
                    }
                }
                const int ppi = p.ptx * p.pty;
                int it = 0;
                for (int k = k0; k < k1; ++k, ++it) {
                    const int img = k / ppi;
                    const int r = k - img * ppi;
                    const int oy0 = (r / p.ptx) * p.ph, ox0 = (r % p.ptx) * p.pw;
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_expect_tx(full_bar(s), A_TILE + n_boxes * WG_BLOCK);
                    for (int b = 0; b < 2; ++b)
                        tma_load_4d(sA + s * A_TILE + b * WG_BLOCK, &tmA, full_bar(s), blk_c[b], p.sA * ox0 + ax[b],
                                    p.sA * oy0 + ay[b], img);
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j) {
                        if (j >= n_boxes) break;
                        const int tj = j / boxes_per_tap, tap = tap_base + tj;
                        const int bx = b_shift ? tap % 3 - 1 : 0, by = b_shift ? tap / 3 - 1 : 0;
                        tma_load_4d(sB + s * B_TILE + j * WG_BLOCK, &tmB, full_bar(s), n0 + 64 * (j - tj * boxes_per_tap),
                                    p.sB * ox0 + bx, p.sB * oy0 + by, img);
                    }
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            {       // all lanes run the issue loop in converged code, one elected lane issues (umma_f16_w)
                // both operands MN-major: bits 15 / 16 of the instruction descriptor
                const int n_cols = p.pair || p.kind == 2 ? BN : nvalid * p.cn_tile;
                const uint32_t idesc = make_idesc(BM, n_cols) | (1u << 15) | (1u << 16);
                for (int it = 0; it < num_k_iters; ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tcgen05_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k) {      // 16 pixels (2 KiB) per MMA
                        const uint64_t adesc = make_sw128_mn_desc(sA + s * A_TILE + k * 2048, WG_BLOCK);
                        const uint64_t bdesc = make_sw128_mn_desc(sB + s * B_TILE + k * 2048, WG_BLOCK);
                        umma_f16_w(tmem_base, adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit_w(empty_bar(s));
                }
                umma_commit_w(tfull_bar);
            }
            __syncwarp();
        } else {
            const int e = warp - 2, quad = warp & 3, half = e >> 2;
            const int b = quad >> 1;                              // which 64-row block this TMEM lane quadrant holds
            const int m = blk_c[b] + (quad & 1) * 32 + lane;      // output channel (G row)
            mbar_wait(tfull_bar, 0);
            tcgen05_fence_after();
#pragma unroll 1
            for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c), r);
                const int tj = c / p.cn_tile;                    // tap slot of this 32-column chunk inside the N tile
                const int cc = c - tj * p.cn_tile;
                if (blk_ok[b] && m < p.Cm && tj < nvalid && n0 + cc < p.Cn) {
                    if (p.partial) {
                        float *dst = p.out + ((static_cast<size_t>(split) * p.Cm + m) * p.taps + blk_tap[b] + tj) * p.Cn + n0 + cc;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            st_v4(dst + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    } else {
                        float *dst = p.out + (static_cast<size_t>(m) * p.taps + blk_tap[b] + tj) * p.out_ld + p.out_coff + n0 + cc;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            red_add_v4(dst + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                       __uint_as_float(r[j + 3]));
                    }
                }
            }
            tcgen05_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

// ------------------------------------------------------------------------------------------
// Weight gradient of the full-resolution layers (conv3x3 stride 1, Cout = 64): halo variant.
//
// With one (tap pair, 64-pixel chunk) per MMA group the generic kernel moves 24 KiB of operands per 4 MMAs and is
// L2 -> smem bound (250-440 TFLOP/s).  Here one CTA owns ALL nine taps of a 64-input-channel slice: per 8x8 pixel
// patch it loads the x patch once (8 KiB) and the dz patch with a 1-pixel halo once (10x10 pixels, 12.5 KiB), and the
// five tap pairs are five accumulators (5 x 64 TMEM columns).  sum_o dz[o] x[o+t] = sum_p dz[p-t] x[p]: the A operand
// of tap t is the halo patch read from pixel offset off(t) = (2-ky)*10 + (2-kx); its 8-pixel K groups are 10 pixels
// (1280 B) apart = the descriptor's stride byte offset, and the second tap of a pair is the "second 64-row block" of
// the MN-major A tile at leading byte offset (off(t2) - off(t1)) * 128 B.  (The 128B swizzle is a function of the
// absolute smem address, so shifted starts read what TMA wrote -- same property the forward halo kernel relies on.)
// ------------------------------------------------------------------------------------------
constexpr int WH_A_SLOT = 13 * 1024;          // 10 x 10 pixels x 128 B = 12800 B, slot rounded to the 1 KiB swizzle period
constexpr int WH_B_SLOT = 8 * 1024;
constexpr int WH_STAGES = 6;
constexpr int WH_DYN = WH_STAGES * (WH_A_SLOT + WH_B_SLOT) + 8 * (2 * WH_STAGES + 1) + 16 + 1024;

__device__ __forceinline__ uint64_t make_sw128_mn_desc2(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_halo64_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradNhwcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *gbase = smem_raw + (base - raw);
    const uint32_t sA = base;
    const uint32_t sB = base + WH_STAGES * WH_A_SLOT;
    const uint32_t sBar = sB + WH_STAGES * WH_B_SLOT;
    auto full_bar = [&](int s) { return sBar + 8u * s; };
    auto empty_bar = [&](int s) { return sBar + 8u * (WH_STAGES + s); };
    const uint32_t tfull_bar = sBar + 8u * (2 * WH_STAGES);
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(gbase + WH_STAGES * (WH_A_SLOT + WH_B_SLOT) + 8 * (2 * WH_STAGES + 1));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int split = blockIdx.x % p.splits;
    const int n_chunk = blockIdx.x / p.splits;
    const int n0 = n_chunk * 64;
    const int patches = p.N * p.ptx * p.pty;
    const int per = p.per;
    const int k0 = split * per;
    const int k1 = min(patches, k0 + per);
    const int num_k_iters = max(0, k1 - k0);

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < WH_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_ptr)), 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (num_k_iters > 0) {
        if (warp == 0) {
            if (lane == 0) {
                const int ppi = p.ptx * p.pty;
                int it = 0;
                for (int k = k0; k < k1; ++k, ++it) {
                    const int img = k / ppi;
                    const int r = k - img * ppi;
                    const int oy0 = (r / p.ptx) * 8, ox0 = (r % p.ptx) * 8;
                    const int s = it % WH_STAGES;
                    const uint32_t ph = (it / WH_STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_expect_tx(full_bar(s), 100 * 128 + 64 * 128);
                    tma_load_4d(sA + s * WH_A_SLOT, &tmA, full_bar(s), 0, ox0 - 1, oy0 - 1, img);     // dz, 10 x 10 halo box
                    tma_load_4d(sB + s * WH_B_SLOT, &tmB, full_bar(s), n0, ox0, oy0, img);           // x, 8 x 8 box
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            {       // all lanes run the issue loop in converged code, one elected lane issues (umma_f16_w)
                constexpr uint32_t idesc = make_idesc(BM, 64) | (1u << 15) | (1u << 16);
                for (int it = 0; it < num_k_iters; ++it) {
                    const int s = it % WH_STAGES;
                    const uint32_t ph = (it / WH_STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tcgen05_fence_after();
                    const uint32_t a_slot = sA + s * WH_A_SLOT, b_slot = sB + s * WH_B_SLOT;
#pragma unroll
                    for (int pr = 0; pr < 5; ++pr) {
                        // pair pr = taps (8 - 2pr, 7 - 2pr); halo pixel offset of tap t: (2 - t/3) * 10 + (2 - t%3)
                        const int t1 = 8 - 2 * pr, t2 = pr < 4 ? 7 - 2 * pr : t1;
                        const int o1 = (2 - t1 / 3) * 10 + (2 - t1 % 3), o2 = (2 - t2 / 3) * 10 + (2 - t2 % 3);
                        const uint32_t lbo = pr < 4 ? static_cast<uint32_t>(o2 - o1) * 128u : 128u;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {      // 16 pixels = two 8-pixel patch rows per MMA
                            const uint64_t adesc = make_sw128_mn_desc2(a_slot + o1 * 128 + k * 2 * 1280, lbo, 1280);
                            const uint64_t bdesc = make_sw128_mn_desc2(b_slot + k * 2048, 8192, 1024);
                            umma_f16_w(tmem_base + static_cast<uint32_t>(pr * 64), adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit_w(empty_bar(s));
                }
                umma_commit_w(tfull_bar);
            }
            __syncwarp();
        } else {
            const int e = warp - 2, quad = warp & 3, half = e >> 2;
            const int blk = quad >> 1;                               // first / second tap of the pair
            const int m = (quad & 1) * 32 + lane;                    // output channel
            mbar_wait(tfull_bar, 0);
            tcgen05_fence_after();
#pragma unroll 1
            for (int c = half * 160; c < (half + 1) * 160; c += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c), r);
                const int pr = c >> 6;
                const int tap = 8 - 2 * pr - blk;
                if (tap >= 0 && p.partial) {
                    float *dst = p.out + ((static_cast<size_t>(split) * 64 + m) * 9 + tap) * p.Cn + n0 + (c & 63);
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        st_v4(dst + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                } else if (tap >= 0) {
                    float *dst = p.out + (static_cast<size_t>(m) * 9 + tap) * p.out_ld + p.out_coff + n0 + (c & 63);
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        red_add_v4(dst + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                   __uint_as_float(r[j + 3]));
                }
            }
            tcgen05_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// Weight gradient of the wide stride-1 layers (Cout, Cin multiples of 128): one filter ROW per CTA.
//
// The generic kernel streams 48 KiB of operands per 512 MMA cycles (94 B/clk/SM); both operands are private to the CTA,
// so nothing is shared in L2 and the chip-wide L2 -> SM bandwidth (~42 B/clk/SM) caps it at ~750 TFLOP/s.  Here a CTA
// owns 128 output channels x 128 input channels x the three taps (ky, 0..2): per 8x8 pixel patch it loads x once
// (2 x 8 KiB) and dz with a 1-pixel halo once (2 x 12.5 KiB) for 3 x 4 MMAs of 64 cycles = 54 B/clk/SM.  Same operand
// trick as wgrad_halo64_kernel: sum_o dz[o] x[o+t] = sum_p dz[p-t] x[p], tap t reads the halo patch from pixel offset
// (2-ky)*10 + (2-kx); the two 64-channel blocks of the A tile are the two halo slots (leading byte offset = slot size).
// ------------------------------------------------------------------------------------------
constexpr int W3_A_SLOT = 13 * 1024;
constexpr int W3_STAGE = 2 * W3_A_SLOT + 2 * WG_BLOCK;         // 42 KiB
constexpr int W3_STAGES = 5;
constexpr int W3_DYN = W3_STAGES * W3_STAGE + 8 * (2 * W3_STAGES + 1) + 16 + 1024;

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_row3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradNhwcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *gbase = smem_raw + (base - raw);
    const uint32_t sBar = base + W3_STAGES * W3_STAGE;
    auto full_bar = [&](int s) { return sBar + 8u * s; };
    auto empty_bar = [&](int s) { return sBar + 8u * (W3_STAGES + s); };
    const uint32_t tfull_bar = sBar + 8u * (2 * W3_STAGES);
    volatile uint32_t *tmem_ptr = reinterpret_cast<volatile uint32_t *>(gbase + W3_STAGES * W3_STAGE + 8 * (2 * W3_STAGES + 1));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int t = blockIdx.x;
    const int split = t % p.splits; t /= p.splits;
    const int ky = t % 3; t /= 3;
    const int n_tile = t % p.n_tiles;
    const int m_tile = t / p.n_tiles;
    const int m0 = m_tile * 128, n0 = n_tile * 128;
    const int patches = p.N * p.ptx * p.pty;
    const int k0 = split * p.per;
    const int k1 = min(patches, k0 + p.per);
    const int num_k_iters = max(0, k1 - k0);

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < W3_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t *>(tmem_ptr)), 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (num_k_iters > 0) {
        if (warp == 0) {
            if (lane == 0) {
                const int ppi = p.ptx * p.pty;
                int it = 0;
                for (int k = k0; k < k1; ++k, ++it) {
                    const int img = k / ppi;
                    const int r = k - img * ppi;
                    const int oy0 = (r / p.ptx) * 8, ox0 = (r % p.ptx) * 8;
                    const int s = it % W3_STAGES;
                    const uint32_t ph = (it / W3_STAGES) & 1;
                    const uint32_t st = base + s * W3_STAGE;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_expect_tx(full_bar(s), 2 * 100 * 128 + 2 * WG_BLOCK);
                    tma_load_4d(st, &tmA, full_bar(s), m0, ox0 - 1, oy0 - 1, img);                      // dz, 10 x 10 halo boxes
                    tma_load_4d(st + W3_A_SLOT, &tmA, full_bar(s), m0 + 64, ox0 - 1, oy0 - 1, img);
                    tma_load_4d(st + 2 * W3_A_SLOT, &tmB, full_bar(s), n0, ox0, oy0, img);              // x, 8 x 8 boxes
                    tma_load_4d(st + 2 * W3_A_SLOT + WG_BLOCK, &tmB, full_bar(s), n0 + 64, ox0, oy0, img);
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            {       // all lanes run the issue loop in converged code, one elected lane issues (umma_f16_w)
                constexpr uint32_t idesc = make_idesc(BM, 128) | (1u << 15) | (1u << 16);
                for (int it = 0; it < num_k_iters; ++it) {
                    const int s = it % W3_STAGES;
                    const uint32_t ph = (it / W3_STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tcgen05_fence_after();
                    const uint32_t a_slot = base + s * W3_STAGE, b_slot = a_slot + 2 * W3_A_SLOT;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int off = (2 - ky) * 10 + (2 - kx);          // halo pixel offset of tap (ky, kx)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {                      // 16 pixels = two 8-pixel patch rows per MMA
                            const uint64_t adesc = make_sw128_mn_desc2(a_slot + off * 128 + k * 2 * 1280, W3_A_SLOT, 1280);
                            const uint64_t bdesc = make_sw128_mn_desc2(b_slot + k * 2048, WG_BLOCK, 1024);
                            umma_f16_w(tmem_base + static_cast<uint32_t>(kx * 128), adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit_w(empty_bar(s));
                }
                umma_commit_w(tfull_bar);
            }
            __syncwarp();
        } else {
            const int e = warp - 2, quad = warp & 3, half = e >> 2;
            const int m = m0 + quad * 32 + lane;                       // output channel (TMEM lane)
            mbar_wait(tfull_bar, 0);
            tcgen05_fence_after();
#pragma unroll 1
            for (int c = half * 192; c < (half + 1) * 192; c += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c), r);
                const int tap = ky * 3 + (c >> 7);
                const int n = n0 + (c & 127);
                if (p.partial) {
                    float *dst = p.out + ((static_cast<size_t>(split) * p.Cm + m) * 9 + tap) * p.Cn + n;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        st_v4(dst + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                } else {
                    float *dst = p.out + (static_cast<size_t>(m) * 9 + tap) * p.out_ld + p.out_coff + n;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        red_add_v4(dst + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                   __uint_as_float(r[j + 3]));
                }
            }
            tcgen05_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// first layer: normalise + pad + Conv2d(1,C,3,p=1) + act + BN(eval), CUDA cores (K = 9)
// ------------------------------------------------------------------------------------------
// Block = 32x8 output pixels.  The normalised (and padded) input patch incl. its 1-pixel halo is staged
// once in shared memory (one IEEE division per input pixel instead of nine per output channel group);
// thread (g, px) keeps the 72 weights of its 8 output channels in registers and walks 8 rows, so the
// only smem traffic in the inner loop is the 9 input taps; a warp stores 4 pixels x 128 B contiguously.
constexpr int FC_TW = 32, FC_TH = 36;   // tall tiles (a multiple of 3 rows): the 96 per-thread weight/affine loads amortise over 36 rows
static_assert(FC_TH % 3 == 0, "the row loop rotates three register rows");

// Persistent: each CTA keeps its 8-channel slice of the weights / affine in registers and walks 32x36-pixel tiles.
template <typename T, bool RELU>
__global__ void __launch_bounds__(256, 2)
first_conv_kernel(const T *__restrict__ img, int H, int W, int pad_y, int pad_x, float lo, float hi,
                  const float *__restrict__ lohi_dev, const float *__restrict__ weight, const float *__restrict__ bias,
                  const float *__restrict__ scale, const float *__restrict__ shift, int C, int act,
                  __nv_bfloat16 *__restrict__ out, int ld, int coff, int tiles_x, int num_tiles) {
    __shared__ float s_in[FC_TH + 2][FC_TW + 2];
    if (lohi_dev) {                 // frame min / max computed on the device (mbs_frame_minmax)
        lo = lohi_dev[0];
        hi = lohi_dev[1];
    }
    const int Hp = H + pad_y, Wp = W + pad_x;
    const float range = hi - lo;
    const int g = threadIdx.x & 7;           // channel group (8 channels) inside a 64-channel slab
    const int px = threadIdx.x >> 3;         // pixel column inside the tile
    const size_t row_stride = static_cast<size_t>(Wp) * ld;
    for (int cbase = 0; cbase < C; cbase += 64) {
        const int c0 = cbase + g * 8;
        const bool live = c0 < C;
        // channel pairs in 64-bit registers: one FFMA2 (fma.rn.f32x2, each half an IEEE fma) per tap and pair
        uint64_t w2[4][9], b2[4], sc2[4], sh2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int ca = live ? c0 + 2 * u : 0, cb = live ? c0 + 2 * u + 1 : 0;
#pragma unroll
            for (int t = 0; t < 9; ++t) w2[u][t] = pack_f32x2(weight[ca * 9 + t], weight[cb * 9 + t]);
            b2[u] = pack_f32x2(bias[ca], bias[cb]);
            sc2[u] = pack_f32x2(scale[ca], scale[cb]);
            sh2[u] = pack_f32x2(shift[ca], shift[cb]);
        }
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int x0 = (tile % tiles_x) * FC_TW, y0 = (tile / tiles_x) * FC_TH;
            __syncthreads();         // everyone is done with the previous tile's patch
            // (prefetching the next patch into registers during the FMA loop was measured: no gain, more spills)
            for (int i = threadIdx.x; i < (FC_TH + 2) * (FC_TW + 2); i += 256) {
                const int r = i / (FC_TW + 2), c = i - r * (FC_TW + 2);
                const int yy = y0 + r - 1, xx = x0 + c - 1;
                float v = 0.0f;  // conv zero padding outside the (padded) model input
                if (yy >= 0 && yy < Hp && xx >= 0 && xx < Wp) {
                    float raw = lo;  // zero_pad_model_input pad value = frame min (utils.py:124, infer.py:256)
                    if (yy >= pad_y && xx >= pad_x) raw = static_cast<float>(img[static_cast<size_t>(yy - pad_y) * W + (xx - pad_x)]);
                    // 2 * (f32(img) - min) / (max - min) - 1, evaluated left to right in f32 (infer.py:346);
                    // hi < lo: the caller already normalised the image (drop-in net(x) path) -> pass through
                    v = hi < lo ? raw : __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(raw, lo)), range), 1.0f);
                }
                s_in[r][c] = v;
            }
            __syncthreads();
            if (!live) continue;
            const int X = x0 + px;
            __nv_bfloat16 *drow = out + (static_cast<size_t>(y0) * Wp + X) * ld + coff + c0;
            // one output row from the three input rows top / mid / bot (3 floats each)
            auto do_row = [&](int r, const float (&top)[3], const float (&mid)[3], const float (&bot)[3]) {
                uint64_t acc2[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint64_t a = b2[u];
#pragma unroll
                    for (int t = 0; t < 3; ++t) a = fma_f32x2(pack_f32x2(top[t], top[t]), w2[u][t], a);
#pragma unroll
                    for (int t = 0; t < 3; ++t) a = fma_f32x2(pack_f32x2(mid[t], mid[t]), w2[u][3 + t], a);
#pragma unroll
                    for (int t = 0; t < 3; ++t) a = fma_f32x2(pack_f32x2(bot[t], bot[t]), w2[u][6 + t], a);
                    acc2[u] = a;
                }
                float v[8];
#pragma unroll
                for (int u = 0; u < 4; ++u) unpack_f32x2(acc2[u], v[2 * u], v[2 * u + 1]);
                if (RELU) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = fmaxf(v[u], 0.0f);
                } else if (act != MBS_ACT_NONE) {        // one uniform branch per row
#pragma unroll 1
                    for (int u = 0; u < 8; ++u) v[u] = apply_act(v[u], act);
                }
                uint32_t packed[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float a0, a1;
                    unpack_f32x2(fma_f32x2(pack_f32x2(v[2 * u], v[2 * u + 1]), sc2[u], sh2[u]), a0, a1);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);
                    packed[u] = *reinterpret_cast<uint32_t *>(&h2);
                }
                if (y0 + r < Hp && X < Wp)
                    *reinterpret_cast<uint4 *>(drow + r * row_stride) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            };
            float ra[3], rb[3], rc[3];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                ra[t] = s_in[0][px + t];
                rb[t] = s_in[1][px + t];
            }
            // rows in groups of three with rotating register names: 3 new smem reads per output row, no moves
#pragma unroll 1
            for (int r = 0; r < FC_TH; r += 3) {
#pragma unroll
                for (int t = 0; t < 3; ++t) rc[t] = s_in[r + 2][px + t];
                do_row(r, ra, rb, rc);
#pragma unroll
                for (int t = 0; t < 3; ++t) ra[t] = s_in[r + 3][px + t];
                do_row(r + 1, rb, rc, ra);
#pragma unroll
                for (int t = 0; t < 3; ++t) rb[t] = s_in[r + 4][px + t];
                do_row(r + 2, rc, ra, rb);
            }
        }
    }
}

// frame min / max (np.min / np.max of the raw frame, infer_script_local.py:124) as ordered uint keys
__device__ __forceinline__ unsigned int order_key(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float order_unkey(unsigned int k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
template <typename T>
__device__ __forceinline__ unsigned int to_key(T v) { return static_cast<unsigned int>(v); }
template <>
__device__ __forceinline__ unsigned int to_key<float>(float v) { return order_key(v); }

__global__ void minmax_init_kernel(unsigned int *keys) {
    keys[0] = 0xFFFFFFFFu;
    keys[1] = 0u;
}
template <typename T>
__global__ void minmax_kernel(const T *__restrict__ img, long long n, unsigned int *keys) {
    unsigned int lo = 0xFFFFFFFFu, hi = 0u;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const unsigned int k = to_key<T>(img[i]);
        lo = k < lo ? k : lo;
        hi = k > hi ? k : hi;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned int a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&keys[0], lo);
        atomicMax(&keys[1], hi);
    }
}
__global__ void minmax_final_kernel(const unsigned int *keys, int is_float, float *lohi) {
    lohi[0] = is_float ? order_unkey(keys[0]) : static_cast<float>(keys[0]);
    lohi[1] = is_float ? order_unkey(keys[1]) : static_cast<float>(keys[1]);
}

__global__ void pack_conv3x3_kernel(const float *__restrict__ w, int Cout, int Cin, __nv_bfloat16 *__restrict__ out) {
    // out[o][t][c] = w[o][c][t]
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(Cout) * 9 * Cin;
    if (i >= total) return;
    const int c = static_cast<int>(i % Cin);
    const int t = static_cast<int>((i / Cin) % 9);
    const int o = static_cast<int>(i / (static_cast<long long>(Cin) * 9));
    out[i] = __float2bfloat16_rn(w[(static_cast<size_t>(o) * Cin + c) * 9 + t]);
}

__global__ void pack_convT2x2_kernel(const float *__restrict__ w, int Cin, int Cout, __nv_bfloat16 *__restrict__ out) {
    // out[(q*Cout + co)][ci] = w[ci][co][q],  q = dy*2+dx
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(4) * Cout * Cin;
    if (i >= total) return;
    const int ci = static_cast<int>(i % Cin);
    const int row = static_cast<int>(i / Cin);
    const int q = row / Cout, co = row % Cout;
    out[i] = __float2bfloat16_rn(w[(static_cast<size_t>(ci) * Cout + co) * 4 + q]);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

// NHWC bf16 activation view -> 4-D tensor map (C, W, H, N), box (64, bw, bh, 1), traversal stride es.
int make_act_map(CUtensorMap *map, const void *base, int N, int H, int W, int C, int ld, int coff, int es,
                 int box_w = TILE_W, int box_h = TILE_H) {
    EncodeTiledFn enc = get_encode_fn();
    MBS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    const char *p = static_cast<const char *>(base) + static_cast<size_t>(coff) * 2;
    MBS_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "activation view must be 16-byte aligned");
    MBS_REQUIRE((ld * 2) % 16 == 0, "activation pixel stride must be a multiple of 16 bytes");
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                          static_cast<cuuint64_t>(N)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(W) * ld * 2,
                             static_cast<cuuint64_t>(H) * W * ld * 2};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_w * es),
                         static_cast<cuuint32_t>(box_h * es), 1};
    cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(es), static_cast<cuuint32_t>(es), 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char *>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MBS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation N=%d H=%d W=%d C=%d ld=%d es=%d) failed: %d", N,
                H, W, C, ld, es, static_cast<int>(r));
    return 0;
}

// packed weights [rows][K] bf16 (K contiguous) -> 2-D tensor map, box (64, bn)
int make_weight_map(CUtensorMap *map, const void *base, int rows, int K, int bn) {
    EncodeTiledFn enc = get_encode_fn();
    MBS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    MBS_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "weights must be 16-byte aligned");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(bn)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MBS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights rows=%d K=%d) failed: %d", rows, K,
                static_cast<int>(r));
    return 0;
}

int sm_count() {
    static int n[mbs::kMaxDevices] = {0};
    const int dev = mbs::current_device();
    if (n[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n[dev] = v;
    }
    return n[dev];
}

// destination of a transposed conv as a 5-D tensor (C, dx, x, dy, y'): pixel (2y+dy, 2x+dx) of image n is
// element (c, dx, x, dy, n*H_in + y); one store box = the (dy,dx) quadrant of an 8x16 input patch.
int make_convT_store_map(CUtensorMap *map, const void *base, int N, int H_in, int W_in, int C, int ld, int coff) {
    EncodeTiledFn enc = get_encode_fn();
    MBS_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    const char *p = static_cast<const char *>(base) + static_cast<size_t>(coff) * 2;
    const cuuint64_t px = static_cast<cuuint64_t>(ld) * 2;          // bytes per output pixel
    const cuuint64_t Wd = 2ull * W_in;
    cuuint64_t dims[5] = {static_cast<cuuint64_t>(C), 2, static_cast<cuuint64_t>(W_in), 2,
                          static_cast<cuuint64_t>(N) * H_in};
    cuuint64_t strides[4] = {px, 2 * px, Wd * px, 2 * Wd * px};
    cuuint32_t box[5] = {static_cast<cuuint32_t>(BK), 1, static_cast<cuuint32_t>(TILE_W), 1, static_cast<cuuint32_t>(TILE_H)};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<char *>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MBS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(convT destination) failed: %d", static_cast<int>(r));
    return 0;
}

template <int BN, int STAGES, int CTAS_PER_SM, bool TMA_EPI, int MT = 1, int NG = 2>
int launch_conv(const CUtensorMap &a0, const CUtensorMap &a1, const CUtensorMap &b, const CUtensorMap &dmap,
                const ConvKParams &kp, cudaStream_t stream) {
    using Plan = SmemPlan<BN, STAGES, TMA_EPI, MT, NG>;
    static int configured_bytes[mbs::kMaxDevices] = {0};
    const int dyn = Plan::dyn_bytes(kp.Cout);
    const int dev = mbs::current_device();
    if (dyn > configured_bytes[dev]) {
        MBS_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, STAGES, CTAS_PER_SM, TMA_EPI, MT, NG>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
        configured_bytes[dev] = dyn;
    }
    const int grid = kp.num_tiles < sm_count() * CTAS_PER_SM ? kp.num_tiles : sm_count() * CTAS_PER_SM;
    conv_gemm_kernel<BN, STAGES, CTAS_PER_SM, TMA_EPI, MT, NG><<<grid, threads_for_groups(NG), dyn, stream>>>(a0, a1, b, dmap, kp);
    MBS_CHECK_LAUNCH();
    return 0;
}

template <int CHUNKS, int STAGES, int NG>
int launch_halo(const CUtensorMap &a0, const CUtensorMap &a1, const CUtensorMap &b, const CUtensorMap &dmap,
                const ConvKParams &kp, cudaStream_t stream) {
    using Plan = HaloPlan<CHUNKS, STAGES, NG>;
    static bool configured[mbs::kMaxDevices] = {false};
    const int dev = mbs::current_device();
    if (!configured[dev]) {
        MBS_CHECK_CUDA(cudaFuncSetAttribute(conv_halo64_kernel<CHUNKS, STAGES, NG>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::DYN_BYTES));
        configured[dev] = true;
    }
    const int grid = kp.num_tiles < sm_count() ? kp.num_tiles : sm_count();
    conv_halo64_kernel<CHUNKS, STAGES, NG><<<grid, threads_for_groups(NG), Plan::DYN_BYTES, stream>>>(a0, a1, b, dmap, kp,
                                                                                                       FirstParams{});
    MBS_CHECK_LAUNCH();
    return 0;
}

template <typename T, bool RELU1>
int launch_halo_first(const CUtensorMap &b, const CUtensorMap &dmap, const ConvKParams &kp, const FirstParams &fp,
                      cudaStream_t stream) {
    constexpr int STAGES = 4, NG = 2;
    using Plan = HaloPlan<1, STAGES, NG>;
    constexpr int DYN = Plan::DYN_BYTES + 1024;      // + the producers' input patch behind the epilogue parameters
    static_assert(FIRST_IN_H * FIRST_IN_W * 4 <= 1024, "input patch");
    static bool configured[mbs::kMaxDevices] = {false};
    const int dev = mbs::current_device();
    if (!configured[dev]) {
        MBS_CHECK_CUDA(cudaFuncSetAttribute(conv_halo64_kernel<1, STAGES, NG, true, T, RELU1>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, DYN));
        configured[dev] = true;
    }
    const int grid = kp.num_tiles < sm_count() ? kp.num_tiles : sm_count();
    conv_halo64_kernel<1, STAGES, NG, true, T, RELU1>
        <<<grid, threads_for_groups(NG) + FIRST_THREADS, DYN, stream>>>(b, b, b, dmap, kp, fp);
    MBS_CHECK_LAUNCH();
    return 0;
}

template <int CHUNKS, int STAGES, int NG>
int launch_halo_pair(const CUtensorMap &a0, const CUtensorMap &a1, const CUtensorMap &b, const CUtensorMap &dmap,
                     const ConvKParams &kp, cudaStream_t stream) {
    using Plan = HaloPairPlan<CHUNKS, STAGES, NG>;
    static int max_clusters[mbs::kMaxDevices] = {0};
    const int dev = mbs::current_device();
    if (!max_clusters[dev]) {
        MBS_CHECK_CUDA(cudaFuncSetAttribute(conv_halo64_pair_kernel<CHUNKS, STAGES, NG>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::DYN_BYTES));
        // persistent kernel: exactly as many 2-CTA clusters as can be co-resident (GPCs with an odd SM count leave one SM out)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sm_count() & ~1);
        cfg.blockDim = dim3(threads_for_groups(NG));
        cfg.dynamicSmemBytes = Plan::DYN_BYTES;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, conv_halo64_pair_kernel<CHUNKS, STAGES, NG>, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = sm_count() / 2;
        }
        max_clusters[dev] = n;
        if (getenv("MBS_VERBOSE")) fprintf(stderr, "[mbseg] halo pair kernel: %d co-resident 2-CTA clusters\n", n);
    }
    int grid = 2 * max_clusters[dev];
    if (grid > ((kp.num_tiles + 1) & ~1)) grid = (kp.num_tiles + 1) & ~1;
    conv_halo64_pair_kernel<CHUNKS, STAGES, NG><<<grid, threads_for_groups(NG), Plan::DYN_BYTES, stream>>>(a0, a1, b, dmap, kp);
    MBS_CHECK_LAUNCH();
    return 0;
}

template <int BN, int MT, int SA, int SB>
int launch_hstream(const CUtensorMap &a0, const CUtensorMap &a1, const CUtensorMap &b, const ConvKParams &kp, cudaStream_t stream) {
    using Plan = HsPlan<BN, MT, SA, SB>;
    static int configured_bytes[mbs::kMaxDevices] = {0};
    const int dyn = Plan::dyn_bytes(kp.Cout);
    const int dev = mbs::current_device();
    if (dyn > configured_bytes[dev]) {
        MBS_CHECK_CUDA(cudaFuncSetAttribute(conv_hstream_kernel<BN, MT, SA, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
        configured_bytes[dev] = dyn;
    }
    const int grid = kp.num_tiles < sm_count() ? kp.num_tiles : sm_count();
    conv_hstream_kernel<BN, MT, SA, SB><<<grid, threads_for_groups(2), dyn, stream>>>(a0, a1, b, kp);
    MBS_CHECK_LAUNCH();
    return 0;
}

template <int SA, int SB>
int launch_hstreamT(const CUtensorMap &a0, const CUtensorMap &a1, const CUtensorMap &b, const ConvKParams &kp, cudaStream_t stream) {
    using Plan = HsPlan<128, 2, SA, SB>;
    static int configured_bytes[mbs::kMaxDevices] = {0};
    const int dyn = Plan::dyn_bytes(128);
    const int dev = mbs::current_device();
    if (dyn > configured_bytes[dev]) {
        MBS_CHECK_CUDA(cudaFuncSetAttribute(conv_hstreamT_kernel<SA, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
        configured_bytes[dev] = dyn;
    }
    const int grid = kp.num_tiles < sm_count() ? kp.num_tiles : sm_count();
    conv_hstreamT_kernel<SA, SB><<<grid, threads_for_groups(2), dyn, stream>>>(a0, a1, b, kp);
    MBS_CHECK_LAUNCH();
    return 0;
}

int hstream_enabled() {     // MBS_NO_HSTREAM=1 (A/B runs): wide stride-1 convs through the generic kernel; =2: no role-swapped Cout = 128 variant
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MBS_NO_HSTREAM");
        v = (e && e[0] == '1') ? 0 : ((e && e[0] == '2') ? 2 : 1);
    }
    return v;
}

int epi_variant() {     // MBS_EPI_VARIANT (A/B runs): 0 default; 1 = transposed convs on 128-column tiles; 3 = 4 epilogue groups in the halo kernel
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MBS_EPI_VARIANT");
        v = e ? atoi(e) : 0;
    }
    return v;
}

bool halo_pair_enabled() {      // MBS_NO_HALO_PAIR=1 (A/B runs): full-resolution Cout = 64 layers on single CTAs
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MBS_NO_HALO_PAIR");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

bool wgrad_halo_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MBS_NO_WGRAD_HALO");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

bool wgrad_row3_enabled() {      // MBS_NO_WGRAD_ROW3=1 (A/B runs): wide stride-1 layers through the generic kernel
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MBS_NO_WGRAD_ROW3");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

bool wgrad_stack_enabled() {     // MBS_NO_WGRAD_STACK=1 (A/B runs): one tap per N tile for the narrow layers
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MBS_NO_WGRAD_STACK");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

bool pair_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MBS_NO_PAIR");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

bool halo_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MBS_NO_HALO");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

}  // namespace

extern "C" int mbs_conv_gemm(const mbs_conv_desc *d, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(d != nullptr, "null descriptor");
    MBS_REQUIRE(d->mode >= 0 && d->mode <= 3, "bad conv mode %d", d->mode);
    MBS_REQUIRE(d->C0 > 0 && d->C0 % BK == 0 && d->C1 >= 0 && d->C1 % BK == 0,
                "input channels must be multiples of %d (got %d + %d)", BK, d->C0, d->C1);
    MBS_REQUIRE(d->Cout > 0 && d->Cout % 64 == 0, "Cout must be a multiple of 64 (got %d)", d->Cout);
    MBS_REQUIRE(d->mode != MBS_CONVT2X2_S2 || d->C1 == 0, "transposed conv takes a single source");
    const bool strided = d->mode == MBS_CONV3X3_S2 || d->mode == MBS_CONV2X2_S2;
    MBS_REQUIRE(!strided || (d->H % 2 == 0 && d->W % 2 == 0), "stride-2 conv needs even H, W");
    MBS_REQUIRE(d->head_out == nullptr ||
                    (d->Cout == 64 && d->mode == MBS_CONV3X3_S1 && d->head_w && d->head_n >= 1 && d->head_n <= 4),
                "fused head needs Cout == 64, stride-1 conv, head weights and 1 <= head_n <= 4");
    MBS_REQUIRE(d->dst != nullptr || d->head_out != nullptr, "no output requested");
    if (d->dst) {
        MBS_REQUIRE(((reinterpret_cast<uintptr_t>(d->dst) + static_cast<size_t>(d->coffd) * 2) & 15) == 0 &&
                        (d->ldd * 2) % 16 == 0,
                    "destination view must be 16-byte aligned");
    }

    ConvKParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.mode = d->mode;
    kp.act = d->act;
    const int es = strided ? 2 : 1;
    kp.Hm = strided ? d->H / 2 : d->H;
    kp.Wm = strided ? d->W / 2 : d->W;
    kp.tiles_x = mbs::cdiv(kp.Wm, TILE_W);
    kp.tiles_y = mbs::cdiv(kp.Hm, TILE_H);
    kp.chunks0 = d->C0 / BK;
    kp.chunks1 = d->C1 / BK;
    kp.taps = d->mode == MBS_CONVT2X2_S2 ? 1 : (d->mode == MBS_CONV2X2_S2 ? 4 : 9);
    kp.Cout = d->Cout;
    kp.bias = d->bias;
    kp.scale = d->scale;
    kp.shift = d->shift;
    kp.dst = static_cast<__nv_bfloat16 *>(d->dst);
    kp.ldd = d->ldd;
    kp.coffd = d->coffd;
    kp.Hd = d->mode == MBS_CONVT2X2_S2 ? 2 * d->H : kp.Hm;
    kp.Wd = d->mode == MBS_CONVT2X2_S2 ? 2 * d->W : kp.Wm;
    kp.head_w = d->head_out ? d->head_w : nullptr;
    kp.head_n = d->head_out ? d->head_n : 0;
    for (int h = 0; h < 4; ++h) kp.head_b[h] = d->head_b[h];
    kp.head_out = d->head_out;

    if (halo_enabled() && d->mode == MBS_CONV3X3_S1 && d->Cout == 64 && d->C0 == 64 && (d->C1 == 0 || d->C1 == 64)) {
        // full-resolution layers: halo tiles + resident weights (see conv_halo64_kernel)
        kp.tiles_x = mbs::cdiv(kp.Wm, HT_W);
        kp.tiles_y = mbs::cdiv(kp.Hm, HT_H);
        kp.n_tiles = 1;
        const long long tiles_ll = static_cast<long long>(d->N) * kp.tiles_x * kp.tiles_y;
        MBS_REQUIRE(tiles_ll > 0 && tiles_ll < (1ll << 31), "too many tiles");
        kp.num_tiles = static_cast<int>(tiles_ll);
        CUtensorMap a0, a1, b;
        int rc = make_act_map(&a0, d->src0, d->N, d->H, d->W, d->C0, d->ld0, d->coff0, 1, HALO_W, HALO_H);
        if (rc) return rc;
        if (d->C1 > 0) {
            rc = make_act_map(&a1, d->src1, d->N, d->H, d->W, d->C1, d->ld1, d->coff1, 1, HALO_W, HALO_H);
            if (rc) return rc;
        } else {
            a1 = a0;
        }
        // measured (2048^2): two sources 0.529 -> 0.459 ms as a pair; one source 0.276 ms either way (the pair costs two SMs per
        // cluster of scheduling freedom and gains nothing there)
        const bool pair2 = halo_pair_enabled() && d->C1 > 0 && kp.num_tiles >= 2 * sm_count();
        rc = make_weight_map(&b, d->weight, 64, 9 * (d->C0 + d->C1), pair2 ? 32 : 64);
        if (rc) return rc;
        CUtensorMap dm = a0;
        if (d->dst) {
            rc = make_act_map(&dm, d->dst, d->N, kp.Hd, kp.Wd, d->Cout, d->ldd, d->coffd, 1, HT_W, HT_H);
            if (rc) return rc;
        }
        if (pair2) {
            // 2-CTA pairs (cta_group::2): each CTA supplies half of every weight tile (see conv_halo64_pair_kernel); the
            // accumulation order is that of the single-CTA kernel, so results are bit-identical whichever runs
            return launch_halo_pair<2, 4, 2>(a0, a1, b, dm, kp, stream);
        }
        if (d->C1 > 0) return launch_halo<2, 2, 2>(a0, a1, b, dm, kp, stream);
        // measured: 4 epilogue groups do not help here (64->64 @2048^2: 0.333 vs 0.324 ms) -- the MMAs' own smem
        // reads bound this layer, and more epilogue warps compete for the same smem bandwidth
        if (epi_variant() == 3) return launch_halo<1, 3, 4>(a0, a1, b, dm, kp, stream);
        return launch_halo<1, 4, 2>(a0, a1, b, dm, kp, stream);
    }

    if (hstream_enabled() && d->mode == MBS_CONV3X3_S1 && d->Cout % 128 == 0 && d->Cout <= 1024 && d->head_out == nullptr) {
        // wide stride-1 layers: halo tiles + streamed weights (see conv_hstream_kernel); the choice depends on the layer
        // shape only, never on the frame size, so tiled and whole-frame inference stay bit-identical
        const bool wide = d->Cout % 256 == 0;
        const int mt = wide ? 1 : 2;
        kp.tiles_x = mbs::cdiv(kp.Wm, HT_W);
        kp.tiles_y = mbs::cdiv(kp.Hm, HT_H * mt);
        kp.n_tiles = d->Cout / (wide ? 256 : 128);
        const long long tiles_ll = static_cast<long long>(d->N) * kp.tiles_x * kp.tiles_y * kp.n_tiles;
        MBS_REQUIRE(tiles_ll > 0 && tiles_ll < (1ll << 31), "too many tiles");
        kp.num_tiles = static_cast<int>(tiles_ll);
        CUtensorMap a0, a1, b;
        int rc = make_act_map(&a0, d->src0, d->N, d->H, d->W, d->C0, d->ld0, d->coff0, 1, HALO_W, mt * HT_H + 2);
        if (rc) return rc;
        if (d->C1 > 0) {
            rc = make_act_map(&a1, d->src1, d->N, d->H, d->W, d->C1, d->ld1, d->coff1, 1, HALO_W, mt * HT_H + 2);
            if (rc) return rc;
        } else {
            a1 = a0;
        }
        rc = make_weight_map(&b, d->weight, d->Cout, 9 * (d->C0 + d->C1), wide ? 256 : 128);
        if (rc) return rc;
        if (wide) return launch_hstream<256, 1, 2, 4>(a0, a1, b, kp, stream);
        // Cout == 128: operand roles swapped (channels = M rows, 256 pixels = N columns), see conv_hstreamT_kernel
        if (d->Cout == 128 && hstream_enabled() != 2) return launch_hstreamT<2, 6>(a0, a1, b, kp, stream);
        return launch_hstream<128, 2, 2, 6>(a0, a1, b, kp, stream);
    }

    const int ncols = d->mode == MBS_CONVT2X2_S2 ? 4 * d->Cout : d->Cout;
    const int K = kp.taps * (d->C0 + d->C1);
    int bn = ncols % 256 == 0 ? 256 : (ncols % 128 == 0 ? 128 : 64);
    // transposed convs have a short K loop and are epilogue/bandwidth bound.  Up to 256 input channels: all four
    // (dy,dx) taps of 64 output channels in one 256-column tile (A read once), TMA-store epilogue with 4 groups
    // (measured 128->64 @1024^2: 0.146 ms vs 0.202 with 2 groups on 128-column tiles; 256->128 @512^2: 0.102 vs
    // 0.122).  Deeper ones keep the direct epilogue (512->256 @256^2: 0.083 both; 1024->512 @128^2: 0.064 vs 0.072).
    // (the TMA-store path needs whole patches per image in its folded (N*H) dimension)
    const bool convT_tma = d->mode == MBS_CONVT2X2_S2 && d->C0 <= 256 && (d->N == 1 || d->H % TILE_H == 0);
    if (convT_tma) bn = (epi_variant() != 1 && ncols % 256 == 0) ? 256 : 128;
    kp.n_tiles = ncols / bn;
    // Cout = 128 layers: two stacked pixel tiles per work item share each weight tile (see SmemPlan)
    const bool pair = pair_enabled() && bn == 128 && d->mode != MBS_CONVT2X2_S2 &&
                      static_cast<long long>(d->N) * kp.tiles_x * mbs::cdiv(kp.Hm, 2 * TILE_H) >= sm_count();
    const int mt = pair ? 2 : 1;
    kp.tiles_y = mbs::cdiv(kp.Hm, TILE_H * mt);

    CUtensorMap a0, a1, b;
    int rc = make_act_map(&a0, d->src0, d->N, d->H, d->W, d->C0, d->ld0, d->coff0, es, TILE_W, TILE_H * mt);
    if (rc) return rc;
    if (d->C1 > 0) {
        rc = make_act_map(&a1, d->src1, d->N, d->H, d->W, d->C1, d->ld1, d->coff1, es, TILE_W, TILE_H * mt);
        if (rc) return rc;
    } else {
        a1 = a0;
    }
    rc = make_weight_map(&b, d->weight, ncols, K, bn);
    if (rc) return rc;

    const long long tiles_ll = static_cast<long long>(d->N) * kp.tiles_x * kp.tiles_y * kp.n_tiles;
    MBS_REQUIRE(tiles_ll > 0 && tiles_ll < (1ll << 31), "too many tiles");
    kp.num_tiles = static_cast<int>(tiles_ll);
    MBS_REQUIRE(d->Cout <= 1024, "Cout > 1024 is not supported by the epilogue parameter staging");
    if (convT_tma) {
        // short-K transposed conv: TMA-store epilogue (5-D pixel-shuffle view of the destination), 4 epilogue groups
        CUtensorMap dm = a0;
        if (d->dst) {
            rc = make_convT_store_map(&dm, d->dst, d->N, d->H, d->W, d->Cout, d->ldd, d->coffd);
            if (rc) return rc;
        }
        if (bn == 256) return launch_conv<256, 3, 1, true, 1, 4>(a0, a1, b, dm, kp, stream);
        return launch_conv<128, 5, 1, true>(a0, a1, b, dm, kp, stream);
    }
    if (bn == 256) return launch_conv<256, 4, 1, false>(a0, a1, b, a0, kp, stream);
    // measured on B200: for the generic BN <= 128 convs two resident CTAs with the direct epilogue beat one CTA
    // with the TMA-store epilogue (128->128 @1024^2: 0.295 vs 0.378 ms)
    if (pair) return launch_conv<128, 4, 1, false, 2>(a0, a1, b, a0, kp, stream);
    if (bn == 128) return launch_conv<128, 3, 2, false>(a0, a1, b, a0, kp, stream);
    return launch_conv<64, 4, 2, false>(a0, a1, b, a0, kp, stream);
}

template <int BN, int STAGES>
int launch_wgrad_nhwc(const CUtensorMap &a, const CUtensorMap &b, const WgradNhwcParams &wp, int grid, cudaStream_t stream) {
    constexpr int dyn = STAGES * (2 * WG_BLOCK + (BN / 64) * WG_BLOCK) + 8 * (2 * STAGES + 1) + 16 + 1024;
    static bool configured[mbs::kMaxDevices] = {false};
    const int dev = mbs::current_device();
    if (!configured[dev]) {
        MBS_CHECK_CUDA(cudaFuncSetAttribute(wgrad_nhwc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
        configured[dev] = true;
    }
    wgrad_nhwc_kernel<BN, STAGES><<<grid, NUM_THREADS, dyn, stream>>>(a, b, wp);
    MBS_CHECK_LAUNCH();
    return 0;
}

namespace {
struct WgradPlan {
    WgradNhwcParams wp;
    bool halo, row3;
    int bn, grid;
};
// The weight-gradient kernels hold one CTA per SM (~190 KB of operand stages): fill ONE wave when the tiles fit, otherwise
// the split (<= 4) with the fewest wave-equivalents; fewer splits = less partial-tile traffic.
int split_rule(int tiles) {
    const int sms = sm_count();
    if (tiles <= sms) return sms / tiles;
    int splits = 1;
    double best = 1e30;
    for (int sp = 1; sp <= 4; ++sp) {
        const double cost = static_cast<double>(mbs::cdiv(tiles * sp, sms)) / sp;
        if (cost < best - 1e-9) { best = cost; splits = sp; }
    }
    return splits;
}
// one place decides tiling and the K split, so that mbs_conv_wgrad_splits() and the launch always agree
int plan_wgrad(const mbs_wgrad_desc *d, WgradPlan &pl) {
    MBS_REQUIRE(d != nullptr && d->kind >= 0 && d->kind <= 2, "wgrad: bad descriptor");
    MBS_REQUIRE(d->Cm > 0 && d->Cm % 64 == 0 && d->Cn > 0 && d->Cn % 64 == 0, "wgrad: channel counts must be multiples of 64 (got %d, %d)", d->Cm, d->Cn);
    MBS_REQUIRE(d->N > 0 && d->Ho > 0 && d->Wo > 0, "wgrad: bad shape");
    WgradNhwcParams &wp = pl.wp;
    memset(&wp, 0, sizeof(wp));
    wp.N = d->N;
    wp.Ho = d->Ho;
    wp.Wo = d->Wo;
    wp.kind = d->kind;
    wp.Cm = d->Cm;
    wp.Cn = d->Cn;
    wp.out = d->out;
    wp.out_ld = d->out_ld;
    wp.out_coff = d->out_coff;
    wp.partial = d->partial ? 1 : 0;
    wp.taps = d->kind == 2 ? 4 : 9;
    wp.sA = d->kind == 2 ? 2 : 1;
    wp.sB = d->kind == 1 ? 2 : 1;
    wp.pair = ((d->kind == 0 || d->kind == 2) && d->Cm == 64) ? 1 : 0;     // two taps x 64 output channels per 128-row tile
    wp.tap_items = wp.pair ? (wp.taps + 1) / 2 : wp.taps;
    // K chunk = 64 pixels: 8x8 patches; narrow grids use taller patches
    wp.pw = d->Wo >= 8 ? 8 : (d->Wo >= 4 ? 4 : 2);
    wp.ph = 64 / wp.pw;
    wp.ptx = mbs::cdiv(d->Wo, wp.pw);
    wp.pty = mbs::cdiv(d->Ho, wp.ph);
    wp.tpi = 1;
    pl.halo = d->kind == 0 && d->Cm == 64 && d->Wo >= 8 && d->Ho >= 8 && wgrad_halo_enabled();
    pl.row3 = !pl.halo && d->kind == 0 && d->Cm % 128 == 0 && d->Cn % 128 == 0 && d->Wo >= 8 && d->Ho >= 8 && wgrad_row3_enabled();
    int splits;
    if (pl.row3) {
        // wide stride-1 layers: one filter row (three taps) per CTA from one halo patch (see wgrad_row3_kernel)
        wp.pw = wp.ph = 8;
        wp.ptx = mbs::cdiv(d->Wo, 8);
        wp.pty = mbs::cdiv(d->Ho, 8);
        pl.bn = 128;
        wp.cn_tile = 128;
        wp.m_tiles = d->Cm / 128;
        wp.n_tiles = d->Cn / 128;
        wp.tap_items = 3;
        splits = split_rule(wp.m_tiles * wp.n_tiles * 3);
    } else if (pl.halo) {
        // full-resolution layers: all nine taps per CTA from one halo patch (see wgrad_halo64_kernel)
        wp.pw = wp.ph = 8;
        wp.ptx = mbs::cdiv(d->Wo, 8);
        wp.pty = mbs::cdiv(d->Ho, 8);
        pl.bn = 64;
        wp.cn_tile = 64;
        splits = mbs::cdiv(sm_count(), d->Cn / 64);      // one resident CTA per SM: one wave, fewest partial-tile reductions
    } else {
        int bn = d->Cn % 256 == 0 ? 256 : (d->Cn % 128 == 0 ? 128 : 64);
        wp.cn_tile = bn;
        if (!wp.pair && d->kind != 2 && (d->Cn == 128 || d->Cn == 64) && wgrad_stack_enabled()) {
            // narrow inputs: stack 2 / 4 taps along a 256-column N tile (one dz tile feeds them all)
            wp.tpi = 256 / d->Cn;
            wp.cn_tile = d->Cn;
            wp.tap_items = mbs::cdiv(wp.taps, wp.tpi);
            bn = 256;
        }
        pl.bn = bn;
        wp.m_tiles = wp.pair ? 1 : mbs::cdiv(d->Cm, 128);
        wp.n_tiles = d->Cn / wp.cn_tile;
        splits = split_rule(wp.m_tiles * wp.n_tiles * wp.tap_items);
    }
    const int patches = d->N * wp.ptx * wp.pty;
    if (splits > patches) splits = patches;
    if (splits < 1) splits = 1;
    wp.per = mbs::cdiv(patches, splits);
    wp.splits = mbs::cdiv(patches, wp.per);              // every split owns at least one patch
    pl.grid = (pl.halo ? d->Cn / 64 : wp.m_tiles * wp.n_tiles * wp.tap_items) * wp.splits;     // row3: tap_items = 3 rows
    return 0;
}
}  // namespace

extern "C" int mbs_conv_wgrad_splits(const mbs_wgrad_desc *d) {
    WgradPlan pl;
    if (plan_wgrad(d, pl)) return -1;
    return pl.wp.splits;
}

extern "C" int mbs_conv_wgrad(const mbs_wgrad_desc *d, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    WgradPlan pl;
    int rc = plan_wgrad(d, pl);
    if (rc) return rc;
    MBS_REQUIRE(d->a && d->b && d->out, "wgrad: null operand");
    MBS_REQUIRE((reinterpret_cast<uintptr_t>(d->out) & 15) == 0 && (d->partial || (d->out_ld % 4 == 0 && d->out_coff % 4 == 0)),
                "wgrad: the gradient buffer must allow 16-byte vector accesses");
    const WgradNhwcParams &wp = pl.wp;
    CUtensorMap a, b;
    if (pl.row3) {
        rc = make_act_map(&a, d->a, d->N, d->Ho, d->Wo, d->Cm, d->lda, d->coffa, 1, 10, 10);
        if (rc) return rc;
        rc = make_act_map(&b, d->b, d->N, d->Ho, d->Wo, d->Cn, d->ldb, d->coffb, 1, 8, 8);
        if (rc) return rc;
        static bool configured[mbs::kMaxDevices] = {false};
        const int dev = mbs::current_device();
        if (!configured[dev]) {
            MBS_CHECK_CUDA(cudaFuncSetAttribute(wgrad_row3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W3_DYN));
            configured[dev] = true;
        }
        wgrad_row3_kernel<<<pl.grid, NUM_THREADS, W3_DYN, stream>>>(a, b, wp);
        MBS_CHECK_LAUNCH();
        return 0;
    }
    if (pl.halo) {
        rc = make_act_map(&a, d->a, d->N, d->Ho, d->Wo, d->Cm, d->lda, d->coffa, 1, 10, 10);
        if (rc) return rc;
        rc = make_act_map(&b, d->b, d->N, d->Ho, d->Wo, d->Cn, d->ldb, d->coffb, 1, 8, 8);
        if (rc) return rc;
        static bool configured[mbs::kMaxDevices] = {false};
        const int dev = mbs::current_device();
        if (!configured[dev]) {
            MBS_CHECK_CUDA(cudaFuncSetAttribute(wgrad_halo64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WH_DYN));
            configured[dev] = true;
        }
        wgrad_halo64_kernel<<<pl.grid, NUM_THREADS, WH_DYN, stream>>>(a, b, wp);
        MBS_CHECK_LAUNCH();
        return 0;
    }
    rc = make_act_map(&a, d->a, d->N, wp.sA * d->Ho, wp.sA * d->Wo, d->Cm, d->lda, d->coffa, wp.sA, wp.pw, wp.ph);
    if (rc) return rc;
    rc = make_act_map(&b, d->b, d->N, wp.sB * d->Ho, wp.sB * d->Wo, d->Cn, d->ldb, d->coffb, wp.sB, wp.pw, wp.ph);
    if (rc) return rc;
    if (pl.bn == 256) return launch_wgrad_nhwc<256, 4>(a, b, wp, pl.grid, stream);
    if (pl.bn == 128) return launch_wgrad_nhwc<128, 6>(a, b, wp, pl.grid, stream);
    return launch_wgrad_nhwc<64, 8>(a, b, wp, pl.grid, stream);
}

extern "C" int mbs_first_conv(const void *img, int in_dtype, int H, int W, int pad_y, int pad_x, float norm_lo,
                              float norm_hi, const float *lohi_dev, const float *weight, const float *bias, const float *scale,
                              const float *shift, int C, int act, void *out, int out_ld, int out_coff,
                              void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(C > 0 && C % 8 == 0 && C <= 256, "first conv: C must be a multiple of 8 and <= 256 (got %d)", C);
    MBS_REQUIRE(((reinterpret_cast<uintptr_t>(out) + static_cast<size_t>(out_coff) * 2) & 15) == 0 &&
                    (out_ld * 2) % 16 == 0,
                "first conv: destination view must be 16-byte aligned");
    const int tiles_x = mbs::cdiv(W + pad_x, FC_TW);
    const int num_tiles = tiles_x * mbs::cdiv(H + pad_y, FC_TH);
    const int fgrid = num_tiles < 2 * sm_count() ? num_tiles : 2 * sm_count();
    const int threads = 256;
    __nv_bfloat16 *o = static_cast<__nv_bfloat16 *>(out);
#define MBS_FIRST_CONV(T, RELU)                                                                                        \
    first_conv_kernel<T, RELU><<<fgrid, threads, 0, stream>>>(static_cast<const T *>(img), H, W, pad_y, pad_x, norm_lo, \
                                                              norm_hi, lohi_dev, weight, bias, scale, shift, C, act, o, \
                                                              out_ld, out_coff, tiles_x, num_tiles)
    const bool relu = act == MBS_ACT_RELU;
    switch (in_dtype) {
        case MBS_IN_U8:
            if (relu) MBS_FIRST_CONV(uint8_t, true); else MBS_FIRST_CONV(uint8_t, false);
            break;
        case MBS_IN_U16:
            if (relu) MBS_FIRST_CONV(uint16_t, true); else MBS_FIRST_CONV(uint16_t, false);
            break;
        case MBS_IN_F32:
            if (relu) MBS_FIRST_CONV(float, true); else MBS_FIRST_CONV(float, false);
            break;
        default: MBS_REQUIRE(false, "first conv: unknown input dtype %d", in_dtype);
    }
#undef MBS_FIRST_CONV
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_first_conv_halo64(const void *img, int in_dtype, int N, int H, int W, int pad_y, int pad_x, float norm_lo,
                                     float norm_hi, const float *lohi_dev, const float *weight, const float *bias,
                                     const float *scale, const float *shift, int act, const mbs_conv_desc *d, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(d != nullptr && img != nullptr, "null argument");
    MBS_REQUIRE(d->mode == MBS_CONV3X3_S1 && d->C0 == 64 && d->C1 == 0 && d->Cout == 64,
                "fused first layer: the second convolution must be 3x3 stride 1, 64 -> 64 channels");
    MBS_REQUIRE(d->N == N && d->H == H + pad_y && d->W == W + pad_x, "fused first layer: descriptor / frame size mismatch");
    MBS_REQUIRE(d->head_out == nullptr && d->dst != nullptr, "fused first layer: plain destination only");
    MBS_REQUIRE(((reinterpret_cast<uintptr_t>(d->dst) + static_cast<size_t>(d->coffd) * 2) & 15) == 0 && (d->ldd * 2) % 16 == 0,
                "destination view must be 16-byte aligned");
    ConvKParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.mode = d->mode;
    kp.act = d->act;
    kp.Hm = kp.Hd = d->H;
    kp.Wm = kp.Wd = d->W;
    kp.tiles_x = mbs::cdiv(kp.Wm, HT_W);
    kp.tiles_y = mbs::cdiv(kp.Hm, HT_H);
    kp.n_tiles = 1;
    kp.chunks0 = 1;
    kp.taps = 9;
    kp.Cout = 64;
    kp.bias = d->bias;
    kp.scale = d->scale;
    kp.shift = d->shift;
    kp.dst = static_cast<__nv_bfloat16 *>(d->dst);
    kp.ldd = d->ldd;
    kp.coffd = d->coffd;
    const long long tiles_ll = static_cast<long long>(N) * kp.tiles_x * kp.tiles_y;
    MBS_REQUIRE(tiles_ll > 0 && tiles_ll < (1ll << 31), "too many tiles");
    kp.num_tiles = static_cast<int>(tiles_ll);
    CUtensorMap b, dm;
    int rc = make_weight_map(&b, d->weight, 64, 9 * 64, 64);
    if (rc) return rc;
    rc = make_act_map(&dm, d->dst, N, kp.Hd, kp.Wd, 64, d->ldd, d->coffd, 1, HT_W, HT_H);
    if (rc) return rc;
    FirstParams fp;
    fp.img = img;
    fp.H = H;
    fp.W = W;
    fp.pad_y = pad_y;
    fp.pad_x = pad_x;
    fp.lo = norm_lo;
    fp.hi = norm_hi;
    fp.lohi_dev = lohi_dev;
    fp.weight = weight;
    fp.bias = bias;
    fp.scale = scale;
    fp.shift = shift;
    fp.act = act;
    switch (in_dtype) {
        case MBS_IN_U8:
            return act == MBS_ACT_RELU ? launch_halo_first<uint8_t, true>(b, dm, kp, fp, stream)
                                       : launch_halo_first<uint8_t, false>(b, dm, kp, fp, stream);
        case MBS_IN_U16:
            return act == MBS_ACT_RELU ? launch_halo_first<uint16_t, true>(b, dm, kp, fp, stream)
                                       : launch_halo_first<uint16_t, false>(b, dm, kp, fp, stream);
        case MBS_IN_F32:
            return act == MBS_ACT_RELU ? launch_halo_first<float, true>(b, dm, kp, fp, stream)
                                       : launch_halo_first<float, false>(b, dm, kp, fp, stream);
        default: MBS_REQUIRE(false, "fused first layer: unknown input dtype %d", in_dtype);
    }
    return 0;
}

extern "C" int mbs_debug_flags(int reset) {
    int v = 0;
    if (cudaMemcpyFromSymbol(&v, g_mbar_timeout, sizeof(int)) != cudaSuccess) return -1;
    if (reset) {
        int z = 0;
        cudaMemcpyToSymbol(g_mbar_timeout, &z, sizeof(int));
    }
    return v;
}

extern "C" int mbs_frame_minmax(const void *img, int in_dtype, long long n, float *lohi_dev, void *scratch8,
                                void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(n > 0 && lohi_dev && scratch8, "frame_minmax: bad arguments");
    unsigned int *keys = static_cast<unsigned int *>(scratch8);
    minmax_init_kernel<<<1, 1, 0, stream>>>(keys);
    MBS_CHECK_LAUNCH();
    const int blocks = static_cast<int>(n / 4096 + 1 < 1184 ? n / 4096 + 1 : 1184);
    switch (in_dtype) {
        case MBS_IN_U8: minmax_kernel<uint8_t><<<blocks, 256, 0, stream>>>(static_cast<const uint8_t *>(img), n, keys); break;
        case MBS_IN_U16: minmax_kernel<uint16_t><<<blocks, 256, 0, stream>>>(static_cast<const uint16_t *>(img), n, keys); break;
        case MBS_IN_F32: minmax_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float *>(img), n, keys); break;
        default: MBS_REQUIRE(false, "frame_minmax: unknown input dtype %d", in_dtype);
    }
    MBS_CHECK_LAUNCH();
    minmax_final_kernel<<<1, 1, 0, stream>>>(keys, in_dtype == MBS_IN_F32, lohi_dev);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_pack_conv3x3_weight(const float *w, int Cout, int Cin, void *packed, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long total = static_cast<long long>(Cout) * 9 * Cin;
    pack_conv3x3_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(
        w, Cout, Cin, static_cast<__nv_bfloat16 *>(packed));
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_pack_convT2x2_weight(const float *w, int Cin, int Cout, void *packed, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long total = static_cast<long long>(4) * Cout * Cin;
    pack_convT2x2_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(
        w, Cin, Cout, static_cast<__nv_bfloat16 *>(packed));
    MBS_CHECK_LAUNCH();
    return 0;
}
