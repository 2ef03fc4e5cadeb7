// Distance-map post-processing on the GPU, bit-exact to the reference's scipy/skimage pipeline.
//
// Replaces src/inference/postprocessing.py:7-59 (distance_postprocessing):
//   :25     gaussian_filter(cell, 0.5)           -> pp_front_kernel (separable 5-tap, f64 accumulate,
//                                                   f32 rounding after each axis, 'reflect' borders)
//   :27-37  clip / tan / thresholds              -> pp_front_kernel
//   :38     measure.label (8-connectivity)       -> union-find CCL (root = first raster pixel)
//   :41-53  regionprops areas, small-seed filter -> area histogram at roots + f64 threshold
//   :54     relabel 1..m in raster order         -> prefix scan over surviving roots
//   :57     watershed(-cell, markers, mask)      -> order-free minimax formulation (see below) with an
//                                                   exact sequential heap flood when value ties make
//                                                   the result order dependent
//   :59     astype(uint16)                       -> final store
//
// Watershed formulation.  skimage's flood pops pixels by (value, age).  The level at which pixel p
// is popped is the minimax path cost from the markers, L(p) = max(v(p), min_{q in N4(p)} L(q)) with
// L(m) = v(m) on markers.  p is labelled by its first-popped neighbour, which has minimal L among
// its neighbours; pixels sharing one L value form plateaus that are flooded from their "sources"
// (members that have a strictly lower neighbour or are markers).  If every pixel's minimal-L
// neighbours (and every plateau's sources) agree on one label the result does not depend on the
// pop order inside a level, hence equals skimage's for ANY tie-break; otherwise the image is
// "ambiguous" (needs exact value ties between competing pixels) and the whole flood is re-run by
// the sequential heap kernel, which restates skimage's algorithm including its binary heap.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cfloat>
#include <climits>
#include <cstdint>
#include <cstring>

#include "../../include/mbseg.h"
#include "common.cuh"

namespace {

// scipy.ndimage gaussian kernel for sigma=0.5, radius 2 (float64), centre / +-1 / +-2
__constant__ double c_gw[3] = {0x1.92b965ef5aaeep-1, 0x1.b405b9842b206p-4, 0x1.14aebe6a24088p-12};

__device__ __forceinline__ int reflect_idx(int i, int n) {
    // scipy 'reflect': d c b a | a b c d | d c b a
    while (i < 0 || i >= n) {
        if (i < 0) i = -i - 1;
        if (i >= n) i = 2 * n - 1 - i;
    }
    return i;
}

__device__ __forceinline__ float gauss5(float xm2, float xm1, float x0, float xp1, float xp2) {
    // NI_Correlate1D symmetric branch: tmp = x0*w0; tmp += (x[-2]+x[+2])*w2; tmp += (x[-1]+x[+1])*w1 (no FMA)
    double t = __dmul_rn(static_cast<double>(x0), c_gw[0]);
    t = __dadd_rn(t, __dmul_rn(__dadd_rn(static_cast<double>(xm2), static_cast<double>(xp2)), c_gw[2]));
    t = __dadd_rn(t, __dmul_rn(__dadd_rn(static_cast<double>(xm1), static_cast<double>(xp1)), c_gw[1]));
    return __double2float_rn(t);
}

constexpr int FT = 32;  // front-end tile edge

// One block = 32x32 output tile, 256 threads.  Phase 1: y-pass for 36 columns into smem (f32
// rounded, as scipy stores the intermediate in the float32 output array); phase 2: x-pass + maps.
__global__ void __launch_bounds__(256)
pp_front_kernel(const float *__restrict__ border, const float *__restrict__ cell, int H, int W, int ld, float th_seed,
                float th_cell, float *__restrict__ cell_s, uint8_t *__restrict__ mask, uint8_t *__restrict__ seed,
                int *__restrict__ ccl_init) {
    __shared__ float s_in[FT + 4][FT + 4];
    __shared__ float s_y[FT][FT + 4];
    const int x0 = blockIdx.x * FT, y0 = blockIdx.y * FT;
    for (int i = threadIdx.x; i < (FT + 4) * (FT + 4); i += 256) {
        const int r = i / (FT + 4), c = i % (FT + 4);
        const int yy = reflect_idx(y0 + r - 2, H), xx = reflect_idx(x0 + c - 2, W);
        s_in[r][c] = cell[static_cast<size_t>(yy) * ld + xx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < FT * (FT + 4); i += 256) {
        const int r = i / (FT + 4), c = i % (FT + 4);
        s_y[r][c] = gauss5(s_in[r][c], s_in[r + 1][c], s_in[r + 2][c], s_in[r + 3][c], s_in[r + 4][c]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < FT * FT; i += 256) {
        const int r = i / FT, c = i % FT;
        const int y = y0 + r, x = x0 + c;
        if (y >= H || x >= W) continue;
        // columns that are reflections at the image border must reflect the *y-filtered* image, which is
        // what s_y holds because the y-pass was evaluated at the reflected source column.
        const float cs = gauss5(s_y[r][c], s_y[r][c + 1], s_y[r][c + 2], s_y[r][c + 3], s_y[r][c + 4]);
        float b = border[static_cast<size_t>(y) * ld + x];
        b = b < 0.0f ? 0.0f : (b > 1.0f ? 1.0f : b);                    // np.clip keeps NaN
        const float sq = __fmul_rn(b, b);
        float t = static_cast<float>(tan(static_cast<double>(sq)));     // float32(tan_f64(x))
        if (t < 0.05f) t = 0.0f;
        t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
        const float cleaned = __fsub_rn(cs, t);
        const size_t o = static_cast<size_t>(y) * W + x;
        cell_s[o] = cs;
        mask[o] = cs > th_cell ? 1 : 0;
        const bool sd = cleaned > th_seed;
        seed[o] = sd ? 1 : 0;
        if (ccl_init) ccl_init[o] = sd ? static_cast<int>(o) : -1;     // union-find initial state (saves a pass)
    }
}

// boundary_postprocessing front end (postprocessing.py:71-77): argmax over the 3 classes (first maximum wins,
// as np.argmax), mask = class 1, seeds = p1 * (1 - p2) > 0.5; the flood image is the mask itself.
__global__ void bp_front_kernel(const float *__restrict__ pred, int n, float *__restrict__ img, uint8_t *__restrict__ mask,
                                uint8_t *__restrict__ seed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p0 = pred[3 * i], p1 = pred[3 * i + 1], p2 = pred[3 * i + 2];
    int am = 0;
    float best = p0;
    if (p1 > best) { best = p1; am = 1; }
    if (p2 > best) { am = 2; }
    const bool m = am == 1;
    mask[i] = m;
    img[i] = m ? 1.0f : 0.0f;
    seed[i] = __fmul_rn(p1, __fsub_rn(1.0f, p2)) > 0.5f;
}

// ------------------------------------------------------------------------------------------
// union-find (root = minimum linear index = first pixel in raster order)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int *L, int a) {
    int p = L[a];
    while (p != a) {
        a = p;
        p = L[a];
    }
    return a;
}
__device__ __forceinline__ int uf_find_v(volatile int *L, int a) {
    int p = L[a];
    while (p != a) {
        a = p;
        p = L[a];
    }
    return a;
}
__device__ __forceinline__ void uf_union(int *L, int a, int b) {
    bool done;
    do {
        a = uf_find_v(L, a);
        b = uf_find_v(L, b);
        if (a < b) {
            int old = atomicMin(&L[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            int old = atomicMin(&L[a], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

__global__ void ccl_init_kernel(const uint8_t *__restrict__ fg, int n, int *__restrict__ L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) L[i] = fg[i] ? i : -1;
}
// 8-connectivity.  A pixel only issues the unions that are not implied by its left neighbour's:
//   W  : always (cheap: the roots usually coincide already after the first pixels of a run)
//   N  : unless both left neighbours (W and NW) are foreground -- then pixel W already joined the two runs
//   NW : only if N is background and W is background (else W's own N union covers it)
//   NE : only if N is background and E is background (else E's N union covers it)
__global__ void ccl_merge8_kernel(const uint8_t *__restrict__ fg, int H, int W, int *L) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int i = y * W + x;
    if (!fg[i]) return;
    const bool w = x > 0 && fg[i - 1];
    if (w) uf_union(L, i, i - 1);
    if (y > 0) {
        const int u = i - W;
        const bool n = fg[u] != 0;
        const bool nw = x > 0 && fg[u - 1];
        if (n) {
            if (!(w && nw)) uf_union(L, i, u);
        } else {
            if (nw && !w) uf_union(L, i, u - 1);
            if (x + 1 < W && fg[u + 1] && !fg[i + 1]) uf_union(L, i, u + 1);
        }
    }
}
// measure.label on an INTEGER image (eval.py:261,313 label the ground truth and the prediction before AJI+):
// 8-connected pixels belong to one component iff they carry the same non-zero value.
__global__ void ccl_init_values_kernel(const uint16_t *__restrict__ v, int n, int *__restrict__ L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) L[i] = v[i] ? i : -1;
}
__global__ void ccl_merge8_values_kernel(const uint16_t *__restrict__ v, int H, int W, int *L) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int i = y * W + x;
    const uint16_t me = v[i];
    if (!me) return;
    if (x > 0 && v[i - 1] == me) uf_union(L, i, i - 1);
    if (y > 0) {
        const int u = i - W;
        if (v[u] == me) uf_union(L, i, u);
        if (x > 0 && v[u - 1] == me) uf_union(L, i, u - 1);
        if (x + 1 < W && v[u + 1] == me) uf_union(L, i, u + 1);
    }
}
__global__ void ccl_compress_kernel(int n, int *L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && L[i] >= 0) L[i] = uf_find(L, i);
}

struct Stats {
    unsigned long long total;   // seed pixels
    unsigned int n_comp;        // components before filtering
    unsigned int n_markers;     // components after filtering
    unsigned int changed[3];    // rotating "some tile changed in sweep s" flags (slot s % 3)
    unsigned int ambiguous;     // ambiguity counter
    unsigned int sweeps;
};

// path compression + component areas + component / pixel counts in one pass
__global__ void compress_area_kernel(int *L, int n, int *__restrict__ area, Stats *st) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int r = -1;
    if (i < n && L[i] >= 0) {
        r = uf_find(L, i);
        L[i] = r;
    }
    // all pixels of a component add to the same counter: group the warp's lanes by root, one atomic per group
    unsigned remaining = __ballot_sync(0xffffffffu, r >= 0);
    const int lane = threadIdx.x & 31;
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const int cur = __shfl_sync(0xffffffffu, r, leader);
        const unsigned grp = __ballot_sync(0xffffffffu, r == cur);
        if (lane == leader) atomicAdd(&area[cur], __popc(grp));
        remaining &= ~grp;
    }
    const int cnt = __syncthreads_count(r >= 0);
    const int roots = __syncthreads_count(r >= 0 && r == i);
    if (threadIdx.x == 0) {
        if (cnt) atomicAdd(&st->total, static_cast<unsigned long long>(cnt));
        if (roots) atomicAdd(&st->n_comp, static_cast<unsigned int>(roots));
    }
}

__device__ __forceinline__ bool keep_root(int area, const Stats *st, int use_mean) {
    // postprocessing.py:46-53: min_area = max(0.10 * mean(areas), 4); drop area <= min_area (float64)
    double min_area = 0.0;
    if (use_mean && st->n_comp > 0)
        min_area = __dmul_rn(0.10, __ddiv_rn(static_cast<double>(st->total), static_cast<double>(st->n_comp)));
    min_area = fmax(min_area, 4.0);
    return !(static_cast<double>(area) <= min_area);
}

constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_TILE = SCAN_ITEMS * SCAN_THREADS;

// pass 1: number of surviving roots per 4096-pixel tile
__global__ void __launch_bounds__(SCAN_THREADS)
rank_count_kernel(const int *__restrict__ L, const int *__restrict__ area, int n, const Stats *st, int use_mean,
                  int *__restrict__ tile_count) {
    const int base = blockIdx.x * SCAN_TILE;
    int c = 0;
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int i = base + k * SCAN_THREADS + threadIdx.x;
        if (i < n && L[i] == i && keep_root(area[i], st, use_mean)) ++c;
    }
    __shared__ int s_red[SCAN_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += s_red[w];
        tile_count[blockIdx.x] = t;
    }
}
// pass 2: exclusive scan of tile counts (single block)
__global__ void __launch_bounds__(1024) rank_scan_kernel(int *tile_count, int n_tiles, Stats *st) {
    __shared__ int s_part[1024];
    const int per = (n_tiles + 1023) / 1024;
    const int b = threadIdx.x * per;
    int sum = 0;
    for (int k = 0; k < per; ++k)
        if (b + k < n_tiles) sum += tile_count[b + k];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (int o = 1; o < 1024; o <<= 1) {
        int v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = threadIdx.x ? s_part[threadIdx.x - 1] : 0;
    for (int k = 0; k < per; ++k)
        if (b + k < n_tiles) {
            const int c = tile_count[b + k];
            tile_count[b + k] = run;
            run += c;
        }
    if (threadIdx.x == 1023) st->n_markers = static_cast<unsigned int>(s_part[1023]);
}
// pass 3: rank of every surviving root (1-based, raster order), 0 otherwise
__global__ void __launch_bounds__(SCAN_THREADS)
rank_assign_kernel(const int *__restrict__ L, const int *__restrict__ area, int n, const Stats *st, int use_mean,
                   const int *__restrict__ tile_off, int *__restrict__ rank) {
    // thread t owns the contiguous items [t*16, t*16+16) of the tile so that ranks follow raster order
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned flags = 0;
    int c = 0;
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int i = base + k;
        if (i < n && L[i] == i && keep_root(area[i], st, use_mean)) {
            flags |= 1u << k;
            ++c;
        }
    }
    // block exclusive scan of c
    __shared__ int s_w[SCAN_THREADS / 32];
    int inc = c;
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += v;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) woff += s_w[w];
    int run = tile_off[blockIdx.x] + woff + inc - c;
    // ranks are only ever read at component roots: write the survivors' ranks and zero the dropped roots
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int i = base + k;
        if (i < n && L[i] == i) rank[i] = (flags >> k) & 1u ? ++run : 0;
    }
}

// markers = relabelled seeds * mask (_validate_inputs) and the initial flood level (markers: v, else +inf)
__global__ void markers_kernel(const int *__restrict__ L, const int *__restrict__ rank, const uint8_t *__restrict__ mask,
                               int n, int *__restrict__ markers, const float *__restrict__ img, int negate,
                               float *__restrict__ Lv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = L[i];
    const int m = (r >= 0 && mask[i]) ? rank[r] : 0;
    markers[i] = m;
    if (Lv) {
        const float v = img[i];
        Lv[i] = m > 0 ? (negate ? -v : v) : __builtin_huge_valf();
    }
}
__global__ void labels_from_roots_kernel(const int *__restrict__ L, const int *__restrict__ rank, int n,
                                         int *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = L[i];
    out[i] = r >= 0 ? rank[r] : 0;
}

// ------------------------------------------------------------------------------------------
// watershed: minimax relaxation
// ------------------------------------------------------------------------------------------
constexpr float kInf = __builtin_huge_valf();
constexpr int RT = 32;  // relaxation tile edge

// image value used by the flood.  negate != 0 -> v = -img (the reference floods -cell)
__device__ __forceinline__ float flood_value(const float *img, int i, int negate) {
    const float v = img[i];
    return negate ? -v : v;
}

__global__ void ws_init_kernel(const float *__restrict__ img, int negate, const int *__restrict__ markers,
                               const uint8_t *__restrict__ mask, int n, float *__restrict__ Lv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Lv[i] = (mask[i] && markers[i] > 0) ? flood_value(img, i, negate) : kInf;
}

// Minimax relaxation, one cooperative launch for all sweeps (no host round trip).  A block relaxes
// 32x32 tiles (1-pixel halo) to their local fixed point in shared memory; sweeps are separated by a
// grid-wide barrier and repeat until no tile changed.  A tile is revisited only if it or one of its
// four neighbours changed in the previous sweep.
__global__ void __launch_bounds__(256)
ws_relax_coop_kernel(const float *__restrict__ img, int negate, const int *__restrict__ markers,
                     const uint8_t *__restrict__ mask, int H, int W, float *Lv, Stats *st, uint8_t *tile_changed) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ float sL[RT + 2][RT + 2];
    const int tiles_x = (W + RT - 1) / RT, tiles_y = (H + RT - 1) / RT;
    const int ntiles = tiles_x * tiles_y;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    int sweep = 0;
    for (;;) {
        const uint8_t *prev = tile_changed + ((sweep + 1) & 1) * ntiles;
        uint8_t *cur = tile_changed + (sweep & 1) * ntiles;
        bool block_changed = false;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int tyi = tile / tiles_x, txi = tile - tyi * tiles_x;
            if (sweep > 0) {
                const bool need = prev[tile] || (txi > 0 && prev[tile - 1]) || (txi + 1 < tiles_x && prev[tile + 1]) ||
                                  (tyi > 0 && prev[tile - tiles_x]) || (tyi + 1 < tiles_y && prev[tile + tiles_x]);
                if (!need) {
                    if (threadIdx.x == 0) cur[tile] = 0;
                    continue;
                }
            }
            const int x0 = txi * RT, y0 = tyi * RT;
            __syncthreads();   // sL reuse across tiles
            for (int i = threadIdx.x; i < (RT + 2) * (RT + 2); i += 256) {
                const int r = i / (RT + 2), c = i % (RT + 2);
                const int y = y0 + r - 1, x = x0 + c - 1;
                sL[r][c] = (y >= 0 && y < H && x >= 0 && x < W) ? Lv[static_cast<size_t>(y) * W + x] : kInf;
            }
            float v[4];
            bool act[4];
            bool any_act = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int y = y0 + ty * 4 + k, x = x0 + tx;
                act[k] = false;
                v[k] = kInf;
                if (y < H && x < W) {
                    const size_t i = static_cast<size_t>(y) * W + x;
                    if (mask[i] && markers[i] == 0) {
                        act[k] = true;
                        v[k] = flood_value(img, static_cast<int>(i), negate);
                    }
                }
                any_act |= act[k];
            }
            bool changed_any = false;
            if (__syncthreads_or(any_act)) {
                for (;;) {
                    bool changed = false;
                    // down then up over this thread's 4-pixel column strip: levels travel the whole strip per iteration
#pragma unroll
                    for (int kk = 0; kk < 7; ++kk) {
                        const int k = kk < 4 ? kk : 6 - kk;
                        if (!act[k]) continue;
                        const int r = ty * 4 + k + 1, c = tx + 1;
                        const float m = fminf(fminf(sL[r - 1][c], sL[r + 1][c]), fminf(sL[r][c - 1], sL[r][c + 1]));
                        const float nl = fmaxf(v[k], m);
                        if (nl < sL[r][c]) {
                            sL[r][c] = nl;
                            changed = true;
                        }
                    }
                    changed_any |= changed;
                    if (!__syncthreads_or(changed)) break;
                }
                if (changed_any) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (act[k]) Lv[static_cast<size_t>(y0 + ty * 4 + k) * W + x0 + tx] = sL[ty * 4 + k + 1][tx + 1];
                }
            }
            const bool tile_ch = __syncthreads_or(changed_any) != 0;
            if (threadIdx.x == 0) cur[tile] = tile_ch ? 1 : 0;
            block_changed |= tile_ch;
        }
        if (block_changed && threadIdx.x == 0) st->changed[sweep % 3] = 1;
        __threadfence();
        grid.sync();
        const unsigned int flag = *reinterpret_cast<volatile unsigned int *>(&st->changed[sweep % 3]);
        if (blockIdx.x == 0 && threadIdx.x == 0) st->changed[(sweep + 2) % 3] = 0;
        ++sweep;
        if (!flag) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) st->sweeps = static_cast<unsigned int>(sweep);
}

// parent codes
constexpr int P_NONE = -1;      // not flooded (outside mask or unreachable)
constexpr int P_PLATEAU = -2;   // plateau member without a strictly lower neighbour (resolved later)

// classify every pixel: marker -> parent = self; strictly lower neighbour -> parent = first minimal-L
// neighbour in skimage's neighbour order (up, left, right, down); else plateau member.
__global__ void ws_parent_kernel(const float *__restrict__ Lv, const int *__restrict__ markers,
                                 const uint8_t *__restrict__ mask, int H, int W, int *__restrict__ parent,
                                 int *__restrict__ uf, int *__restrict__ src) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int i = y * W + x;
    src[i] = INT_MAX;
    uf[i] = i;
    const float l = Lv[i];
    if (!mask[i] || l == kInf) {
        parent[i] = P_NONE;
        return;
    }
    if (markers[i] > 0) {
        parent[i] = i;
        return;
    }
    float m = kInf;
    int q = -1;
    if (y > 0 && Lv[i - W] < m) { m = Lv[i - W]; q = i - W; }
    if (x > 0 && Lv[i - 1] < m) { m = Lv[i - 1]; q = i - 1; }
    if (x + 1 < W && Lv[i + 1] < m) { m = Lv[i + 1]; q = i + 1; }
    if (y + 1 < H && Lv[i + W] < m) { m = Lv[i + W]; q = i + W; }
    parent[i] = (m < l) ? q : P_PLATEAU;
}
// unite adjacent unresolved plateau members of equal level
__global__ void ws_plateau_union_kernel(const float *__restrict__ Lv, const int *__restrict__ parent, int H, int W,
                                        int *uf) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int i = y * W + x;
    if (parent[i] != P_PLATEAU) return;
    const float l = Lv[i];
    if (x > 0 && parent[i - 1] == P_PLATEAU && Lv[i - 1] == l) uf_union(uf, i, i - 1);
    if (y > 0 && parent[i - W] == P_PLATEAU && Lv[i - W] == l) uf_union(uf, i, i - W);
}
// every plateau component picks the smallest-index adjacent resolved pixel of the same level as its source
__global__ void ws_plateau_source_kernel(const float *__restrict__ Lv, const int *__restrict__ parent, int H, int W,
                                         const int *uf, int *src) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int i = y * W + x;
    if (parent[i] != P_PLATEAU) return;
    const float l = Lv[i];
    int best = INT_MAX;
    if (y > 0 && parent[i - W] >= 0 && Lv[i - W] == l) best = min(best, i - W);
    if (x > 0 && parent[i - 1] >= 0 && Lv[i - 1] == l) best = min(best, i - 1);
    if (x + 1 < W && parent[i + 1] >= 0 && Lv[i + 1] == l) best = min(best, i + 1);
    if (y + 1 < H && parent[i + W] >= 0 && Lv[i + W] == l) best = min(best, i + W);
    if (best != INT_MAX) atomicMin(&src[uf_find(uf, i)], best);
}
// resolve labels by chasing parent pointers (plateau members hop to their component's source)
__global__ void ws_label_kernel(const int *__restrict__ parent, const int *__restrict__ uf, const int *__restrict__ src,
                                const int *__restrict__ markers, int n, int *__restrict__ lab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = i;
    int out = 0;
    for (int guard = 0; guard < (1 << 26); ++guard) {
        const int q = parent[p];
        if (q == P_NONE) break;
        if (q == p) {
            out = markers[p];
            break;
        }
        if (q == P_PLATEAU) {
            const int s = src[uf_find(uf, p)];
            if (s == INT_MAX) break;
            p = s;
        } else {
            p = q;
        }
    }
    lab[i] = out;
}
// order-independence check (see file header)
__global__ void ws_check_kernel(const float *__restrict__ Lv, const int *__restrict__ parent,
                                const int *__restrict__ lab, int H, int W, Stats *st, uint16_t *__restrict__ out16) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    bool bad = false;
    if (x < W && y < H) {
        const int i = y * W + x;
        const int pr = parent[i];
        if (out16) out16[i] = static_cast<uint16_t>(static_cast<unsigned int>(lab[i]));   // astype(uint16) wraps
        if (pr != P_NONE && pr != i) {
            const int me = lab[i];
            // level of the neighbours that may have labelled this pixel
            const float lref = (pr == P_PLATEAU) ? Lv[i] : Lv[pr];
            if (y > 0 && parent[i - W] != P_NONE && Lv[i - W] == lref && lab[i - W] != me) bad = true;
            if (x > 0 && parent[i - 1] != P_NONE && Lv[i - 1] == lref && lab[i - 1] != me) bad = true;
            if (x + 1 < W && parent[i + 1] != P_NONE && Lv[i + 1] == lref && lab[i + 1] != me) bad = true;
            if (y + 1 < H && parent[i + W] != P_NONE && Lv[i + W] == lref && lab[i + W] != me) bad = true;
        }
    }
    const int c = __syncthreads_count(bad);
    if (c && threadIdx.x == 0 && threadIdx.y == 0) atomicAdd(&st->ambiguous, static_cast<unsigned int>(c));
}

// ------------------------------------------------------------------------------------------
// exact sequential flood (one thread): restatement of skimage's watershed_raveled + binary heap.
// Only runs when the order-free result is ambiguous (or when a test forces it).
// ------------------------------------------------------------------------------------------
struct HeapItem {
    float value;   // float32 image values compare exactly like their float64 promotions
    int age;
    int index;
};
__device__ __forceinline__ bool item_smaller(const HeapItem &a, const HeapItem &b) {
    if (a.value != b.value) return a.value < b.value;
    return a.age < b.age;
}
__device__ void heap_push(HeapItem *h, int &n, const HeapItem &e) {
    int child = n++;
    h[child] = e;
    while (child > 0) {
        const int par = (child + 1) / 2 - 1;
        if (item_smaller(h[child], h[par])) {
            HeapItem t = h[par];
            h[par] = h[child];
            h[child] = t;
            child = par;
        } else {
            break;
        }
    }
}
__device__ void heap_pop(HeapItem *h, int &n, HeapItem &dst) {
    dst = h[0];
    n -= 1;
    if (n == 0) return;
    h[0] = h[n];
    int i = 0, smallest = 0;
    for (;;) {
        const int l = 2 * i + 1, r = 2 * i + 2;
        if (l < n) {
            if (item_smaller(h[l], h[i])) smallest = l;
            if (r < n && item_smaller(h[r], h[smallest])) smallest = r;
        } else {
            break;
        }
        if (smallest == i) break;
        HeapItem t = h[i];
        h[i] = h[smallest];
        h[smallest] = t;
        i = smallest;
    }
}
__global__ void ws_sequential_kernel(const float *__restrict__ img, int negate, const int *__restrict__ markers,
                                     const uint8_t *__restrict__ mask, int H, int W, int *lab, HeapItem *heap,
                                     const Stats *st, int force, uint16_t *out16) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (!force && st->ambiguous == 0) return;
    const int n = H * W;
    int hn = 0;
    for (int i = 0; i < n; ++i) {
        const int m = mask[i] ? markers[i] : 0;
        lab[i] = m;
        if (m) {
            HeapItem e;
            e.value = flood_value(img, i, negate);
            e.age = 0;
            e.index = i;
            heap_push(heap, hn, e);
        }
    }
    int age = 1;
    HeapItem e, ne;
    while (hn > 0) {
        heap_pop(heap, hn, e);
        const int y = e.index / W, x = e.index - y * W;
        const int l = lab[e.index];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int q;
            if (k == 0) { if (y == 0) continue; q = e.index - W; }
            else if (k == 1) { if (x == 0) continue; q = e.index - 1; }
            else if (k == 2) { if (x + 1 >= W) continue; q = e.index + 1; }
            else { if (y + 1 >= H) continue; q = e.index + W; }
            if (!mask[q]) continue;
            if (lab[q]) continue;
            age += 1;
            lab[q] = l;
            ne.value = flood_value(img, q, negate);
            ne.age = age;
            ne.index = q;
            heap_push(heap, hn, ne);
        }
    }
    if (out16)
        for (int i = 0; i < n; ++i) out16[i] = static_cast<uint16_t>(static_cast<unsigned int>(lab[i]));
}

// ------------------------------------------------------------------------------------------
// workspace carving
// ------------------------------------------------------------------------------------------
struct Carver {
    char *p;
    size_t left;
    bool ok = true;
    template <typename T>
    T *take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~static_cast<size_t>(255);
        if (bytes > left) {
            ok = false;
            return nullptr;
        }
        T *r = reinterpret_cast<T *>(p);
        p += bytes;
        left -= bytes;
        return r;
    }
};
inline size_t r256(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }

size_t label_ws_bytes(size_t n) {
    const size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    return r256(sizeof(Stats)) + r256(n * 4) /*area*/ + r256(n * 4) /*rank*/ + r256(tiles * 4) + r256(n * 4) /*roots*/;
}
size_t watershed_ws_bytes(size_t n) {
    return r256(sizeof(Stats)) + 4 * r256(n * 4) /*Lv,parent,uf,src*/ + r256(n * sizeof(HeapItem)) + r256(n / 64 + 4096);
}

// label + filter + rank.  roots: [n] int scratch/out (component root per pixel, -1 background)
int run_label_rank(const uint8_t *fg, int H, int W, int *roots, int *area, int *rank, int *tile_cnt, Stats *st,
                   int use_mean, int roots_initialised, cudaStream_t stream) {
    const int n = H * W;
    const int nb = mbs::cdiv(n, 256);
    dim3 b2(32, 8), g2(mbs::cdiv(W, 32), mbs::cdiv(H, 8));
    if (!roots_initialised) {
        ccl_init_kernel<<<nb, 256, 0, stream>>>(fg, n, roots);
        MBS_CHECK_LAUNCH();
    }
    ccl_merge8_kernel<<<g2, b2, 0, stream>>>(fg, H, W, roots);
    MBS_CHECK_LAUNCH();
    MBS_CHECK_CUDA(cudaMemsetAsync(area, 0, static_cast<size_t>(n) * 4, stream));
    compress_area_kernel<<<nb, 256, 0, stream>>>(roots, n, area, st);
    MBS_CHECK_LAUNCH();
    const int tiles = mbs::cdiv(n, SCAN_TILE);
    rank_count_kernel<<<tiles, SCAN_THREADS, 0, stream>>>(roots, area, n, st, use_mean, tile_cnt);
    MBS_CHECK_LAUNCH();
    rank_scan_kernel<<<1, 1024, 0, stream>>>(tile_cnt, tiles, st);
    MBS_CHECK_LAUNCH();
    rank_assign_kernel<<<tiles, SCAN_THREADS, 0, stream>>>(roots, area, n, st, use_mean, tile_cnt, rank);
    MBS_CHECK_LAUNCH();
    return 0;
}

int run_watershed(const float *img, int negate, const int *markers, const uint8_t *mask, int H, int W, int *lab,
                  float *Lv, int *parent, int *uf, int *src, HeapItem *heap, Stats *st, uint8_t *tile_changed,
                  int force_sequential, int lv_initialised, uint16_t *out16, cudaStream_t stream) {
    const int n = H * W;
    const int nb = mbs::cdiv(n, 256);
    dim3 b2(32, 8), g2(mbs::cdiv(W, 32), mbs::cdiv(H, 8));
    dim3 gr(mbs::cdiv(W, RT), mbs::cdiv(H, RT));
    if (!force_sequential) {
        if (!lv_initialised) {
            ws_init_kernel<<<nb, 256, 0, stream>>>(img, negate, markers, mask, n, Lv);
            MBS_CHECK_LAUNCH();
        }
        // all sweeps in one cooperative launch; convergence is detected on the device
        {
            static int coop_blocks = 0;
            if (coop_blocks == 0) {
                int dev = 0, sms = 0, per_sm = 0;
                MBS_CHECK_CUDA(cudaGetDevice(&dev));
                MBS_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
                MBS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ws_relax_coop_kernel, 256, 0));
                coop_blocks = sms * (per_sm > 0 ? per_sm : 1);
            }
            const int ntiles = static_cast<int>(gr.x * gr.y);
            int blocks = ntiles < coop_blocks ? ntiles : coop_blocks;
            MBS_CHECK_CUDA(cudaMemsetAsync(tile_changed, 0, 2 * static_cast<size_t>(ntiles), stream));
            void *args[] = {(void *)&img, (void *)&negate, (void *)&markers, (void *)&mask, (void *)&H, (void *)&W,
                            (void *)&Lv, (void *)&st, (void *)&tile_changed};
            MBS_CHECK_CUDA(cudaLaunchCooperativeKernel((void *)ws_relax_coop_kernel, dim3(blocks), dim3(256), args, 0, stream));
            mbs::count_launch();
        }
        ws_parent_kernel<<<g2, b2, 0, stream>>>(Lv, markers, mask, H, W, parent, uf, src);
        MBS_CHECK_LAUNCH();
        ws_plateau_union_kernel<<<g2, b2, 0, stream>>>(Lv, parent, H, W, uf);
        MBS_CHECK_LAUNCH();
        ws_plateau_source_kernel<<<g2, b2, 0, stream>>>(Lv, parent, H, W, uf, src);
        MBS_CHECK_LAUNCH();
        ws_label_kernel<<<nb, 256, 0, stream>>>(parent, uf, src, markers, n, lab);
        MBS_CHECK_LAUNCH();
        ws_check_kernel<<<g2, b2, 0, stream>>>(Lv, parent, lab, H, W, st, out16);
        MBS_CHECK_LAUNCH();
    }
    ws_sequential_kernel<<<1, 32, 0, stream>>>(img, negate, markers, mask, H, W, lab, heap, st, force_sequential, out16);
    MBS_CHECK_LAUNCH();
    return 0;
}

Stats *pinned_stats() {
    static thread_local Stats *p = nullptr;
    if (!p) {
        if (cudaMallocHost(&p, sizeof(Stats)) != cudaSuccess) p = nullptr;
    }
    return p;
}

}  // namespace

extern "C" size_t mbs_postproc_workspace_bytes(int H, int W) {
    const size_t n = static_cast<size_t>(H) * W;
    const size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    // cell_s, mask, seed, roots, area, rank, tile counts, markers, Lv, parent, uf, src, lab, heap, stats
    return r256(n * 4) + 2 * r256(n) + 3 * r256(n * 4) + r256(tiles * 4) + r256(n * 4) + 4 * r256(n * 4) +
           r256(n * 4) + r256(n * sizeof(HeapItem)) + r256(sizeof(Stats)) + r256(n / 64 + 4096) + 4096;
}

extern "C" int mbs_pp_front(const float *border, const float *cell, int H, int W, int ld, float th_seed, float th_cell,
                            float *cell_smooth, uint8_t *mask, uint8_t *seed, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(H > 0 && W > 0 && ld >= W, "pp_front: bad shape H=%d W=%d ld=%d", H, W, ld);
    dim3 grid(mbs::cdiv(W, FT), mbs::cdiv(H, FT));
    pp_front_kernel<<<grid, 256, 0, stream>>>(border, cell, H, W, ld, th_seed, th_cell, cell_smooth, mask, seed, nullptr);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_pp_label8(const uint8_t *binary, int H, int W, int32_t *labels, int32_t *n_out, void *workspace,
                             size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t n = static_cast<size_t>(H) * W;
    MBS_REQUIRE(H > 0 && W > 0 && n < (1ull << 31), "label8: bad shape");
    MBS_REQUIRE(workspace_bytes >= label_ws_bytes(n), "label8: workspace too small (%zu < %zu)", workspace_bytes,
                label_ws_bytes(n));
    Carver cv{static_cast<char *>(workspace), workspace_bytes};
    Stats *st = cv.take<Stats>(1);
    int *area = cv.take<int>(n);
    int *rank = cv.take<int>(n);
    int *tiles = cv.take<int>((n + SCAN_TILE - 1) / SCAN_TILE);
    int *roots = cv.take<int>(n);
    MBS_REQUIRE(cv.ok, "label8: workspace carve failed");
    MBS_CHECK_CUDA(cudaMemsetAsync(st, 0, sizeof(Stats), stream));
    // plain labelling == filter that keeps everything: use_mean = 0 and area floor handled by caller,
    // here we want *all* components, so rank with a threshold that never drops: emulate by marking
    // areas as huge is unnecessary -- keep_root drops area <= 4, so label8 ranks via a dedicated path:
    const int nn = static_cast<int>(n);
    const int nb = mbs::cdiv(nn, 256);
    dim3 b2(32, 8), g2(mbs::cdiv(W, 32), mbs::cdiv(H, 8));
    ccl_init_kernel<<<nb, 256, 0, stream>>>(binary, nn, roots);
    MBS_CHECK_LAUNCH();
    ccl_merge8_kernel<<<g2, b2, 0, stream>>>(binary, H, W, roots);
    MBS_CHECK_LAUNCH();
    ccl_compress_kernel<<<nb, 256, 0, stream>>>(nn, roots);
    MBS_CHECK_LAUNCH();
    // areas := large constant at roots so that every component survives keep_root()
    MBS_CHECK_CUDA(cudaMemsetAsync(area, 0x3f, n * 4, stream));
    const int tl = mbs::cdiv(nn, SCAN_TILE);
    rank_count_kernel<<<tl, SCAN_THREADS, 0, stream>>>(roots, area, nn, st, 0, tiles);
    MBS_CHECK_LAUNCH();
    rank_scan_kernel<<<1, 1024, 0, stream>>>(tiles, tl, st);
    MBS_CHECK_LAUNCH();
    rank_assign_kernel<<<tl, SCAN_THREADS, 0, stream>>>(roots, area, nn, st, 0, tiles, rank);
    MBS_CHECK_LAUNCH();
    labels_from_roots_kernel<<<nb, 256, 0, stream>>>(roots, rank, nn, labels);
    MBS_CHECK_LAUNCH();
    if (n_out) {
        MBS_CHECK_CUDA(cudaMemcpyAsync(n_out, &st->n_markers, sizeof(int), cudaMemcpyDeviceToDevice, stream));
    }
    return 0;
}

extern "C" int mbs_label8_instances(const uint16_t *image, int H, int W, int32_t *labels, int32_t *n_out, void *workspace,
                                    size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t n = static_cast<size_t>(H) * W;
    MBS_REQUIRE(H > 0 && W > 0 && n < (1ull << 31), "label8_instances: bad shape");
    MBS_REQUIRE(workspace_bytes >= label_ws_bytes(n), "label8_instances: workspace too small (%zu < %zu)", workspace_bytes,
                label_ws_bytes(n));
    Carver cv{static_cast<char *>(workspace), workspace_bytes};
    Stats *st = cv.take<Stats>(1);
    int *area = cv.take<int>(n);
    int *rank = cv.take<int>(n);
    int *tiles = cv.take<int>((n + SCAN_TILE - 1) / SCAN_TILE);
    int *roots = cv.take<int>(n);
    MBS_REQUIRE(cv.ok, "label8_instances: workspace carve failed");
    MBS_CHECK_CUDA(cudaMemsetAsync(st, 0, sizeof(Stats), stream));
    const int nn = static_cast<int>(n);
    const int nb = mbs::cdiv(nn, 256);
    dim3 b2(32, 8), g2(mbs::cdiv(W, 32), mbs::cdiv(H, 8));
    ccl_init_values_kernel<<<nb, 256, 0, stream>>>(image, nn, roots);
    MBS_CHECK_LAUNCH();
    ccl_merge8_values_kernel<<<g2, b2, 0, stream>>>(image, H, W, roots);
    MBS_CHECK_LAUNCH();
    ccl_compress_kernel<<<nb, 256, 0, stream>>>(nn, roots);
    MBS_CHECK_LAUNCH();
    MBS_CHECK_CUDA(cudaMemsetAsync(area, 0x3f, n * 4, stream));      // every component survives the area filter
    const int tl = mbs::cdiv(nn, SCAN_TILE);
    rank_count_kernel<<<tl, SCAN_THREADS, 0, stream>>>(roots, area, nn, st, 0, tiles);
    MBS_CHECK_LAUNCH();
    rank_scan_kernel<<<1, 1024, 0, stream>>>(tiles, tl, st);
    MBS_CHECK_LAUNCH();
    rank_assign_kernel<<<tl, SCAN_THREADS, 0, stream>>>(roots, area, nn, st, 0, tiles, rank);
    MBS_CHECK_LAUNCH();
    labels_from_roots_kernel<<<nb, 256, 0, stream>>>(roots, rank, nn, labels);
    MBS_CHECK_LAUNCH();
    if (n_out) {
        MBS_CHECK_CUDA(cudaMemcpyAsync(n_out, &st->n_markers, sizeof(int), cudaMemcpyDeviceToDevice, stream));
    }
    return 0;
}

extern "C" int mbs_pp_watershed(const float *image, const int32_t *markers, const uint8_t *mask, int H, int W,
                                int32_t *labels_out, void *workspace, size_t workspace_bytes, int64_t *info_host,
                                int force_sequential, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t n = static_cast<size_t>(H) * W;
    MBS_REQUIRE(H > 0 && W > 0 && n < (1ull << 31), "watershed: bad shape");
    MBS_REQUIRE(workspace_bytes >= watershed_ws_bytes(n), "watershed: workspace too small (%zu < %zu)",
                workspace_bytes, watershed_ws_bytes(n));
    Carver cv{static_cast<char *>(workspace), workspace_bytes};
    Stats *st = cv.take<Stats>(1);
    float *Lv = cv.take<float>(n);
    int *parent = cv.take<int>(n);
    int *uf = cv.take<int>(n);
    int *src = cv.take<int>(n);
    HeapItem *heap = cv.take<HeapItem>(n);
    uint8_t *tile_changed = cv.take<uint8_t>(n / 64 + 4096);
    MBS_REQUIRE(cv.ok, "watershed: workspace carve failed");
    Stats *hp = pinned_stats();
    MBS_REQUIRE(hp != nullptr, "watershed: cannot allocate pinned host memory");
    MBS_CHECK_CUDA(cudaMemsetAsync(st, 0, sizeof(Stats), stream));
    int rc = run_watershed(image, 0, markers, mask, H, W, labels_out, Lv, parent, uf, src, heap, st, tile_changed,
                           force_sequential, /*lv_initialised=*/0, nullptr, stream);
    if (rc) return rc;
    if (info_host) {
        MBS_CHECK_CUDA(cudaMemcpyAsync(hp, st, sizeof(Stats), cudaMemcpyDeviceToHost, stream));
        MBS_CHECK_CUDA(cudaStreamSynchronize(stream));
        memset(info_host, 0, 8 * sizeof(int64_t));
        info_host[2] = hp->sweeps;
        info_host[3] = (force_sequential || hp->ambiguous) ? 1 : 0;
        info_host[4] = hp->ambiguous;
    }
    return 0;
}

extern "C" int mbs_distance_postprocessing(const float *border, const float *cell, int H, int W, int ld, float th_seed,
                                           float th_cell, uint16_t *out, void *workspace, size_t workspace_bytes,
                                           int64_t *info_host, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t n = static_cast<size_t>(H) * W;
    MBS_REQUIRE(H > 0 && W > 0 && ld >= W && n < (1ull << 31), "distance_postprocessing: bad shape H=%d W=%d ld=%d", H,
                W, ld);
    MBS_REQUIRE(workspace_bytes >= mbs_postproc_workspace_bytes(H, W),
                "distance_postprocessing: workspace too small (%zu < %zu)", workspace_bytes,
                mbs_postproc_workspace_bytes(H, W));
    Carver cv{static_cast<char *>(workspace), workspace_bytes};
    Stats *st = cv.take<Stats>(1);
    float *cell_s = cv.take<float>(n);
    uint8_t *mask = cv.take<uint8_t>(n);
    uint8_t *seed = cv.take<uint8_t>(n);
    int *roots = cv.take<int>(n);
    int *area = cv.take<int>(n);
    int *rank = cv.take<int>(n);
    int *tiles = cv.take<int>((n + SCAN_TILE - 1) / SCAN_TILE);
    int *markers = cv.take<int>(n);
    float *Lv = cv.take<float>(n);
    int *parent = cv.take<int>(n);
    int *uf = cv.take<int>(n);
    int *src = cv.take<int>(n);
    int *lab = cv.take<int>(n);
    HeapItem *heap = cv.take<HeapItem>(n);
    uint8_t *tile_changed = cv.take<uint8_t>(n / 64 + 4096);
    MBS_REQUIRE(cv.ok, "distance_postprocessing: workspace carve failed");
    Stats *hp = pinned_stats();
    MBS_REQUIRE(hp != nullptr, "distance_postprocessing: cannot allocate pinned host memory");
    const int nn = static_cast<int>(n);
    const int nb = mbs::cdiv(nn, 256);

    MBS_CHECK_CUDA(cudaMemsetAsync(st, 0, sizeof(Stats), stream));
    dim3 grid(mbs::cdiv(W, FT), mbs::cdiv(H, FT));
    pp_front_kernel<<<grid, 256, 0, stream>>>(border, cell, H, W, ld, th_seed, th_cell, cell_s, mask, seed, roots);
    MBS_CHECK_LAUNCH();
    int rc = run_label_rank(seed, H, W, roots, area, rank, tiles, st, /*use_mean=*/1, /*roots_initialised=*/1, stream);
    if (rc) return rc;
    markers_kernel<<<nb, 256, 0, stream>>>(roots, rank, mask, nn, markers, cell_s, 1, Lv);
    MBS_CHECK_LAUNCH();
    rc = run_watershed(cell_s, /*negate=*/1, markers, mask, H, W, lab, Lv, parent, uf, src, heap, st, tile_changed, 0,
                       /*lv_initialised=*/1, out, stream);
    if (rc) return rc;
    if (info_host) {
        MBS_CHECK_CUDA(cudaMemcpyAsync(hp, st, sizeof(Stats), cudaMemcpyDeviceToHost, stream));
        MBS_CHECK_CUDA(cudaStreamSynchronize(stream));
        memset(info_host, 0, 8 * sizeof(int64_t));
        info_host[0] = hp->n_comp;
        info_host[1] = hp->n_markers;
        info_host[2] = hp->sweeps;
        info_host[3] = hp->ambiguous ? 1 : 0;
        info_host[4] = hp->ambiguous;
    }
    return 0;
}

extern "C" int mbs_boundary_postprocessing(const float *prediction_hwc, int H, int W, uint16_t *out, void *workspace,
                                           size_t workspace_bytes, int64_t *info_host, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t n = static_cast<size_t>(H) * W;
    MBS_REQUIRE(H > 0 && W > 0 && n < (1ull << 31), "boundary_postprocessing: bad shape H=%d W=%d", H, W);
    MBS_REQUIRE(workspace_bytes >= mbs_postproc_workspace_bytes(H, W), "boundary_postprocessing: workspace too small");
    Carver cv{static_cast<char *>(workspace), workspace_bytes};
    Stats *st = cv.take<Stats>(1);
    float *img = cv.take<float>(n);
    uint8_t *mask = cv.take<uint8_t>(n);
    uint8_t *seed = cv.take<uint8_t>(n);
    int *roots = cv.take<int>(n);
    int *area = cv.take<int>(n);
    int *rank = cv.take<int>(n);
    int *tiles = cv.take<int>((n + SCAN_TILE - 1) / SCAN_TILE);
    int *markers = cv.take<int>(n);
    float *Lv = cv.take<float>(n);
    int *parent = cv.take<int>(n);
    int *uf = cv.take<int>(n);
    int *src = cv.take<int>(n);
    int *lab = cv.take<int>(n);
    HeapItem *heap = cv.take<HeapItem>(n);
    uint8_t *tile_changed = cv.take<uint8_t>(n / 64 + 4096);
    MBS_REQUIRE(cv.ok, "boundary_postprocessing: workspace carve failed");
    Stats *hp = pinned_stats();
    MBS_REQUIRE(hp != nullptr, "boundary_postprocessing: cannot allocate pinned host memory");
    const int nn = static_cast<int>(n);
    const int nb = mbs::cdiv(nn, 256);
    MBS_CHECK_CUDA(cudaMemsetAsync(st, 0, sizeof(Stats), stream));
    bp_front_kernel<<<nb, 256, 0, stream>>>(prediction_hwc, nn, img, mask, seed);
    MBS_CHECK_LAUNCH();
    int rc = run_label_rank(seed, H, W, roots, area, rank, tiles, st, /*use_mean=*/0, /*roots_initialised=*/0, stream);   // drop area <= 4
    if (rc) return rc;
    markers_kernel<<<nb, 256, 0, stream>>>(roots, rank, mask, nn, markers, img, 0, Lv);
    MBS_CHECK_LAUNCH();
    rc = run_watershed(img, /*negate=*/0, markers, mask, H, W, lab, Lv, parent, uf, src, heap, st, tile_changed, 0,
                       /*lv_initialised=*/1, out, stream);
    if (rc) return rc;
    if (info_host) {
        MBS_CHECK_CUDA(cudaMemcpyAsync(hp, st, sizeof(Stats), cudaMemcpyDeviceToHost, stream));
        MBS_CHECK_CUDA(cudaStreamSynchronize(stream));
        memset(info_host, 0, 8 * sizeof(int64_t));
        info_host[0] = hp->n_comp;
        info_host[1] = hp->n_markers;
        info_host[2] = hp->sweeps;
        info_host[3] = hp->ambiguous ? 1 : 0;
        info_host[4] = hp->ambiguous;
    }
    return 0;
}
