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
    unsigned int n_roots;       // tiled path: entries of the tile-local root list
    unsigned int overflow;      // tiled path: plateau hop counter saturated / root list full -> exact sequential flood
    // instrumentation of flood_kernel (tools/time_pp.py): tiles visited per sweep, %globaltimer at the end of each sweep
    unsigned int dbg_tiles[32];
    unsigned long long dbg_t[36];
    unsigned int dbg_rounds[32];    // sum over the visited tiles of the number of event rounds, per sweep
    unsigned int dbg_maxrounds[32];
    unsigned int dbg_items[32];     // queue entries processed, per sweep
    unsigned long long dbg_phase[5]; // block-wide visits in sweeps >= 3: clocks in load / queue build / rounds / store, count
    // exact fallback (tiled path)
    unsigned int n_bad;             // order-dependent pixels recorded by the flood (entries of the bad-pixel list)
    unsigned int need_global;       // a component could not be re-flooded on its own: whole-image sequential flood
    unsigned int n_reflooded;       // mask components re-flooded on their own
    unsigned int pool_top;          // heap pool allocator of the per-component floods
    unsigned int tile_counter;      // sweep 0 of flood_kernel: next tile to visit (dynamic assignment: visit costs vary 3x)
    unsigned int tile_counter2;     // the same for the final phase (4096 tiles on 444 blocks: 9.2 -> 10 visits when dealt statically)
    unsigned int dbg_maxclk[32];    // longest block-wide visit of each sweep (clock64 ticks)
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
    if (!force && st->ambiguous == 0 && st->overflow == 0) return;
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


// ==========================================================================================
// Tiled pipeline (round 2): 3 kernels + a gated exact fallback instead of 13 streaming passes.
//
//   front_ccl_kernel   gaussian / clip / tan / thresholds, and the 8-connected labelling of the seed pixels of one
//                      64x64 tile in SHARED memory (union-find with the smallest raster index as root).  Per pixel it
//                      writes the smoothed cell value (f32) and ONE 16-bit word: mask bit, seed bit, index of the
//                      tile-local root.  Global union-find state exists only at the tile-local roots (sparse).
//   mid_kernel         one cooperative launch over compact lists: unions across tile seams, path compression and
//                      component areas at the local roots, the area filter (float64, postprocessing.py:46-53), a bitmap
//                      of the surviving component roots and its prefix popcount = the raster-order relabelling.
//   flood_kernel       marker watershed, one cooperative launch.  State per pixel = 64 bits [level | hops | label]:
//                      level L(p) = minimax path cost from the markers, hops = distance to the place where the path
//                      reached that level (makes the "labelled by an earlier popped neighbour" relation well founded
//                      on plateaus), label = label of the neighbour with the smallest (level, hops).  64x64 tiles are
//                      relaxed to their fixed point in shared memory, sweeps are separated by grid barriers, only tiles
//                      next to a change are revisited.  The last phase writes the uint16 mask and runs the
//                      order-independence check of the file header (every neighbour at the minimal level must carry
//                      the pixel's label); any violation triggers the exact sequential flood.
// ==========================================================================================
constexpr int CT = 64;                       // tile edge of the tiled pipeline
constexpr unsigned L16_MASK = 0x8000u, L16_SEED = 0x4000u, L16_IDX = 0x0FFFu;
constexpr unsigned ORD_INF = 0xFF800000u;    // ordered image of +inf
constexpr unsigned long long ST_OUTSIDE = ~0ull;                                         // not in the mask: never updated
constexpr unsigned long long ST_UNREACHED = (static_cast<unsigned long long>(ORD_INF) << 32) | (0xFFFEull << 16);
constexpr unsigned HOP_MAX = 0x7FFEu;

// float -> unsigned with the same order (finite values and infinities); -0.0 and +0.0 map to the same code
__device__ __forceinline__ unsigned ord_f32(float v) {
    v = v + 0.0f;
    const unsigned b = __float_as_uint(v);
    return b ^ ((static_cast<unsigned>(static_cast<int>(b) >> 31)) | 0x80000000u);
}

__device__ __forceinline__ int l16_root_index(unsigned v16, int y, int x, int W) {
    const int idx = static_cast<int>(v16 & L16_IDX);
    return ((y & ~(CT - 1)) + (idx >> 6)) * W + (x & ~(CT - 1)) + (idx & (CT - 1));
}

template <bool BOUNDARY>
__global__ void __launch_bounds__(256, 5)
front_ccl_kernel(const float *__restrict__ border, const float *__restrict__ cell, int H, int W, int ld, float th_seed,
                 float th_cell, float *__restrict__ cell_s, uint16_t *__restrict__ L16, int *G, int *area,
                 int *__restrict__ roots_list, int list_cap, unsigned *bitmap, Stats *st) {
    __shared__ float s_in[CT + 4][CT + 4];      // later: tile-local union-find (int[4096])
    __shared__ float s_y[CT][CT + 4];           // later: tile-local component areas (int[4096])
    int *s_uf = reinterpret_cast<int *>(&s_in[0][0]);
    int *s_cnt = reinterpret_cast<int *>(&s_y[0][0]);
    const int x0 = blockIdx.x * CT, y0 = blockIdx.y * CT;
    const int tid = threadIdx.x;
    const int c = tid & (CT - 1), rq = tid >> 6;            // pixel k of this thread: row k*4 + rq, column c

    if (!BOUNDARY) {
        // tile + 2-pixel apron of the cell map.  Full tiles with 16-byte aligned rows: four independent float4 loads
        // per thread for the interior and ~2 scalar loads for the apron ring, all in flight together (the plain loop
        // serialises 18 load latencies per tile)
        const bool vec = x0 + CT <= W && y0 + CT <= H && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(cell) & 15) == 0;
        if (vec) {
            const int row = tid >> 2, col0 = (tid & 3) * 16;
            const float4 *src = reinterpret_cast<const float4 *>(cell + static_cast<size_t>(y0 + row) * ld + x0 + col0);
            const float4 f0 = src[0], f1 = src[1], f2 = src[2], f3 = src[3];
            float ring[3];
            int rr[3], rc[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int h = tid + j * 256;
                rr[j] = -1;
                if (h < 4 * (CT + 4) + 4 * CT) {
                    int r, cc;
                    if (h < 4 * (CT + 4)) {                       // rows 0, 1, CT+2, CT+3 over the full width
                        const int q = h / (CT + 4);
                        r = q < 2 ? q : CT + q;
                        cc = h % (CT + 4);
                    } else {                                       // columns 0, 1, CT+2, CT+3 of the interior rows
                        const int g = h - 4 * (CT + 4);
                        r = 2 + (g >> 2);
                        const int q = g & 3;
                        cc = q < 2 ? q : CT + q;
                    }
                    rr[j] = r;
                    rc[j] = cc;
                    const int yy = reflect_idx(y0 + r - 2, H), xx = reflect_idx(x0 + cc - 2, W);
                    ring[j] = cell[static_cast<size_t>(yy) * ld + xx];
                }
            }
            float *d = &s_in[row + 2][col0 + 2];
            d[0] = f0.x; d[1] = f0.y; d[2] = f0.z; d[3] = f0.w; d[4] = f1.x; d[5] = f1.y; d[6] = f1.z; d[7] = f1.w;
            d[8] = f2.x; d[9] = f2.y; d[10] = f2.z; d[11] = f2.w; d[12] = f3.x; d[13] = f3.y; d[14] = f3.z; d[15] = f3.w;
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (rr[j] >= 0) s_in[rr[j]][rc[j]] = ring[j];
        } else {
            for (int i = tid; i < (CT + 4) * (CT + 4); i += 256) {
                const int r = i / (CT + 4), cc = i % (CT + 4);
                const int yy = reflect_idx(y0 + r - 2, H), xx = reflect_idx(x0 + cc - 2, W);
                s_in[r][cc] = cell[static_cast<size_t>(yy) * ld + xx];
            }
        }
        __syncthreads();
        for (int i = tid; i < CT * (CT + 4); i += 256) {
            const int r = i / (CT + 4), cc = i % (CT + 4);
            s_y[r][cc] = gauss5(s_in[r][cc], s_in[r + 1][cc], s_in[r + 2][cc], s_in[r + 3][cc], s_in[r + 4][cc]);
        }
        __syncthreads();
    }
    // the bitmap of surviving roots is filled by mid_kernel: clear the words this tile's pixels fall into
    for (int r = tid; r < CT; r += 256) {
        const int y = y0 + r;
        if (y < H) {
            const long long b0 = static_cast<long long>(y) * W + x0;
            long long b1 = b0 + CT - 1;
            const long long rowend = static_cast<long long>(y) * W + W - 1;
            if (b1 > rowend) b1 = rowend;
            for (long long w = b0 >> 5; w <= (b1 >> 5); ++w) bitmap[w] = 0u;
        }
    }
    unsigned sdbits = 0, mkbits = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int r = k * 4 + rq;
        const int y = y0 + r, x = x0 + c;
        bool sd = false, mk = false;
        if (y < H && x < W) {
            const size_t o = static_cast<size_t>(y) * W + x;
            if (BOUNDARY) {
                // boundary_postprocessing front end (postprocessing.py:71-77): `border` holds the (H,W,3) probabilities
                const float p0 = border[3 * o], p1 = border[3 * o + 1], p2 = border[3 * o + 2];
                int am = 0;
                float best = p0;
                if (p1 > best) { best = p1; am = 1; }
                if (p2 > best) { am = 2; }
                mk = am == 1;
                sd = __fmul_rn(p1, __fsub_rn(1.0f, p2)) > 0.5f;
                cell_s[o] = mk ? 1.0f : 0.0f;
            } else {
                const float cs = gauss5(s_y[r][c], s_y[r][c + 1], s_y[r][c + 2], s_y[r][c + 3], s_y[r][c + 4]);
                float b = border[static_cast<size_t>(y) * ld + x];
                b = b < 0.0f ? 0.0f : (b > 1.0f ? 1.0f : b);                    // np.clip keeps NaN
                const float sq = __fmul_rn(b, b);
                // float32(tan_f64(x)); tan is increasing on [0,1] and tan(0.0499) = 0.04994 < 0.05, so
                // every sq below 0.0499 lands in the `borders < 0.05 -> 0` branch without evaluating tan
                cell_s[o] = cs;
                mk = cs > th_cell;
                if (sq < 0.0499f) {
                    sd = cs > th_seed;                                          // borders == 0: cleaned = cs - 0
                } else {
                    // Only the DECISION (cs - borders) > th_seed is needed, not the value of tan.  tanf is within 4 ulp
                    // (< 6e-7 on [0, 1.56]) of the exact tangent and so is float32(tan_f64); if the float32 estimate keeps
                    // the three comparisons that use it (borders < 0.05, borders > 1, cleaned > th_seed) at least 1e-5
                    // away from their thresholds, the exact evaluation cannot decide differently.  Otherwise (about one
                    // pixel in 10^4) the float64 tangent is evaluated as before.  (The double-precision tan costs ~150
                    // FP64 instructions per warp and was half of this kernel's time.)
                    const float tf = tanf(sq);
                    const float te = tf > 1.0f ? 1.0f : tf;
                    const bool sure = fabsf(tf - 0.05f) > 1e-5f && fabsf(tf - 1.0f) > 1e-5f && fabsf((cs - te) - th_seed) > 1e-5f &&
                                      tf >= 0.05f;
                    if (sure) {
                        sd = (cs - te) > th_seed;
                    } else {
                        float t = static_cast<float>(tan(static_cast<double>(sq)));
                        if (t < 0.05f) t = 0.0f;
                        t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
                        sd = __fsub_rn(cs, t) > th_seed;
                    }
                }
            }
        }
        sdbits |= (sd ? 1u : 0u) << k;
        mkbits |= (mk ? 1u : 0u) << k;
    }
    __syncthreads();                      // s_in / s_y are dead: reuse them as s_uf / s_cnt
    // ---- tile-local 8-connected labelling on RUNS: the seed pixels of a row are a 64-bit mask, a pixel finds the start
    // of its horizontal run with two bit operations, and only run starts take part in the union-find (one union per
    // pair of touching runs in consecutive rows instead of up to four per pixel)
    __shared__ unsigned s_row32[CT][2];
    const int lane = tid & 31;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const unsigned bal = __ballot_sync(0xffffffffu, (sdbits >> k) & 1u);
        if (lane == 0) s_row32[k * 4 + rq][(tid >> 5) & 1] = bal;
    }
    __syncthreads();
    auto row_mask = [&](int r) -> unsigned long long {
        return static_cast<unsigned long long>(s_row32[r][0]) | (static_cast<unsigned long long>(s_row32[r][1]) << 32);
    };
    auto run_start = [](unsigned long long m, int col) -> int {       // first column of the run of ones that contains `col`
        const unsigned long long below = (1ull << col) - 1ull;
        const unsigned long long zeros = ~m & below;
        return zeros ? 64 - __clzll(static_cast<long long>(zeros)) : 0;
    };
    auto run_mask = [](unsigned long long m, int start) -> unsigned long long {     // the run of ones that starts at `start`
        const unsigned long long t = m >> start;
        const int len = (~t) ? __ffsll(static_cast<long long>(~t)) - 1 : 64;
        const unsigned long long ones = len >= 64 ? ~0ull : ((1ull << len) - 1ull);
        return ones << start;
    };
    unsigned startbits = 0;               // pixel k is the first pixel of its run
    unsigned char rs[16];                 // first column of pixel k's run
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        rs[k] = 0;
        if (!((sdbits >> k) & 1u)) continue;
        const int r = k * 4 + rq;
        const int st0 = run_start(row_mask(r), c);
        rs[k] = static_cast<unsigned char>(st0);
        if (st0 == c) {
            startbits |= 1u << k;
            s_uf[r * CT + c] = r * CT + c;
            s_cnt[r * CT + c] = 0;
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < 16; ++k) {
        if (!((startbits >> k) & 1u)) continue;
        const int r = k * 4 + rq;
        if (r == 0) continue;
        const unsigned long long run = run_mask(row_mask(r), c);
        const unsigned long long above = row_mask(r - 1);
        unsigned long long touch = above & (run | (run << 1) | (run >> 1));
        while (touch) {
            const int bcol = __ffsll(static_cast<long long>(touch)) - 1;
            const int sa = run_start(above, bcol);
            uf_union(s_uf, r * CT + c, (r - 1) * CT + sa);
            touch &= ~run_mask(above, sa);
        }
    }
    __syncthreads();
    // flatten the run starts, areas of the tile-local components
#pragma unroll 1
    for (int k = 0; k < 16; ++k) {
        if (!((startbits >> k) & 1u)) continue;
        const int i = (k * 4 + rq) * CT + c;
        const int rt = uf_find_v(s_uf, i);
        const unsigned long long run = run_mask(row_mask(k * 4 + rq), c);
        atomicAdd(&s_cnt[rt], __popcll(run));
        if (rt != i) *reinterpret_cast<volatile int *>(&s_uf[i]) = rt;      // chains stay valid: rt is a root
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int r = k * 4 + rq;
        const int y = y0 + r, x = x0 + c;
        const bool sd = (sdbits >> k) & 1u;
        int rt = 0;
        if (sd) rt = uf_find_v(s_uf, r * CT + rs[k]);
        const bool isroot = sd && rt == r * CT + c;
        if (y < H && x < W) {
            const unsigned v = ((mkbits >> k) & 1u ? L16_MASK : 0u) | (sd ? (L16_SEED | static_cast<unsigned>(rt)) : 0u);
            L16[static_cast<size_t>(y) * W + x] = static_cast<uint16_t>(v);
        }
        const unsigned m = __ballot_sync(0xffffffffu, isroot);
        if (m) {
            int base = 0;
            if (lane == __ffs(m) - 1) base = static_cast<int>(atomicAdd(&st->n_roots, static_cast<unsigned>(__popc(m))));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (isroot) {
                const int g = y * W + x;
                G[g] = g;
                area[g] = s_cnt[r * CT + c];
                const int slot = base + __popc(m & ((1u << lane) - 1u));
                if (slot < list_cap) roots_list[slot] = g;
                else atomicExch(&st->overflow, 1u);
            }
        }
    }
    // seed pixels of the tile -> st->total (one atomic per block)
    int cnt = __popc(sdbits);
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    __shared__ int s_red[8];
    if (lane == 0) s_red[tid >> 5] = cnt;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        if (t) atomicAdd(&st->total, static_cast<unsigned long long>(t));
    }
}

__device__ __forceinline__ bool keep_area(int area, unsigned long long total, unsigned n_comp, int use_mean) {
    // postprocessing.py:46-53: min_area = max(0.10 * mean(areas), 4); drop area <= min_area (float64)
    double min_area = 0.0;
    if (use_mean && n_comp > 0)
        min_area = __dmul_rn(0.10, __ddiv_rn(static_cast<double>(total), static_cast<double>(n_comp)));
    min_area = fmax(min_area, 4.0);
    return !(static_cast<double>(area) <= min_area);
}

__global__ void __launch_bounds__(256)
mid_kernel(const uint16_t *__restrict__ L16, int H, int W, int *G, int *area, const int *__restrict__ roots_list,
           int list_cap, unsigned *bitmap, int *prefix, int *block_sums, Stats *st, int use_mean) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int gsz = gridDim.x * blockDim.x;
    const int tiles_x = (W + CT - 1) / CT, tiles_y = (H + CT - 1) / CT;
    __shared__ int s_red[8];
    __shared__ int s_off;
    // ---- phase 0: unions across tile seams (8-connectivity: three neighbours on the other side of the seam)
    {
        const long long nv = static_cast<long long>(tiles_x - 1) * H, nh = static_cast<long long>(tiles_y - 1) * W;
        for (long long it = gtid; it < nv + nh; it += gsz) {
            int ya, xa, yb0, xb0, dy, dx;
            if (it < nv) {
                const int k = static_cast<int>(it / H) + 1;
                ya = static_cast<int>(it % H); xa = k * CT - 1; yb0 = ya; xb0 = k * CT; dy = 1; dx = 0;
            } else {
                const long long j = it - nv;
                const int k = static_cast<int>(j / W) + 1;
                xa = static_cast<int>(j % W); ya = k * CT - 1; yb0 = k * CT; xb0 = xa; dy = 0; dx = 1;
            }
            const unsigned va = L16[static_cast<size_t>(ya) * W + xa];
            if (!(va & L16_SEED)) continue;
            const int ra = l16_root_index(va, ya, xa, W);
            for (int s = -1; s <= 1; ++s) {
                const int yb = yb0 + s * dy, xb = xb0 + s * dx;
                if (yb < 0 || yb >= H || xb < 0 || xb >= W) continue;
                const unsigned vb = L16[static_cast<size_t>(yb) * W + xb];
                if (vb & L16_SEED) uf_union(G, ra, l16_root_index(vb, yb, xb, W));
            }
        }
    }
    __threadfence();
    grid.sync();
    // ---- phase 1: compress the local roots, accumulate areas at the component roots, count components
    int n_roots = static_cast<int>(*reinterpret_cast<volatile unsigned int *>(&st->n_roots));
    if (n_roots > list_cap) n_roots = list_cap;
    {
        int comps = 0;
        for (int i = gtid; i < n_roots; i += gsz) {
            const int lr = roots_list[i];
            const int r = uf_find_v(G, lr);
            if (r != lr) {
                G[lr] = r;
                atomicAdd(&area[r], area[lr]);
            } else {
                ++comps;
            }
        }
        int v = comps;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += s_red[w];
            if (t) atomicAdd(&st->n_comp, static_cast<unsigned>(t));
        }
    }
    __threadfence();
    grid.sync();
    // ---- phase 2: area filter -> bitmap of the surviving component roots
    {
        const unsigned long long total = *reinterpret_cast<volatile unsigned long long *>(&st->total);
        const unsigned n_comp = *reinterpret_cast<volatile unsigned int *>(&st->n_comp);
        for (int i = gtid; i < n_roots; i += gsz) {
            const int lr = roots_list[i];
            if (*reinterpret_cast<volatile int *>(&G[lr]) != lr) continue;
            const int a = *reinterpret_cast<volatile int *>(&area[lr]);
            if (keep_area(a, total, n_comp, use_mean)) atomicOr(&bitmap[lr >> 5], 1u << (lr & 31));
        }
    }
    __threadfence();
    grid.sync();
    // ---- phase 3: prefix popcount over the bitmap words (rank of a root = 1 + number of surviving roots before it)
    const int nwords = static_cast<int>((static_cast<long long>(H) * W + 31) >> 5);
    const int per = (nwords + gridDim.x - 1) / gridDim.x;
    const int w0 = blockIdx.x * per;
    const int w1 = (w0 + per < nwords) ? w0 + per : nwords;
    {
        int v = 0;
        for (int w = w0 + threadIdx.x; w < w1; w += blockDim.x) v += __popc(*reinterpret_cast<volatile unsigned *>(&bitmap[w]));
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += s_red[w];
            block_sums[blockIdx.x] = t;
        }
    }
    __threadfence();
    grid.sync();
    {
        int v = 0;
        for (int b = threadIdx.x; b < static_cast<int>(blockIdx.x); b += blockDim.x) v += *reinterpret_cast<volatile int *>(&block_sums[b]);
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += s_red[w];
            s_off = t;
        }
        __syncthreads();
        int running = s_off;
        for (int base = w0; base < w1; base += blockDim.x) {
            const int w = base + threadIdx.x;
            const int pc = w < w1 ? __popc(*reinterpret_cast<volatile unsigned *>(&bitmap[w])) : 0;
            int inc = pc;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if ((threadIdx.x & 31) >= o) inc += t;
            }
            __syncthreads();
            if ((threadIdx.x & 31) == 31) s_red[threadIdx.x >> 5] = inc;
            __syncthreads();
            int woff = 0, tot = 0;
            for (int q = 0; q < 8; ++q) {
                if (q < (threadIdx.x >> 5)) woff += s_red[q];
                tot += s_red[q];
            }
            if (w < w1) prefix[w] = running + woff + inc - pc;
            running += tot;
        }
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) st->n_markers = static_cast<unsigned>(running);
    }
    __threadfence();
    grid.sync();
    // ---- phase 4: marker id (1-based raster rank of the surviving component, 0 = dropped) of every tile-local root,
    // stored over its (no longer needed) area: the flood resolves a seed pixel with ONE gather instead of three
    for (int i = gtid; i < n_roots; i += gsz) {
        const int lr = roots_list[i];
        const int r = *reinterpret_cast<volatile int *>(&G[lr]);
        const unsigned w = *reinterpret_cast<volatile unsigned *>(&bitmap[r >> 5]);
        int id = 0;
        if ((w >> (r & 31)) & 1u) id = *reinterpret_cast<volatile int *>(&prefix[r >> 5]) + __popc(w & ((1u << (r & 31)) - 1u)) + 1;
        area[lr] = id;
    }
}

// marker id (1-based raster rank of the surviving seed component) of a seed pixel, 0 if its component was dropped:
// mid_kernel phase 4 left it at the pixel's tile-local root
__device__ __forceinline__ int marker_of(unsigned v16, int y, int x, int W, const int *__restrict__ marker_at_root) {
    return marker_at_root[l16_root_index(v16, y, x, W)];
}

// generic entry (mbs_pp_watershed): state / labels from an explicit marker image
__global__ void flood_init_kernel(const float *__restrict__ img, const int *__restrict__ markers, const uint8_t *__restrict__ mask,
                                  int n, unsigned long long *__restrict__ state, int *__restrict__ lab32) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool m = mask[i] != 0;
    const int mk = m ? markers[i] : 0;
    unsigned long long s = ST_OUTSIDE;
    if (m) s = mk != 0 ? ((static_cast<unsigned long long>(ord_f32(img[i])) << 32) | (1ull << 16)) : ST_UNREACHED;
    state[i] = s;
    lab32[i] = mk;
}

struct FloodParams {
    const float *img;                // flood image (negate != 0: -img)
    int negate;
    const uint16_t *L16;             // fused mode: seeds / mask / local roots from front_ccl_kernel
    const int *marker_at_root;       // fused mode: marker id of every tile-local root (mid_kernel phase 4)
    int light_init_max, light_items_max;   // a warp-level revisit is abandoned (and finished block-wide) beyond these sizes
    int *badlist;                          // fused mode: order-dependent pixels (exact fallback re-floods their mask components)
    int bad_cap;
    int light_min_tiles;                   // warp-level revisits only if the block has at least this many tiles to revisit
    int chase_heavy, chase_light;          // pixels a thread follows straight ahead per queue entry (block-wide / warp-level visit)
    float th_mask;                   // threshold sweeps: mask = img > th_mask instead of the mask bit stored by the front end
    int use_th_mask;
    unsigned long long *state;
    int *lab32;                      // LAB32 only
    int H, W;
    Stats *st;
    uint8_t *tile_changed;
    uint16_t *out16;                 // fused mode
    int *out32;                      // LAB32 mode
};

// shared memory of flood_kernel: state tile with a 1-pixel halo, flood values, two work queues, "queued" bits
constexpr unsigned EDGE_DIRTY = 16u;          // a light revisit ran out of queue space: full block-wide visit next sweep
constexpr int LQ = 1024;                      // queue entries of a warp-level ("light") revisit
constexpr int LIGHT_WARP_BYTES = 2 * LQ * 2 + (CT * CT / 32) * 4 + 16;
constexpr int FL_LIST_MAX = 256;              // light revisits a block can defer per sweep (more: handled block-wide)
constexpr int FL_STATE_BYTES = (CT + 2) * (CT + 2) * 8;
constexpr int FL_V_BYTES = CT * CT * 4;
constexpr int FL_Q_BYTES = 2 * CT * CT * 2;
constexpr int FL_FLAG_BYTES = (CT * CT / 32) * 4;
constexpr int FL_MISC_BYTES = 64 + 2 * (4 * FL_LIST_MAX + 16);       // queue sizes / edge bits / flags, list of light revisits
constexpr int FL_LAB_BYTES = (CT + 2) * (CT + 2) * 4;
constexpr int FLOOD_SMEM16 = FL_STATE_BYTES + FL_V_BYTES + FL_Q_BYTES + FL_FLAG_BYTES + FL_MISC_BYTES;
constexpr int FLOOD_SMEM32 = FLOOD_SMEM16 + FL_LAB_BYTES;
constexpr unsigned EDGE_TOP = 1u, EDGE_BOTTOM = 2u, EDGE_LEFT = 4u, EDGE_RIGHT = 8u;

// The in-tile relaxation is EVENT DRIVEN: a pixel is (re)evaluated only when one of its 4-neighbours changed.  Work is
// proportional to the number of floodable pixels (a few evaluations each) instead of tile area x wavefront depth.
// Tile loads / stores are bulk 16-byte accesses issued back to back (a tile visit used to cost ~20 us of serialised
// load latency); thread t owns the 16 consecutive pixels (row t/4, columns (t%4)*16 ..) in those phases.
template <bool LAB32>
__global__ void __launch_bounds__(256, 3)
flood_kernel(const FloodParams p) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ unsigned long long smem_u64[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(smem_u64);
    unsigned long long(*sS)[CT + 2] = reinterpret_cast<unsigned long long(*)[CT + 2]>(smem);
    unsigned *sV = reinterpret_cast<unsigned *>(smem + FL_STATE_BYTES);
    unsigned short *sQ = reinterpret_cast<unsigned short *>(smem + FL_STATE_BYTES + FL_V_BYTES);      // [2][4096]
    unsigned *sFlag = reinterpret_cast<unsigned *>(smem + FL_STATE_BYTES + FL_V_BYTES + FL_Q_BYTES);
    int *sMisc = reinterpret_cast<int *>(smem + FL_STATE_BYTES + FL_V_BYTES + FL_Q_BYTES + FL_FLAG_BYTES);   // [0],[1]: queue sizes, [2]: edge bits, [3]: any change
    int(*sLab)[CT + 2] = reinterpret_cast<int(*)[CT + 2]>(smem + FLOOD_SMEM16);
    int *sList = sMisc + 16;                     // [0, FL_LIST_MAX): tiles deferred to a warp-level revisit, [FL_LIST_MAX]: "one of them changed"
    int *sAbort = sList + FL_LIST_MAX + 4;       // [0, FL_LIST_MAX): abandoned light revisits, [FL_LIST_MAX]: their count
    const int H = p.H, W = p.W;
    const int tiles_x = (W + CT - 1) / CT, tiles_y = (H + CT - 1) / CT;
    const int ntiles = tiles_x * tiles_y;
    const int lane = threadIdx.x & 31;
    int sweep = 0;
    bool overflow = false;
    auto stamp = [&](int slot) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && slot < 36) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.st->dbg_t[slot] = t;
        }
    };
    // halo ring of the tile at (y0, x0): 260 elements, element h of thread t = t and (t < 4) 256 + t
    auto load_halo = [&](int x0, int y0) {
        for (int h = threadIdx.x; h < 4 * CT + 4; h += 256) {
            int r, c;
            if (h < CT + 2) { r = 0; c = h; }
            else if (h < 2 * (CT + 2)) { r = CT + 1; c = h - (CT + 2); }
            else if (h < 2 * (CT + 2) + CT) { r = h - 2 * (CT + 2) + 1; c = 0; }
            else { r = h - 2 * (CT + 2) - CT + 1; c = CT + 1; }
            const int y = y0 + r - 1, x = x0 + c - 1;
            const bool in = y >= 0 && y < H && x >= 0 && x < W;
            sS[r][c] = in ? p.state[static_cast<size_t>(y) * W + x] : ST_OUTSIDE;
            if (LAB32) sLab[r][c] = in ? p.lab32[static_cast<size_t>(y) * W + x] : 0;
        }
    };
    // ---- block-wide visit of one tile: stage it in shared memory, relax to the local fixed point, write it back
    long long ph[5] = {0, 0, 0, 0, 0};          // clock64 spent in load / queue build / rounds / store, visits (sweeps >= 3)
    const uint8_t *prev = nullptr;
    uint8_t *cur = nullptr;
    bool block_changed = false;
    auto heavy_visit = [&](const int tile, const bool full_scan, const unsigned extra_edges) {
            const int tyi = tile / tiles_x, txi = tile - tyi * tiles_x;
            const int x0 = txi * CT, y0 = tyi * CT;
            __syncthreads();   // shared tile reuse
            const long long tk0 = clock64();
            const bool fused_init = !LAB32 && sweep == 0;
            if (threadIdx.x == 0 && sweep < 32) atomicAdd(&p.st->dbg_tiles[sweep], 1u);
            if (threadIdx.x < CT * CT / 32) sFlag[threadIdx.x] = 0u;
            if (threadIdx.x < 4) sMisc[threadIdx.x] = 0;
            // thread t owns column t % 64 of rows k*4 + t/64 (k = 0..15): a warp touches 32 adjacent columns of one row, so
            // global accesses are coalesced and shared-memory accesses conflict free; the sixteen loads of each array are
            // independent and issued back to back
            const int pc = threadIdx.x & (CT - 1), pr0 = threadIdx.x >> 6;
            const int gx = x0 + pc;
            if (fused_init) {
                // halo: the neighbouring tiles have no state yet; they are picked up in sweep 1
                for (int i = threadIdx.x; i < 4 * (CT + 2); i += 256) {
                    const int side = i / (CT + 2), j = i % (CT + 2);
                    const int r = side == 0 ? 0 : (side == 1 ? CT + 1 : j);
                    const int c = side <= 1 ? j : (side == 2 ? 0 : CT + 1);
                    sS[r][c] = ST_OUTSIDE;
                }
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {           // two batches of eight rows (register budget)
                    unsigned short v16[8];
                    float v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int y = y0 + (h * 8 + k) * 4 + pr0;
                        const bool in = y < H && gx < W;
                        v16[k] = in ? p.L16[static_cast<size_t>(y) * W + gx] : static_cast<unsigned short>(0);
                        v[k] = in ? p.img[static_cast<size_t>(y) * W + gx] : 0.0f;
                    }
                    if (p.use_th_mask) {
                        // threshold sweep: one front end / labelling pass serves several cell thresholds (the seeds do not
                        // depend on th_cell); mask = cell_s > th_cell is re-derived from the smoothed map
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const bool in = y0 + (h * 8 + k) * 4 + pr0 < H && gx < W;
                            v16[k] = static_cast<unsigned short>((v16[k] & ~L16_MASK) | ((in && v[k] > p.th_mask) ? L16_MASK : 0u));
                        }
                    }
                    int mk[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {          // independent gathers, all in flight together
                        mk[k] = 0;
                        if ((v16[k] & (L16_MASK | L16_SEED)) == (L16_MASK | L16_SEED))
                            mk[k] = marker_of(v16[k], y0 + (h * 8 + k) * 4 + pr0, gx, W, p.marker_at_root);
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int r = (h * 8 + k) * 4 + pr0;
                        unsigned long long s = ST_OUTSIDE;
                        if (v16[k] & L16_MASK) {
                            const unsigned vo = ord_f32(p.negate ? -v[k] : v[k]);
                            sV[r * CT + pc] = vo;
                            s = mk[k] > 0 ? ((static_cast<unsigned long long>(vo) << 32) | (1ull << 16) | static_cast<unsigned>(mk[k] & 0xFFFF))
                                          : ST_UNREACHED;
                        }
                        sS[r + 1][pc + 1] = s;
                    }
                }
            } else {
                load_halo(x0, y0);
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {           // two batches of eight rows: 24 values in flight per thread
                    unsigned long long sv[8];
                    float fv[8];
                    int lv[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int y = y0 + (h * 8 + k) * 4 + pr0;
                        const bool in = y < H && gx < W;
                        sv[k] = in ? p.state[static_cast<size_t>(y) * W + gx] : ST_OUTSIDE;
                        fv[k] = in ? p.img[static_cast<size_t>(y) * W + gx] : 0.0f;
                        if (LAB32) lv[k] = in ? p.lab32[static_cast<size_t>(y) * W + gx] : 0;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int r = (h * 8 + k) * 4 + pr0;
                        sS[r + 1][pc + 1] = sv[k];
                        sV[r * CT + pc] = ord_f32(p.negate ? -fv[k] : fv[k]);
                        if (LAB32) sLab[r + 1][pc + 1] = lv[k];
                    }
                }
            }
            __syncthreads();
            const long long tk1 = clock64();
            // initial work list = one bit per pixel.  First visit: every floodable pixel next to a flooded one; later visits:
            // the floodable pixels on the tile edges next to a flooded halo pixel (the inside is already at its fixed point).
            // A warp covers 32 adjacent columns of one row = exactly one word of the bitmap: a ballot and a plain store.
            const bool first_visit = full_scan;
            if (first_visit) {
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const int r = i * 4 + (threadIdx.x >> 6), c = threadIdx.x & (CT - 1);
                    const unsigned long long s = sS[r + 1][c + 1];
                    bool want = false;
                    if (!((s >> 16) & 1ull))
                        want = static_cast<unsigned>(sS[r][c + 1] >> 32) < ORD_INF || static_cast<unsigned>(sS[r + 2][c + 1] >> 32) < ORD_INF ||
                               static_cast<unsigned>(sS[r + 1][c] >> 32) < ORD_INF || static_cast<unsigned>(sS[r + 1][c + 2] >> 32) < ORD_INF;
                    const unsigned m = __ballot_sync(0xffffffffu, want);
                    if (lane == 0) sFlag[(r * CT + c) >> 5] = m;
                }
            } else {
                // the bitmap was cleared before the tile was loaded; 4 x 64 edge pixels = one per thread
                const int side = threadIdx.x >> 6, k = threadIdx.x & (CT - 1);
                const int r = side == 0 ? 0 : (side == 1 ? CT - 1 : k), c = side <= 1 ? k : (side == 2 ? 0 : CT - 1);
                const int hr = side == 0 ? 0 : (side == 1 ? CT + 1 : k + 1), hc = side <= 1 ? k + 1 : (side == 2 ? 0 : CT + 1);
                const unsigned long long s = sS[r + 1][c + 1];
                if (!((s >> 16) & 1ull) && static_cast<unsigned>(sS[hr][hc] >> 32) < ORD_INF) {
                    const int idx = r * CT + c;
                    atomicOr(&sFlag[idx >> 5], 1u << (idx & 31));
                }
            }
            const long long tk2 = clock64();
            // Rounds.  The work list IS the bitmap: a thread that changes a pixel marks the lateral neighbours with
            // fire-and-forget shared-memory reductions (no returned value, no queue counter: nothing in the dependent chain of
            // a chase step waits for an atomic).  Each round first compacts the set bits into a list (so that the items are
            // dealt evenly to the threads) and clears the bitmap; the block barriers between compaction and evaluation order
            // "cleared before the neighbours are read" and "state written before the flag is consumed".
            unsigned n_rounds = 0, n_items = 0;
            int *sWarpTot = sMisc + 4;                   // [8]
            for (;;) {
                __syncthreads();                         // flags of the previous round (or of the initial scan) are complete
                const int w = threadIdx.x >> 1, hbit = (threadIdx.x & 1) * 16;
                unsigned bits = (sFlag[w] >> hbit) & 0xFFFFu;
                const int cnt = __popc(bits);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                if (lane == 31) sWarpTot[threadIdx.x >> 5] = incl;
                if (!(threadIdx.x & 1)) sFlag[w] = 0u;   // both readers of the word are adjacent lanes of this (converged) warp
                __syncthreads();
                int basep = 0, n = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int v = sWarpTot[k];
                    if (k < (threadIdx.x >> 5)) basep += v;
                    n += v;
                }
                if (n == 0) break;
                int pos = basep + incl - cnt;
                while (bits) {
                    const int bpos = __ffs(bits) - 1;
                    bits &= bits - 1;
                    sQ[pos++] = static_cast<unsigned short>(w * 32 + hbit + bpos);
                }
                __syncthreads();
                ++n_rounds;
                n_items += static_cast<unsigned>(n);
                const int CHASE = p.chase_heavy;
                for (int i = threadIdx.x; i < n; i += 256) {
                    int idx = sQ[i];
                    // One list entry = one evaluation plus a CHASE: while a pixel changes, the thread carries on with the
                    // neighbour straight ahead (away from the neighbour the new state came from), so a front crosses up to
                    // CHASE pixels per round along rows / columns instead of one.
#pragma unroll 1
                    for (int step = 0; step <= CHASE; ++step) {
                        const int r = (idx >> 6) + 1, c = (idx & (CT - 1)) + 1;
                        // skimage's neighbour order: up, left, right, down; the first minimal (level, hops) wins
                        const unsigned long long n0 = *reinterpret_cast<volatile unsigned long long *>(&sS[r - 1][c]);
                        const unsigned long long n1 = *reinterpret_cast<volatile unsigned long long *>(&sS[r][c - 1]);
                        const unsigned long long n2 = *reinterpret_cast<volatile unsigned long long *>(&sS[r][c + 1]);
                        const unsigned long long n3 = *reinterpret_cast<volatile unsigned long long *>(&sS[r + 1][c]);
                        const unsigned long long old = *reinterpret_cast<volatile unsigned long long *>(&sS[r][c]);
                        const unsigned vo = sV[idx];
                        unsigned long long best = n0;
                        int bi = 0;
                        if ((n1 >> 16) < (best >> 16)) { best = n1; bi = 1; }
                        if ((n2 >> 16) < (best >> 16)) { best = n2; bi = 2; }
                        if ((n3 >> 16) < (best >> 16)) { best = n3; bi = 3; }
                        const unsigned Lq = static_cast<unsigned>(best >> 32);
                        if (Lq >= ORD_INF) break;                        // no flooded neighbour yet
                        unsigned long long ns;
                        if (vo > Lq) {
                            ns = (static_cast<unsigned long long>(vo) << 32) | (best & 0xFFFFull);
                        } else {
                            unsigned h = ((static_cast<unsigned>(best >> 16) & 0xFFFFu) >> 1) + 1u;
                            if (h > HOP_MAX) { h = HOP_MAX; overflow = true; }
                            ns = (static_cast<unsigned long long>(Lq) << 32) | (static_cast<unsigned long long>(h << 1) << 16) | (best & 0xFFFFull);
                        }
                        bool ch = ns != old;
                        if (LAB32) {
                            const int nl = bi == 0 ? sLab[r - 1][c] : (bi == 1 ? sLab[r][c - 1] : (bi == 2 ? sLab[r][c + 1] : sLab[r + 1][c]));
                            if (nl != sLab[r][c]) { *reinterpret_cast<volatile int *>(&sLab[r][c]) = nl; ch = true; }
                        }
                        if (!ch) break;
                        *reinterpret_cast<volatile unsigned long long *>(&sS[r][c]) = ns;
                        // the pixel straight ahead is handled by this thread in the next step; the other floodable
                        // neighbours inside the tile are marked (their "fixed" bit never changes: reuse n0..n3)
                        const int ahead = 3 - bi;                        // came from up -> go down, left -> right, ...
                        bool w0 = r > 1 && !((n0 >> 16) & 1ull), w1 = c > 1 && !((n1 >> 16) & 1ull);
                        bool w2 = c < CT && !((n2 >> 16) & 1ull), w3 = r < CT && !((n3 >> 16) & 1ull);
                        const bool go = step < CHASE && (ahead == 0 ? w0 : (ahead == 1 ? w1 : (ahead == 2 ? w2 : w3)));
                        if (go) {
                            if (ahead == 0) w0 = false; else if (ahead == 1) w1 = false; else if (ahead == 2) w2 = false; else w3 = false;
                        }
                        const int i0 = idx - CT, i1 = idx - 1, i2 = idx + 1, i3 = idx + CT;
                        if (w0) atomicOr(&sFlag[i0 >> 5], 1u << (i0 & 31));
                        if (w1) atomicOr(&sFlag[i1 >> 5], 1u << (i1 & 31));
                        if (w2) atomicOr(&sFlag[i2 >> 5], 1u << (i2 & 31));
                        if (w3) atomicOr(&sFlag[i3 >> 5], 1u << (i3 & 31));
                        const unsigned eb = (r == 1 ? EDGE_TOP : 0u) | (r == CT ? EDGE_BOTTOM : 0u) | (c == 1 ? EDGE_LEFT : 0u) | (c == CT ? EDGE_RIGHT : 0u);
                        if (eb) atomicOr(reinterpret_cast<unsigned *>(&sMisc[2]), eb);
                        sMisc[3] = 1;
                        if (!go) break;
                        idx = ahead == 0 ? i0 : (ahead == 1 ? i1 : (ahead == 2 ? i2 : i3));
                    }
                }
            }
            const long long tk3 = clock64();
            if (threadIdx.x == 0 && sweep < 32) {
                atomicAdd(&p.st->dbg_rounds[sweep], n_rounds);
                atomicMax(&p.st->dbg_maxrounds[sweep], n_rounds);
                atomicAdd(&p.st->dbg_items[sweep], n_items);
            }
            const bool changed_any = *reinterpret_cast<volatile int *>(&sMisc[3]) != 0;
            unsigned edge_bits = *reinterpret_cast<volatile unsigned *>(&sMisc[2]);
            unsigned my_edges = 0;
            if (changed_any || fused_init || (LAB32 && sweep == 0)) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int r = k * 4 + pr0;
                    const int y = y0 + r;
                    if (y < H && gx < W) {
                        const unsigned long long s = sS[r + 1][pc + 1];
                        if (fused_init || !((s >> 16) & 1ull)) {
                            p.state[static_cast<size_t>(y) * W + gx] = s;
                            if (LAB32) p.lab32[static_cast<size_t>(y) * W + gx] = sLab[r + 1][pc + 1];
                        }
                        // first visit: the neighbours have not seen this tile yet -- every flooded edge pixel counts
                        if (sweep == 0 && static_cast<unsigned>(s >> 32) < ORD_INF)
                            my_edges |= (r == 0 ? EDGE_TOP : 0u) | (r == CT - 1 ? EDGE_BOTTOM : 0u) | (pc == 0 ? EDGE_LEFT : 0u) |
                                        (pc == CT - 1 ? EDGE_RIGHT : 0u);
                    }
                }
            }
            if (sweep == 0) {
                for (int o = 16; o > 0; o >>= 1) my_edges |= __shfl_xor_sync(0xffffffffu, my_edges, o);
                if (lane == 0 && my_edges) atomicOr(reinterpret_cast<unsigned *>(&sMisc[2]), my_edges);
                __syncthreads();
                edge_bits = *reinterpret_cast<volatile unsigned *>(&sMisc[2]);
            }
            edge_bits |= extra_edges;                 // changes an abandoned light revisit already made on the tile's edges
            if (threadIdx.x == 0) cur[tile] = static_cast<uint8_t>(edge_bits);
            block_changed |= edge_bits != 0;
            if (sweep >= 3) {
                const long long tk4 = clock64();
                ph[0] += tk1 - tk0; ph[1] += tk2 - tk1; ph[2] += tk3 - tk2; ph[3] += tk4 - tk3; ph[4] += 1;
                if (threadIdx.x == 0 && sweep < 32) atomicMax(&p.st->dbg_maxclk[sweep], static_cast<unsigned int>(tk4 - tk0));
            }
    };
    stamp(0);
    for (;;) {
        prev = p.tile_changed + ((sweep + 1) & 1) * ntiles;
        cur = p.tile_changed + (sweep & 1) * ntiles;
        block_changed = false;
        if (threadIdx.x == 0) { sList[FL_LIST_MAX] = 0; sAbort[FL_LIST_MAX] = 0; }
        int n_light = 0;                              // uniform across the block
        if (sweep == 0) {
            // first visits cost between 1 and 3x the average (number of event rounds): blocks draw tiles from a counter
            for (;;) {
                __syncthreads();
                if (threadIdx.x == 0) sMisc[15] = static_cast<int>(atomicAdd(&p.st->tile_counter, 1u));
                __syncthreads();
                const int tile = sMisc[15];
                if (tile >= ntiles) break;
                heavy_visit(tile, true, 0u);
            }
        }
        for (int tile = blockIdx.x; sweep > 0 && tile < ntiles; tile += gridDim.x) {
            const int tyi = tile / tiles_x, txi = tile - tyi * tiles_x;
            // a tile is revisited only if a neighbour changed pixels on the edge that faces it ...
            const bool need = (txi > 0 && (prev[tile - 1] & EDGE_RIGHT)) || (txi + 1 < tiles_x && (prev[tile + 1] & EDGE_LEFT)) ||
                              (tyi > 0 && (prev[tile - tiles_x] & EDGE_BOTTOM)) || (tyi + 1 < tiles_y && (prev[tile + tiles_x] & EDGE_TOP));
            // ... or if its own light revisit ran out of queue space in the previous sweep
            const bool dirty = (prev[tile] & EDGE_DIRTY) != 0;
            if (!need && !dirty) {
                if (threadIdx.x == 0) cur[tile] = 0;
                continue;
            }
            if (!LAB32 && !dirty && n_light < FL_LIST_MAX) {
                // a revisit usually changes a few dozen pixels next to an edge: one WARP handles it straight on the
                // global state (below) instead of the block staging the whole 48 KB tile in shared memory
                if (threadIdx.x == 0) sList[n_light] = tile;
                ++n_light;
                continue;
            }
            heavy_visit(tile, dirty, 0u);
        }
        if (n_light > 0 && n_light < p.light_min_tiles) {
            // latency regime (the tail sweeps: a handful of tiles with long dependency chains): shared memory is faster
            __syncthreads();
            for (int li = 0; li < n_light; ++li) heavy_visit(sList[li], false, 0u);
            n_light = 0;
        }
        if (n_light > 0) {
            // throughput regime: many tiles with little work each, eight of them in flight per block
            __syncthreads();                          // the block-wide tile (if any) is done: its shared memory is free
            const int warp = threadIdx.x >> 5;
            unsigned char *wbase = smem + warp * LIGHT_WARP_BYTES;
            unsigned short *wq = reinterpret_cast<unsigned short *>(wbase);
            unsigned *wflag = reinterpret_cast<unsigned *>(wbase + 2 * LQ * 2);
            int *wc = reinterpret_cast<int *>(wbase + 2 * LQ * 2 + (CT * CT / 32) * 4);      // [0],[1] queue sizes, [2] edge bits, [3] overflow
            auto gstate = [&](int y, int x) -> unsigned long long {
                unsigned long long v = ~0ull;                  // ST_OUTSIDE
                if (y >= 0 && y < H && x >= 0 && x < W) v = *reinterpret_cast<volatile unsigned long long *>(&p.state[static_cast<size_t>(y) * W + x]);
                return v;
            };
            for (int li = warp; li < n_light; li += 8) {
                const int tile = sList[li];
                const int tyi = tile / tiles_x, txi = tile - tyi * tiles_x;
                const int x0 = txi * CT, y0 = tyi * CT;
                __syncwarp();
                for (int i = lane; i < CT * CT / 32; i += 32) wflag[i] = 0u;
                if (lane < 4) wc[lane] = 0;
                __syncwarp();
                if (lane == 0 && sweep < 32) atomicAdd(&p.st->dbg_tiles[sweep], 1u);
                // work list: floodable edge pixels next to a flooded pixel of the neighbouring tile (all sixteen loads of
                // a lane are issued before the first one is used)
                {
                    unsigned long long own[8], halo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int e = lane + 32 * j;
                        const int side = e >> 6, k = e & (CT - 1);
                        const int r = side == 0 ? 0 : (side == 1 ? CT - 1 : k), c = side <= 1 ? k : (side == 2 ? 0 : CT - 1);
                        const int hy = y0 + (side == 0 ? -1 : (side == 1 ? CT : k)), hx = x0 + (side <= 1 ? k : (side == 2 ? -1 : CT));
                        own[j] = gstate(y0 + r, x0 + c);
                        halo[j] = gstate(hy, hx);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int e = lane + 32 * j;
                        const int side = e >> 6, k = e & (CT - 1);
                        const int r = side == 0 ? 0 : (side == 1 ? CT - 1 : k), c = side <= 1 ? k : (side == 2 ? 0 : CT - 1);
                        if ((own[j] >> 16) & 1ull) continue;                       // marker, outside the mask or outside the image
                        if (static_cast<unsigned>(halo[j] >> 32) >= ORD_INF) continue;
                        const int idx = r * CT + c;
                        const unsigned bit = 1u << (idx & 31);
                        if (atomicOr(&wflag[idx >> 5], bit) & bit) continue;
                        const int pos = atomicAdd(&wc[0], 1);
                        if (pos < LQ) wq[pos] = static_cast<unsigned short>(idx); else wc[3] = 1;
                    }
                }
                __syncwarp();
                int qc = 0;
                int processed = 0;
                bool abandon = *reinterpret_cast<volatile int *>(&wc[0]) > p.light_init_max;
                for (; !abandon;) {
                    int n = *reinterpret_cast<volatile int *>(&wc[qc]);
                    if (n == 0) break;
                    if (n > LQ) n = LQ;
                    processed += n;
                    if (processed > p.light_items_max) {             // not a small revisit after all: finish it block-wide
                        abandon = true;
                        break;
                    }
                    const unsigned short *qin = wq + qc * LQ;
                    unsigned short *qout = wq + (qc ^ 1) * LQ;
                    for (int i = lane; i < n; i += 32) {
                        int idx = qin[i];
                        atomicAnd(&wflag[idx >> 5], ~(1u << (idx & 31)));
                        __threadfence_block();
                        const int CHASE = p.chase_light;
#pragma unroll 1
                        for (int step = 0; step <= CHASE; ++step) {
                            const int r = idx >> 6, c = idx & (CT - 1);
                            const int y = y0 + r, x = x0 + c;
                            const unsigned long long n0 = gstate(y - 1, x), n1 = gstate(y, x - 1), n2 = gstate(y, x + 1), n3 = gstate(y + 1, x);
                            const unsigned long long old = gstate(y, x);
                            const float v = p.img[static_cast<size_t>(y) * W + x];
                            const unsigned vo = ord_f32(p.negate ? -v : v);
                            unsigned long long best = n0;
                            int bi = 0;
                            if ((n1 >> 16) < (best >> 16)) { best = n1; bi = 1; }
                            if ((n2 >> 16) < (best >> 16)) { best = n2; bi = 2; }
                            if ((n3 >> 16) < (best >> 16)) { best = n3; bi = 3; }
                            const unsigned Lq = static_cast<unsigned>(best >> 32);
                            if (Lq >= ORD_INF) break;
                            unsigned long long ns;
                            if (vo > Lq) {
                                ns = (static_cast<unsigned long long>(vo) << 32) | (best & 0xFFFFull);
                            } else {
                                unsigned h = ((static_cast<unsigned>(best >> 16) & 0xFFFFu) >> 1) + 1u;
                                if (h > HOP_MAX) { h = HOP_MAX; overflow = true; }
                                ns = (static_cast<unsigned long long>(Lq) << 32) | (static_cast<unsigned long long>(h << 1) << 16) | (best & 0xFFFFull);
                            }
                            if (ns == old) break;
                            *reinterpret_cast<volatile unsigned long long *>(&p.state[static_cast<size_t>(y) * W + x]) = ns;
                            __threadfence_block();
                            const int ahead = 3 - bi;
                            bool w0 = r > 0 && !((n0 >> 16) & 1ull), w1 = c > 0 && !((n1 >> 16) & 1ull);
                            bool w2 = c < CT - 1 && !((n2 >> 16) & 1ull), w3 = r < CT - 1 && !((n3 >> 16) & 1ull);
                            const bool go = step < CHASE && (ahead == 0 ? w0 : (ahead == 1 ? w1 : (ahead == 2 ? w2 : w3)));
                            if (go) {
                                if (ahead == 0) w0 = false; else if (ahead == 1) w1 = false; else if (ahead == 2) w2 = false; else w3 = false;
                            }
                            const int i0 = idx - CT, i1 = idx - 1, i2 = idx + 1, i3 = idx + CT;
                            unsigned o0 = ~0u, o1 = ~0u, o2 = ~0u, o3 = ~0u;
                            if (w0) o0 = atomicOr(&wflag[i0 >> 5], 1u << (i0 & 31)) & (1u << (i0 & 31));
                            if (w1) o1 = atomicOr(&wflag[i1 >> 5], 1u << (i1 & 31)) & (1u << (i1 & 31));
                            if (w2) o2 = atomicOr(&wflag[i2 >> 5], 1u << (i2 & 31)) & (1u << (i2 & 31));
                            if (w3) o3 = atomicOr(&wflag[i3 >> 5], 1u << (i3 & 31)) & (1u << (i3 & 31));
                            const int cnt = (o0 == 0) + (o1 == 0) + (o2 == 0) + (o3 == 0);
                            if (cnt) {
                                // every reserved slot below LQ is written by its owner; slots beyond are dropped and the
                                // tile gets a full block-wide visit next sweep
                                int pos = atomicAdd(&wc[qc ^ 1], cnt);
                                if (pos + cnt > LQ) wc[3] = 1;
                                if (o0 == 0) { if (pos < LQ) qout[pos] = static_cast<unsigned short>(i0); ++pos; }
                                if (o1 == 0) { if (pos < LQ) qout[pos] = static_cast<unsigned short>(i1); ++pos; }
                                if (o2 == 0) { if (pos < LQ) qout[pos] = static_cast<unsigned short>(i2); ++pos; }
                                if (o3 == 0) { if (pos < LQ) qout[pos] = static_cast<unsigned short>(i3); ++pos; }
                            }
                            const unsigned eb = (r == 0 ? EDGE_TOP : 0u) | (r == CT - 1 ? EDGE_BOTTOM : 0u) | (c == 0 ? EDGE_LEFT : 0u) |
                                                (c == CT - 1 ? EDGE_RIGHT : 0u);
                            if (eb) atomicOr(reinterpret_cast<unsigned *>(&wc[2]), eb);
                            if (!go) break;
                            idx = ahead == 0 ? i0 : (ahead == 1 ? i1 : (ahead == 2 ? i2 : i3));
                        }
                    }
                    __syncwarp();
                    if (lane == 0) {
                        // entries beyond the queue were dropped (flagged): clamp so the next round reads valid slots only
                        if (*reinterpret_cast<volatile int *>(&wc[qc ^ 1]) > LQ) wc[qc ^ 1] = LQ;
                        wc[qc] = 0;
                    }
                    qc ^= 1;
                    __syncwarp();
                }
                if (lane == 0) {
                    const unsigned edges = static_cast<unsigned>(*reinterpret_cast<volatile int *>(&wc[2])) & 0xFu;
                    if (abandon || *reinterpret_cast<volatile int *>(&wc[3])) {
                        // hand the tile to the block (this sweep): tile index + the edge bits already produced
                        const int pos = atomicAdd(&sAbort[FL_LIST_MAX], 1);
                        sAbort[pos] = tile | static_cast<int>(edges << 28);
                    } else {
                        cur[tile] = static_cast<uint8_t>(edges);
                        if (edges) *reinterpret_cast<volatile int *>(&sList[FL_LIST_MAX]) = 1;
                    }
                }
            }
            __syncthreads();
            if (*reinterpret_cast<volatile int *>(&sList[FL_LIST_MAX]) != 0) block_changed = true;
        }
        {
            // light revisits that turned out to be large were abandoned half way: finish them block-wide (full scan)
            __syncthreads();
            const int n_abort = *reinterpret_cast<volatile int *>(&sAbort[FL_LIST_MAX]);
            for (int h = 0; h < n_abort; ++h) heavy_visit(sAbort[h] & 0x0FFFFFFF, true, static_cast<unsigned>(sAbort[h]) >> 28);
        }
        if (block_changed && threadIdx.x == 0) p.st->changed[sweep % 3] = 1;
        __threadfence();
        grid.sync();
        const unsigned int flag = *reinterpret_cast<volatile unsigned int *>(&p.st->changed[sweep % 3]);
        if (blockIdx.x == 0 && threadIdx.x == 0) p.st->changed[(sweep + 2) % 3] = 0;
        ++sweep;
        stamp(sweep < 34 ? sweep : 34);
        if (!flag) break;
    }
    if (overflow) atomicExch(&p.st->overflow, 1u);
    if (threadIdx.x == 0 && ph[4])
        for (int k = 0; k < 5; ++k) atomicAdd(&p.st->dbg_phase[k], static_cast<unsigned long long>(ph[k]));
    if (blockIdx.x == 0 && threadIdx.x == 0) p.st->sweeps = static_cast<unsigned int>(sweep);
    // ---- final phase: labels out + order-independence check (see the file header)
    int bad_total = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sMisc[15] = static_cast<int>(atomicAdd(&p.st->tile_counter2, 1u));
        __syncthreads();
        const int tile = sMisc[15];
        if (tile >= ntiles) break;
        const int tyi = tile / tiles_x, txi = tile - tyi * tiles_x;
        const int x0 = txi * CT, y0 = tyi * CT;
        {
            const int pc = threadIdx.x & (CT - 1), pr0 = threadIdx.x >> 6;
            load_halo(x0, y0);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                unsigned long long sv[8];
                int lv[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int y = y0 + (h * 8 + k) * 4 + pr0, x = x0 + pc;
                    const bool in = y < H && x < W;
                    sv[k] = in ? p.state[static_cast<size_t>(y) * W + x] : ST_OUTSIDE;
                    if (LAB32) lv[k] = in ? p.lab32[static_cast<size_t>(y) * W + x] : 0;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    sS[(h * 8 + k) * 4 + pr0 + 1][pc + 1] = sv[k];
                    if (LAB32) sLab[(h * 8 + k) * 4 + pr0 + 1][pc + 1] = lv[k];
                }
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int rr = i * 4 + (threadIdx.x >> 6), cc = threadIdx.x & (CT - 1);       // a warp = 32 adjacent columns
            const int y = y0 + rr, x = x0 + cc;
            if (y >= H || x >= W) continue;
            const int r = rr + 1, c = cc + 1;
            const unsigned long long s = sS[r][c];
            const bool flooded = static_cast<unsigned>(s >> 32) < ORD_INF;
            const int me = LAB32 ? sLab[r][c] : static_cast<int>(s & 0xFFFFull);
            const size_t o = static_cast<size_t>(y) * W + x;
            if (LAB32) p.out32[o] = flooded ? me : 0;
            else p.out16[o] = flooded ? static_cast<uint16_t>(me) : static_cast<uint16_t>(0);
            if (flooded && !((s >> 16) & 1ull)) {
                const unsigned long long n0 = sS[r - 1][c], n1 = sS[r][c - 1], n2 = sS[r][c + 1], n3 = sS[r + 1][c];
                unsigned lmin = static_cast<unsigned>(n0 >> 32);
                lmin = min(lmin, static_cast<unsigned>(n1 >> 32));
                lmin = min(lmin, static_cast<unsigned>(n2 >> 32));
                lmin = min(lmin, static_cast<unsigned>(n3 >> 32));
                bool bad = false;
                if (static_cast<unsigned>(n0 >> 32) == lmin && (LAB32 ? sLab[r - 1][c] : static_cast<int>(n0 & 0xFFFFull)) != me) bad = true;
                if (static_cast<unsigned>(n1 >> 32) == lmin && (LAB32 ? sLab[r][c - 1] : static_cast<int>(n1 & 0xFFFFull)) != me) bad = true;
                if (static_cast<unsigned>(n2 >> 32) == lmin && (LAB32 ? sLab[r][c + 1] : static_cast<int>(n2 & 0xFFFFull)) != me) bad = true;
                if (static_cast<unsigned>(n3 >> 32) == lmin && (LAB32 ? sLab[r + 1][c] : static_cast<int>(n3 & 0xFFFFull)) != me) bad = true;
                if (bad) {
                    ++bad_total;
                    if (p.badlist) {
                        const unsigned pos = atomicAdd(&p.st->n_bad, 1u);
                        if (pos < static_cast<unsigned>(p.bad_cap)) p.badlist[pos] = static_cast<int>(o);
                    }
                }
            }
        }
    }
    if (bad_total) atomicAdd(&p.st->ambiguous, static_cast<unsigned int>(bad_total));
    __syncthreads();
    stamp(35);          // block 0's end of the final phase
}

// ------------------------------------------------------------------------------------------
// Exact fallback of the tiled path, one cooperative launch that returns at once unless the flood flagged an
// order-dependent pixel.  The flood never leaves a 4-connected component of the mask, so only the components that
// hold a flagged pixel are re-flooded, each by ONE thread running skimage's algorithm (binary heap keyed (value, age))
// on its own heap region; every other pixel keeps the parallel result.  A component is independent of the rest of
// the image as long as its pop order does not depend on the heap's internal order for equal keys: that only happens
// among marker pixels (age 0) of equal value, and only matters if they carry different labels.  Such a pair pops
// consecutively, so it is detected during the component's own flood; then (and if the hop counter overflowed or the
// bad-pixel list was too small) the whole image is re-flooded by the single-thread restatement, exact in every case.
// ------------------------------------------------------------------------------------------
__device__ void sequential_flood(const float *img, int negate, const int *markers, const uint8_t *mask, int H, int W, int *lab,
                                 HeapItem *heap, uint16_t *out16) {
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

struct FallbackParams {
    const uint16_t *L16;
    const int *marker_at_root;
    const float *img;
    int negate;
    float th_mask;
    int use_th_mask;
    int H, W;
    Stats *st;
    const int *badlist;
    int bad_cap;
    int *markers;        // [n]
    uint8_t *mask8;      // [n]
    int *uf;             // [n] union-find of the 4-connected mask components
    int *comp_area;      // [n] at roots: pixels of the component
    int *comp_base;      // [n] at roots: start of the component's heap region
    int *comp_cnt;       // [n] at roots: marker pixels scattered into the heap region
    int *comp_flag;      // [n] at roots: holds an order-dependent pixel
    int *lab;            // [n]
    HeapItem *heap;      // [n]
    uint16_t *out16;
};

__global__ void __launch_bounds__(256)
fallback_kernel(const FallbackParams p) {
    namespace cg = cooperative_groups;
    Stats *st = p.st;
    if (st->ambiguous == 0 && st->overflow == 0) return;            // uniform across the grid: written by the previous launch
    cg::grid_group grid = cg::this_grid();
    const int H = p.H, W = p.W;
    const long long n = static_cast<long long>(H) * W;
    const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long gsz = static_cast<long long>(gridDim.x) * blockDim.x;
    // P1: explicit marker / mask images, union-find init
    for (long long i = gtid; i < n; i += gsz) {
        const int y = static_cast<int>(i / W), x = static_cast<int>(i - static_cast<long long>(y) * W);
        const unsigned v16 = p.L16[i];
        const bool m = p.use_th_mask ? (p.img[i] > p.th_mask) : ((v16 & L16_MASK) != 0);
        const int mk = (m && (v16 & L16_SEED)) ? marker_of(v16, y, x, W, p.marker_at_root) : 0;
        p.mask8[i] = m ? 1 : 0;
        p.markers[i] = mk;
        p.lab[i] = mk;
        p.uf[i] = m ? static_cast<int>(i) : -1;
        p.comp_area[i] = 0;
        p.comp_cnt[i] = 0;
        p.comp_flag[i] = 0;
    }
    __threadfence();
    grid.sync();
    // P2: 4-connected unions
    for (long long i = gtid; i < n; i += gsz) {
        if (!p.mask8[i]) continue;
        const int y = static_cast<int>(i / W), x = static_cast<int>(i - static_cast<long long>(y) * W);
        if (x > 0 && p.mask8[i - 1]) uf_union(p.uf, static_cast<int>(i), static_cast<int>(i - 1));
        if (y > 0 && p.mask8[i - W]) uf_union(p.uf, static_cast<int>(i), static_cast<int>(i - W));
    }
    __threadfence();
    grid.sync();
    // P3: flatten, component sizes
    for (long long i = gtid; i < n; i += gsz) {
        if (!p.mask8[i]) continue;
        const int r = uf_find_v(p.uf, static_cast<int>(i));
        p.uf[i] = r;
        atomicAdd(&p.comp_area[r], 1);
    }
    __threadfence();
    grid.sync();
    // P4: components that hold an order-dependent pixel
    {
        unsigned nb = *reinterpret_cast<volatile unsigned *>(&st->n_bad);
        if (gtid == 0 && (nb > static_cast<unsigned>(p.bad_cap) || *reinterpret_cast<volatile unsigned *>(&st->overflow))) st->need_global = 1;
        if (nb > static_cast<unsigned>(p.bad_cap)) nb = static_cast<unsigned>(p.bad_cap);
        for (long long i = gtid; i < nb; i += gsz) {
            const int b = p.badlist[i];
            const int r = uf_find_v(p.uf, b);
            p.comp_flag[r] = 1;
        }
    }
    __threadfence();
    grid.sync();
    // P5: heap regions (bump allocation: the regions of disjoint components add up to at most n entries)
    for (long long i = gtid; i < n; i += gsz) {
        if (*reinterpret_cast<volatile int *>(&p.uf[i]) == static_cast<int>(i) && *reinterpret_cast<volatile int *>(&p.comp_flag[i])) {
            p.comp_base[i] = static_cast<int>(atomicAdd(&st->pool_top, static_cast<unsigned>(*reinterpret_cast<volatile int *>(&p.comp_area[i]))));
            atomicAdd(&st->n_reflooded, 1u);
        }
    }
    __threadfence();
    grid.sync();
    // P6: the markers of the flagged components go into their heap regions (any order: see the header comment);
    // every other pixel of such a component becomes unlabelled again
    for (long long i = gtid; i < n; i += gsz) {
        if (!p.mask8[i]) continue;
        const int r = *reinterpret_cast<volatile int *>(&p.uf[i]);
        if (!*reinterpret_cast<volatile int *>(&p.comp_flag[r])) continue;
        const int mk = p.markers[i];
        if (mk) {
            const int slot = atomicAdd(&p.comp_cnt[r], 1);
            HeapItem e;
            e.value = flood_value(p.img, static_cast<int>(i), p.negate);
            e.age = 0;
            e.index = static_cast<int>(i);
            p.heap[*reinterpret_cast<volatile int *>(&p.comp_base[r]) + slot] = e;
        } else {
            p.out16[i] = 0;
        }
    }
    __threadfence();
    grid.sync();
    // P7: one thread per flagged component: heapify, then skimage's flood
    for (long long i = gtid; i < n; i += gsz) {
        if (*reinterpret_cast<volatile int *>(&p.uf[i]) != static_cast<int>(i) || !*reinterpret_cast<volatile int *>(&p.comp_flag[i])) continue;
        HeapItem *h = p.heap + *reinterpret_cast<volatile int *>(&p.comp_base[i]);
        int hn = *reinterpret_cast<volatile int *>(&p.comp_cnt[i]);
        for (int s0 = hn / 2 - 1; s0 >= 0; --s0) {                  // bottom-up heap construction
            int k = s0;
            for (;;) {
                const int l = 2 * k + 1, r = 2 * k + 2;
                int sm = k;
                if (l < hn && item_smaller(h[l], h[sm])) sm = l;
                if (r < hn && item_smaller(h[r], h[sm])) sm = r;
                if (sm == k) break;
                const HeapItem t = h[k];
                h[k] = h[sm];
                h[sm] = t;
                k = sm;
            }
        }
        int age = 1;
        HeapItem e, ne;
        bool conflict = false;
        while (hn > 0) {
            heap_pop(h, hn, e);
            const int l = p.lab[e.index];
            if (e.age == 0 && hn > 0 && h[0].age == 0 && h[0].value == e.value && p.lab[h[0].index] != l) conflict = true;
            const int y = e.index / W, x = e.index - y * W;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int q;
                if (k == 0) { if (y == 0) continue; q = e.index - W; }
                else if (k == 1) { if (x == 0) continue; q = e.index - 1; }
                else if (k == 2) { if (x + 1 >= W) continue; q = e.index + 1; }
                else { if (y + 1 >= H) continue; q = e.index + W; }
                if (!p.mask8[q]) continue;
                if (p.lab[q]) continue;
                age += 1;
                p.lab[q] = l;
                p.out16[q] = static_cast<uint16_t>(static_cast<unsigned int>(l));
                ne.value = flood_value(p.img, q, p.negate);
                ne.age = age;
                ne.index = q;
                heap_push(h, hn, ne);
            }
        }
        if (conflict) st->need_global = 1;
    }
    __threadfence();
    grid.sync();
    // P8: last resort
    if (gtid == 0 && *reinterpret_cast<volatile unsigned *>(&st->need_global))
        sequential_flood(p.img, p.negate, p.markers, p.mask8, H, W, p.lab, p.heap, p.out16);
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

static size_t legacy_postproc_workspace_bytes(int H, int W) {
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

static int legacy_pp_watershed(const float *image, const int32_t *markers, const uint8_t *mask, int H, int W,
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

static int legacy_distance_postprocessing(const float *border, const float *cell, int H, int W, int ld, float th_seed,
                                          float th_cell, uint16_t *out, void *workspace, size_t workspace_bytes,
                                          int64_t *info_host, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t n = static_cast<size_t>(H) * W;
    MBS_REQUIRE(H > 0 && W > 0 && ld >= W && n < (1ull << 31), "distance_postprocessing: bad shape H=%d W=%d ld=%d", H,
                W, ld);
    MBS_REQUIRE(workspace_bytes >= legacy_postproc_workspace_bytes(H, W),
                "distance_postprocessing: workspace too small (%zu < %zu)", workspace_bytes,
                legacy_postproc_workspace_bytes(H, W));
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

static int legacy_boundary_postprocessing(const float *prediction_hwc, int H, int W, uint16_t *out, void *workspace,
                                          size_t workspace_bytes, int64_t *info_host, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t n = static_cast<size_t>(H) * W;
    MBS_REQUIRE(H > 0 && W > 0 && n < (1ull << 31), "boundary_postprocessing: bad shape H=%d W=%d", H, W);
    MBS_REQUIRE(workspace_bytes >= legacy_postproc_workspace_bytes(H, W), "boundary_postprocessing: workspace too small");
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


// ==========================================================================================
// Tiled pipeline: host side
// ==========================================================================================
namespace {

bool legacy_path() {        // MBS_PP_LEGACY=1: the round-1 streaming pipeline (A/B runs)
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MBS_PP_LEGACY");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

int light_limit(int which) {      // MBS_PP_LIGHT=init,items,chase_heavy,chase_light,min_tiles (A/B knob); 0,0 = no warp-level revisits
    static int v[5] = {-1, -1, -1, -1, -1};
    if (v[0] < 0) {
        int a = 64, b = 256, c = 16, d = 8, m = 4;
        const char *e = getenv("MBS_PP_LIGHT");
        if (e) sscanf(e, "%d,%d,%d,%d,%d", &a, &b, &c, &d, &m);
        v[1] = b;
        v[2] = c;
        v[3] = d;
        v[4] = m;
        v[0] = a;
    }
    return v[which];
}

struct TiledWs {
    Stats *st;
    float *cell_s;
    uint16_t *L16;
    int *G, *area, *list;
    int list_cap;
    unsigned *bitmap;
    int *prefix, *block_sums;
    unsigned long long *state;
    uint8_t *tile_changed;
    int *lab32;              // generic watershed only
    int *comp_cnt, *comp_flag;   // exact fallback: per-component counters (at the component roots)
    // exact sequential fallback
    int *markers;
    uint8_t *mask8;
    int *lab;
    HeapItem *heap;
};
constexpr int MID_MAX_BLOCKS = 4096;

size_t tiled_ws_bytes(size_t n, int H, int W, bool generic) {
    const size_t tiles = static_cast<size_t>((H + CT - 1) / CT) * ((W + CT - 1) / CT);
    size_t b = r256(sizeof(Stats)) + r256(n * 4) + r256(n * 2) + 2 * r256(n * 4) + r256((n / 4 + 1024) * 4) +
               2 * r256((n / 32 + 2) * 4) + r256(MID_MAX_BLOCKS * 4) + r256(n * 8) + r256(2 * tiles + 256) +
               r256(n * 4) /*markers*/ + r256(n) /*mask*/ + r256(n * 4) /*lab*/ + r256(n * sizeof(HeapItem)) +
               2 * r256(n * 4) /*component counters / flags of the exact fallback (area / base live in the state array)*/;
    if (generic) b += r256(n * 4);
    return b + 4096;
}

bool carve_tiled(TiledWs &t, void *workspace, size_t bytes, size_t n, int H, int W, bool generic) {
    const size_t tiles = static_cast<size_t>((H + CT - 1) / CT) * ((W + CT - 1) / CT);
    Carver cv{static_cast<char *>(workspace), bytes};
    t.st = cv.take<Stats>(1);
    t.cell_s = cv.take<float>(n);
    t.L16 = cv.take<uint16_t>(n);
    t.G = cv.take<int>(n);
    t.area = cv.take<int>(n);
    t.list_cap = static_cast<int>(n / 4 + 1024);
    t.list = cv.take<int>(t.list_cap);
    t.bitmap = cv.take<unsigned>(n / 32 + 2);
    t.prefix = cv.take<int>(n / 32 + 2);
    t.block_sums = cv.take<int>(MID_MAX_BLOCKS);
    t.state = cv.take<unsigned long long>(n);
    t.tile_changed = cv.take<uint8_t>(2 * tiles + 256);
    t.markers = cv.take<int>(n);
    t.mask8 = cv.take<uint8_t>(n);
    t.lab = cv.take<int>(n);
    t.heap = cv.take<HeapItem>(n);
    t.comp_cnt = cv.take<int>(n);
    t.comp_flag = cv.take<int>(n);
    t.lab32 = generic ? cv.take<int>(n) : nullptr;
    return cv.ok;
}

struct CoopCfg {
    int mid_blocks, flood16_blocks, flood32_blocks, fallback_blocks;
    bool ready;
};

int coop_config(CoopCfg **out) {
    static CoopCfg cfg[mbs::kMaxDevices] = {};
    const int dev = mbs::current_device();
    CoopCfg &c = cfg[dev];
    if (!c.ready) {
        int sms = 0, per = 0;
        MBS_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        MBS_CHECK_CUDA(cudaFuncSetAttribute(flood_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FLOOD_SMEM16));
        MBS_CHECK_CUDA(cudaFuncSetAttribute(flood_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FLOOD_SMEM32));
        MBS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, mid_kernel, 256, 0));
        // the phases are latency bound; measured (config 3 / 2048^2 maps, whole pipeline): 1184 blocks 0.675 / 0.268 ms,
        // 592 blocks 0.674 / 0.255 ms, 296 blocks 0.703 / 0.264 ms, 148 blocks 0.768 / 0.275 ms
        c.mid_blocks = sms * (per > 4 ? 4 : (per > 0 ? per : 1));
        if (c.mid_blocks > MID_MAX_BLOCKS) c.mid_blocks = MID_MAX_BLOCKS;
        if (const char *e = getenv("MBS_MID_BLOCKS")) {       // A/B knob
            const int v = atoi(e);
            if (v >= 1 && v <= c.mid_blocks) c.mid_blocks = v;
        }
        MBS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, flood_kernel<false>, 256, FLOOD_SMEM16));
        c.flood16_blocks = sms * (per > 0 ? per : 1);
        MBS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, flood_kernel<true>, 256, FLOOD_SMEM32));
        c.flood32_blocks = sms * (per > 0 ? per : 1);
        MBS_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, fallback_kernel, 256, 0));
        c.fallback_blocks = sms * (per > 0 ? (per > 4 ? 4 : per) : 1);
        c.ready = true;
    }
    *out = &c;
    return 0;
}

int read_info(const TiledWs &t, int64_t *info_host, cudaStream_t stream) {
    Stats *hp = pinned_stats();
    MBS_REQUIRE(hp != nullptr, "post-processing: cannot allocate pinned host memory");
    MBS_CHECK_CUDA(cudaMemcpyAsync(hp, t.st, sizeof(Stats), cudaMemcpyDeviceToHost, stream));
    MBS_CHECK_CUDA(cudaStreamSynchronize(stream));
    memset(info_host, 0, 8 * sizeof(int64_t));
    info_host[0] = hp->n_comp;
    info_host[1] = hp->n_markers;
    info_host[2] = hp->sweeps;
    info_host[3] = (hp->ambiguous || hp->overflow) ? 1 : 0;
    info_host[4] = hp->ambiguous;
    info_host[5] = hp->overflow;
    info_host[6] = hp->need_global;          // whole-image sequential flood ran
    info_host[7] = hp->n_reflooded;          // mask components re-flooded on their own
    return 0;
}

// exact fallback of the fused paths (returns at once on the device unless the flood flagged an order-dependent pixel)
int launch_fallback(const TiledWs &t, const CoopCfg *cfg, int H, int W, int negate, float th_mask, int use_th_mask, uint16_t *out,
                    cudaStream_t stream) {
    const size_t n = static_cast<size_t>(H) * W;
    FallbackParams fb;
    fb.L16 = t.L16;
    fb.marker_at_root = t.area;
    fb.img = t.cell_s;
    fb.negate = negate;
    fb.th_mask = th_mask;
    fb.use_th_mask = use_th_mask;
    fb.H = H;
    fb.W = W;
    fb.st = t.st;
    fb.badlist = t.list;
    fb.bad_cap = t.list_cap;
    fb.markers = t.markers;
    fb.mask8 = t.mask8;
    fb.uf = t.G;                                                    // the seed union-find is dead after mid_kernel
    fb.comp_area = reinterpret_cast<int *>(t.state);                // the flood state is dead after its final phase
    fb.comp_base = reinterpret_cast<int *>(t.state) + n;
    fb.comp_cnt = t.comp_cnt;
    fb.comp_flag = t.comp_flag;
    fb.lab = t.lab;
    fb.heap = t.heap;
    fb.out16 = out;
    void *args[] = {(void *)&fb};
    MBS_CHECK_CUDA(cudaLaunchCooperativeKernel((void *)fallback_kernel, dim3(cfg->fallback_blocks), dim3(256), args, 0, stream));
    mbs::count_launch();
    return 0;
}

// front end + labelling + flood for both methods; `a` / `b`: border + cell maps (distance) or the (H,W,3) probabilities
template <bool BOUNDARY>
int run_tiled(const float *a, const float *b, int H, int W, int ld, float th_seed, float th_cell, uint16_t *out,
              void *workspace, size_t workspace_bytes, int64_t *info_host, cudaStream_t stream) {
    const size_t n = static_cast<size_t>(H) * W;
    TiledWs t;
    MBS_REQUIRE(carve_tiled(t, workspace, workspace_bytes, n, H, W, false), "post-processing: workspace carve failed");
    CoopCfg *cfg = nullptr;
    int rc = coop_config(&cfg);
    if (rc) return rc;
    MBS_CHECK_CUDA(cudaMemsetAsync(t.st, 0, sizeof(Stats), stream));
    dim3 tgrid(mbs::cdiv(W, CT), mbs::cdiv(H, CT));
    front_ccl_kernel<BOUNDARY><<<tgrid, 256, 0, stream>>>(a, b, H, W, ld, th_seed, th_cell, t.cell_s, t.L16, t.G, t.area, t.list,
                                                          t.list_cap, t.bitmap, t.st);
    MBS_CHECK_LAUNCH();
    {
        const int use_mean = BOUNDARY ? 0 : 1;       // boundary method: drop area <= 4 only (postprocessing.py:79-84)
        const uint16_t *L16 = t.L16;
        const int *list = t.list;
        void *args[] = {(void *)&L16, (void *)&H, (void *)&W, (void *)&t.G, (void *)&t.area, (void *)&list, (void *)&t.list_cap,
                        (void *)&t.bitmap, (void *)&t.prefix, (void *)&t.block_sums, (void *)&t.st, (void *)&use_mean};
        MBS_CHECK_CUDA(cudaLaunchCooperativeKernel((void *)mid_kernel, dim3(cfg->mid_blocks), dim3(256), args, 0, stream));
        mbs::count_launch();
    }
    {
        FloodParams fp;
        fp.img = t.cell_s;
        fp.negate = BOUNDARY ? 0 : 1;
        fp.L16 = t.L16;
        fp.marker_at_root = t.area;
        fp.th_mask = 0.0f;
        fp.use_th_mask = 0;
        fp.light_init_max = light_limit(0);
        fp.light_items_max = light_limit(1);
        fp.chase_heavy = light_limit(2);
        fp.chase_light = light_limit(3);
        fp.light_min_tiles = light_limit(4);
        fp.badlist = t.list;        // the root list is dead after mid_kernel
        fp.bad_cap = t.list_cap;
        fp.state = t.state;
        fp.lab32 = nullptr;
        fp.H = H;
        fp.W = W;
        fp.st = t.st;
        fp.tile_changed = t.tile_changed;
        fp.out16 = out;
        fp.out32 = nullptr;
        const int ntiles = static_cast<int>(tgrid.x * tgrid.y);
        const int blocks = ntiles < cfg->flood16_blocks ? ntiles : cfg->flood16_blocks;
        void *args[] = {(void *)&fp};
        MBS_CHECK_CUDA(cudaLaunchCooperativeKernel((void *)flood_kernel<false>, dim3(blocks), dim3(256), args, FLOOD_SMEM16, stream));
        mbs::count_launch();
    }
    rc = launch_fallback(t, cfg, H, W, BOUNDARY ? 0 : 1, 0.0f, 0, out, stream);
    if (rc) return rc;
    if (info_host) return read_info(t, info_host, stream);
    return 0;
}

}  // namespace

namespace {
__global__ void flood_reset_kernel(Stats *st) {
    st->changed[0] = st->changed[1] = st->changed[2] = 0;
    st->ambiguous = 0;
    st->sweeps = 0;
    st->overflow = 0;
    st->n_bad = st->need_global = st->n_reflooded = st->pool_top = 0;
    st->tile_counter = st->tile_counter2 = 0;
    for (int i = 0; i < 32; ++i) st->dbg_tiles[i] = st->dbg_rounds[i] = st->dbg_maxrounds[i] = st->dbg_items[i] = 0;
}
}  // namespace

// Threshold sweep of the evaluation (src/evaluation/eval.py:128-129, 395-412): the reference post-processes ONE
// prediction with every pair of product(th_cell, th_seed).  The smoothed map, the seed image, its labelling and the area
// filter depend on th_seed only, so they run once per seed threshold; each cell threshold then costs one flood.
extern "C" int mbs_distance_postprocessing_sweep(const float *border, const float *cell, int H, int W, int ld, const float *th_seeds_host,
                                                 int n_seed, const float *th_cells_host, int n_cell, uint16_t *out, void *workspace,
                                                 size_t workspace_bytes, int64_t *info_host, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t n = static_cast<size_t>(H) * W;
    MBS_REQUIRE(H > 0 && W > 0 && ld >= W && n < (1ull << 31) && n_seed > 0 && n_cell > 0 && th_seeds_host && th_cells_host,
                "distance_postprocessing_sweep: bad arguments");
    MBS_REQUIRE(workspace_bytes >= tiled_ws_bytes(n, H, W, false), "distance_postprocessing_sweep: workspace too small");
    TiledWs t;
    MBS_REQUIRE(carve_tiled(t, workspace, workspace_bytes, n, H, W, false), "distance_postprocessing_sweep: workspace carve failed");
    CoopCfg *cfg = nullptr;
    int rc = coop_config(&cfg);
    if (rc) return rc;
    dim3 tgrid(mbs::cdiv(W, CT), mbs::cdiv(H, CT));
    const int ntiles = static_cast<int>(tgrid.x * tgrid.y);
    for (int is = 0; is < n_seed; ++is) {
        MBS_CHECK_CUDA(cudaMemsetAsync(t.st, 0, sizeof(Stats), stream));
        front_ccl_kernel<false><<<tgrid, 256, 0, stream>>>(border, cell, H, W, ld, th_seeds_host[is], th_cells_host[0], t.cell_s, t.L16, t.G,
                                                           t.area, t.list, t.list_cap, t.bitmap, t.st);
        MBS_CHECK_LAUNCH();
        {
            const int use_mean = 1;
            const uint16_t *L16 = t.L16;
            const int *list = t.list;
            void *args[] = {(void *)&L16, (void *)&H, (void *)&W, (void *)&t.G, (void *)&t.area, (void *)&list, (void *)&t.list_cap,
                            (void *)&t.bitmap, (void *)&t.prefix, (void *)&t.block_sums, (void *)&t.st, (void *)&use_mean};
            MBS_CHECK_CUDA(cudaLaunchCooperativeKernel((void *)mid_kernel, dim3(cfg->mid_blocks), dim3(256), args, 0, stream));
            mbs::count_launch();
        }
        for (int ic = 0; ic < n_cell; ++ic) {
            uint16_t *o = out + (static_cast<size_t>(is) * n_cell + ic) * n;
            if (ic > 0) {
                flood_reset_kernel<<<1, 1, 0, stream>>>(t.st);
                MBS_CHECK_LAUNCH();
            }
            FloodParams fp;
            fp.img = t.cell_s;
            fp.negate = 1;
            fp.L16 = t.L16;
            fp.marker_at_root = t.area;
            fp.th_mask = th_cells_host[ic];
            fp.use_th_mask = 1;
            fp.light_init_max = light_limit(0);
            fp.light_items_max = light_limit(1);
            fp.chase_heavy = light_limit(2);
            fp.chase_light = light_limit(3);
            fp.light_min_tiles = light_limit(4);
            fp.badlist = t.list;
            fp.bad_cap = t.list_cap;
            fp.state = t.state;
            fp.lab32 = nullptr;
            fp.H = H;
            fp.W = W;
            fp.st = t.st;
            fp.tile_changed = t.tile_changed;
            fp.out16 = o;
            fp.out32 = nullptr;
            const int blocks = ntiles < cfg->flood16_blocks ? ntiles : cfg->flood16_blocks;
            void *args[] = {(void *)&fp};
            MBS_CHECK_CUDA(cudaLaunchCooperativeKernel((void *)flood_kernel<false>, dim3(blocks), dim3(256), args, FLOOD_SMEM16, stream));
            mbs::count_launch();
            rc = launch_fallback(t, cfg, H, W, 1, th_cells_host[ic], 1, o, stream);
            if (rc) return rc;
            if (info_host) {
                rc = read_info(t, info_host + (static_cast<size_t>(is) * n_cell + ic) * 8, stream);
                if (rc) return rc;
            }
        }
    }
    return 0;
}

// softmax over the three class planes of the boundary net + crop + channel-last layout in one pass (replaces
// F.softmax(prediction, dim=1)[..., pads:, pads:].permute(...) of infer.py:371-374): exp(z - max) / sum as torch evaluates it
namespace {
__global__ void softmax3_hwc_kernel(const float *__restrict__ logits, size_t plane, int ld, int y0, int x0, int H, int W,
                                    float *__restrict__ prob) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t i = static_cast<size_t>(y0 + y) * ld + x0 + x;
    const float z0 = logits[i], z1 = logits[plane + i], z2 = logits[2 * plane + i];
    const float m = fmaxf(z0, fmaxf(z1, z2));
    const float e0 = expf(z0 - m), e1 = expf(z1 - m), e2 = expf(z2 - m);
    const float sum = (e0 + e1) + e2;
    float *o = prob + (static_cast<size_t>(y) * W + x) * 3;
    o[0] = e0 / sum;
    o[1] = e1 / sum;
    o[2] = e2 / sum;
}
}  // namespace

extern "C" int mbs_softmax3_hwc(const float *logits, size_t plane_stride, int ld, int y0, int x0, int H, int W, float *prob,
                                void *stream_) {
    MBS_REQUIRE(logits && prob && H > 0 && W > 0 && ld >= x0 + W && y0 >= 0 && x0 >= 0, "softmax3_hwc: bad arguments");
    softmax3_hwc_kernel<<<dim3(mbs::cdiv(W, 256), H), 256, 0, static_cast<cudaStream_t>(stream_)>>>(logits, plane_stride, ld, y0, x0, H, W, prob);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t mbs_postproc_workspace_bytes(int H, int W) {
    const size_t n = static_cast<size_t>(H) * W;
    const size_t a = tiled_ws_bytes(n, H, W, true), b = legacy_postproc_workspace_bytes(H, W);
    return a > b ? a : b;
}

extern "C" int mbs_distance_postprocessing(const float *border, const float *cell, int H, int W, int ld, float th_seed,
                                           float th_cell, uint16_t *out, void *workspace, size_t workspace_bytes,
                                           int64_t *info_host, void *stream_) {
    if (legacy_path())
        return legacy_distance_postprocessing(border, cell, H, W, ld, th_seed, th_cell, out, workspace, workspace_bytes, info_host, stream_);
    const size_t n = static_cast<size_t>(H) * W;
    MBS_REQUIRE(H > 0 && W > 0 && ld >= W && n < (1ull << 31), "distance_postprocessing: bad shape H=%d W=%d ld=%d", H, W, ld);
    MBS_REQUIRE(workspace_bytes >= mbs_postproc_workspace_bytes(H, W), "distance_postprocessing: workspace too small (%zu < %zu)",
                workspace_bytes, mbs_postproc_workspace_bytes(H, W));
    return run_tiled<false>(border, cell, H, W, ld, th_seed, th_cell, out, workspace, workspace_bytes, info_host,
                            static_cast<cudaStream_t>(stream_));
}

extern "C" int mbs_boundary_postprocessing(const float *prediction_hwc, int H, int W, uint16_t *out, void *workspace,
                                           size_t workspace_bytes, int64_t *info_host, void *stream_) {
    if (legacy_path()) return legacy_boundary_postprocessing(prediction_hwc, H, W, out, workspace, workspace_bytes, info_host, stream_);
    const size_t n = static_cast<size_t>(H) * W;
    MBS_REQUIRE(H > 0 && W > 0 && n < (1ull << 31), "boundary_postprocessing: bad shape H=%d W=%d", H, W);
    MBS_REQUIRE(workspace_bytes >= mbs_postproc_workspace_bytes(H, W), "boundary_postprocessing: workspace too small");
    return run_tiled<true>(prediction_hwc, nullptr, H, W, W, 0.0f, 0.0f, out, workspace, workspace_bytes, info_host,
                           static_cast<cudaStream_t>(stream_));
}

extern "C" int mbs_pp_watershed(const float *image, const int32_t *markers, const uint8_t *mask, int H, int W,
                                int32_t *labels_out, void *workspace, size_t workspace_bytes, int64_t *info_host,
                                int force_sequential, void *stream_) {
    if (legacy_path())
        return legacy_pp_watershed(image, markers, mask, H, W, labels_out, workspace, workspace_bytes, info_host, force_sequential, stream_);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t n = static_cast<size_t>(H) * W;
    MBS_REQUIRE(H > 0 && W > 0 && n < (1ull << 31), "watershed: bad shape");
    // own carve (a subset of the fused pipeline's workspace): stats, state, labels, tile flags, heap of the exact flood
    const size_t tiles = static_cast<size_t>((H + CT - 1) / CT) * ((W + CT - 1) / CT);
    const size_t need = r256(sizeof(Stats)) + r256(n * 8) + r256(n * 4) + r256(2 * tiles + 256) + r256(n * sizeof(HeapItem));
    MBS_REQUIRE(workspace_bytes >= need, "watershed: workspace too small (%zu < %zu)", workspace_bytes, need);
    TiledWs t;
    {
        Carver cv{static_cast<char *>(workspace), workspace_bytes};
        t.st = cv.take<Stats>(1);
        t.state = cv.take<unsigned long long>(n);
        t.lab32 = cv.take<int>(n);
        t.tile_changed = cv.take<uint8_t>(2 * tiles + 256);
        t.heap = cv.take<HeapItem>(n);
        MBS_REQUIRE(cv.ok, "watershed: workspace carve failed");
    }
    CoopCfg *cfg = nullptr;
    int rc = coop_config(&cfg);
    if (rc) return rc;
    MBS_CHECK_CUDA(cudaMemsetAsync(t.st, 0, sizeof(Stats), stream));
    const int nn = static_cast<int>(n);
    if (!force_sequential) {
        flood_init_kernel<<<mbs::cdiv(nn, 256), 256, 0, stream>>>(image, markers, mask, nn, t.state, t.lab32);
        MBS_CHECK_LAUNCH();
        FloodParams fp;
        fp.img = image;
        fp.negate = 0;
        fp.L16 = nullptr;
        fp.marker_at_root = nullptr;
        fp.th_mask = 0.0f;
        fp.use_th_mask = 0;
        fp.light_init_max = light_limit(0);
        fp.light_items_max = light_limit(1);
        fp.chase_heavy = light_limit(2);
        fp.chase_light = light_limit(3);
        fp.light_min_tiles = light_limit(4);
        fp.badlist = nullptr;
        fp.bad_cap = 0;
        fp.state = t.state;
        fp.lab32 = t.lab32;
        fp.H = H;
        fp.W = W;
        fp.st = t.st;
        fp.tile_changed = t.tile_changed;
        fp.out16 = nullptr;
        fp.out32 = labels_out;
        const int ntiles = mbs::cdiv(W, CT) * mbs::cdiv(H, CT);
        const int blocks = ntiles < cfg->flood32_blocks ? ntiles : cfg->flood32_blocks;
        void *args[] = {(void *)&fp};
        MBS_CHECK_CUDA(cudaLaunchCooperativeKernel((void *)flood_kernel<true>, dim3(blocks), dim3(256), args, FLOOD_SMEM32, stream));
        mbs::count_launch();
    }
    ws_sequential_kernel<<<1, 32, 0, stream>>>(image, 0, markers, mask, H, W, labels_out, t.heap, t.st, force_sequential, nullptr);
    MBS_CHECK_LAUNCH();
    if (info_host) {
        rc = read_info(t, info_host, stream);
        if (rc) return rc;
        if (force_sequential) info_host[3] = 1;
    }
    return 0;
}
