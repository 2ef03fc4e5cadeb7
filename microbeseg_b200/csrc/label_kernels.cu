// Training-label generation for the distance method on the GPU (batched over crops).
//
// Replaces, per instance-mask crop:
//   src/training/train.py:74-79                       max_mal = ceil(max regionprops.major_axis_length)
//   src/training/train_data_representations.py:11-37  get_label(.., 'distance', max_mal)
//   train_data_representations.py:261-361             distance_label  (cell + neighbour distances)
//   train_data_representations.py:102-126             border_label
//   train_data_representations.py:40-72               bottom_hat_closing
//
// The reference loops over instances in Python and runs two scipy EDTs per instance on a
// (2R)x(2R) window plus a full-image binary_closing per instance.  Here one CTA owns one instance:
// the window lives in shared memory, the two exact Euclidean distance transforms are done as a
// column scan + row minimisation on integer squared distances (sqrt/division in float64 exactly as
// scipy/numpy), and every pixel is written by the CTA of its own instance only, so the reference's
// "+=" accumulation onto zeros is reproduced without atomics.  Everything else (border pixels,
// bottom-hat gaps, gap statistics, rescaling, grey closing) is per-pixel / per-gap work.
#include <cuda_runtime.h>

#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "../../include/mbseg.h"
#include "common.cuh"

namespace {

struct CellStats {
    unsigned long long cnt, sy, sx, syy, sxx, sxy;
    int y0, y1, x0, x1;        // bounding box (inclusive)
    int wy0, wy1, wx0, wx1;    // search window [wy0,wy1) x [wx0,wx1)
};
struct GapStats {
    unsigned long long cnt, sy, sx, syy, sxx, sxy;
    double bsum;               // sum of the raw neighbour map over the gap's outer 3x3 boundary
};
struct CropInfo {
    int max_mal;               // ceil(max major axis length)
    int radius;                // search radius actually used
    int n_gaps;
    int error;                 // bit 0: window too large for shared memory, bit 1: too many gaps
};

// "no site in this column": 2^15, so that k^2 + g^2 stays below 2^32 without a test (crop sides are < 32768, checked)
constexpr unsigned short kInf16 = 0x8000;

__global__ void lab_init_stats_kernel(CellStats *cs, int total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    CellStats z;
    memset(&z, 0, sizeof(z));
    z.y0 = INT_MAX; z.x0 = INT_MAX; z.y1 = -1; z.x1 = -1;
    cs[i] = z;
}

// Per-instance integer statistics.  All pixels of an instance hit the same ten addresses, so plain atomics
// serialise; a warp (32 consecutive pixels of one row) first groups its lanes by instance id and reduces each
// group with REDUX (__reduce_*_sync), then ONE lane per group issues the atomics.
// A block of 32 x 8 threads walks kRowsPerBlock rows (a warp is a piece of one row): with one row group per block these
// light kernels were bound by the block launch rate (80 k blocks per 200 crops), not by their work.
constexpr int kRowsPerBlock = 64;
__global__ void lab_accum_kernel(const uint16_t *__restrict__ masks, int H, int W, int ids, CellStats *cs) {
    const int crop = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int y_end = min(H, static_cast<int>(blockIdx.y + 1) * kRowsPerBlock);
  for (int y = blockIdx.y * kRowsPerBlock + threadIdx.y; y < y_end; y += blockDim.y) {      // warp-uniform trip count
    int id = 0;
    if (x < W) id = masks[(static_cast<size_t>(crop) * H + y) * W + x];
    const bool valid = id != 0 && id < ids;
    unsigned remaining = __ballot_sync(0xffffffffu, valid);
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const int cur = __shfl_sync(0xffffffffu, id, leader);
        const bool mine = valid && id == cur;
        const unsigned grp = __ballot_sync(0xffffffffu, mine);
        if (mine) {
            const unsigned ux = static_cast<unsigned>(x), uy = static_cast<unsigned>(y);
            const unsigned c = __popc(grp);
            const unsigned sy = __reduce_add_sync(grp, uy), sx = __reduce_add_sync(grp, ux);
            const unsigned syy = __reduce_add_sync(grp, uy * uy), sxx = __reduce_add_sync(grp, ux * ux);
            const unsigned sxy = __reduce_add_sync(grp, ux * uy);
            const int y0 = __reduce_min_sync(grp, y), y1 = __reduce_max_sync(grp, y);
            const int x0 = __reduce_min_sync(grp, x), x1 = __reduce_max_sync(grp, x);
            if (lane == leader) {
                CellStats *s = cs + static_cast<size_t>(crop) * ids + cur;
                atomicAdd(&s->cnt, static_cast<unsigned long long>(c));
                atomicAdd(&s->sy, static_cast<unsigned long long>(sy));
                atomicAdd(&s->sx, static_cast<unsigned long long>(sx));
                atomicAdd(&s->syy, static_cast<unsigned long long>(syy));
                atomicAdd(&s->sxx, static_cast<unsigned long long>(sxx));
                atomicAdd(&s->sxy, static_cast<unsigned long long>(sxy));
                atomicMin(&s->y0, y0); atomicMax(&s->y1, y1);
                atomicMin(&s->x0, x0); atomicMax(&s->x1, x1);
            }
        }
        remaining &= ~grp;
    }
  }
}

// 4*sqrt(eigenvalue) of the inertia tensor [[mu02, -mu11], [-mu11, mu20]] / n  (skimage regionprops)
__device__ void axis_lengths(unsigned long long cnt, unsigned long long sy, unsigned long long sx,
                             unsigned long long syy, unsigned long long sxx, unsigned long long sxy, double *major,
                             double *minor) {
    const double n = static_cast<double>(cnt);
    const double cy = static_cast<double>(sy) / n, cx = static_cast<double>(sx) / n;
    const double mu20 = static_cast<double>(syy) - n * cy * cy;
    const double mu02 = static_cast<double>(sxx) - n * cx * cx;
    const double mu11 = static_cast<double>(sxy) - n * cx * cy;
    const double a = mu02 / n, c = mu20 / n, b = -mu11 / n;
    const double h = 0.5 * (a + c), d = sqrt(0.25 * (a - c) * (a - c) + b * b);
    const double e1 = fmax(h + d, 0.0), e2 = fmax(h - d, 0.0);
    *major = 4.0 * sqrt(e1);
    *minor = 4.0 * sqrt(e2);
}

__global__ void lab_axes_kernel(const CellStats *cs, int ids, int total, CropInfo *info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const CellStats s = cs[i];
    if (s.cnt == 0) return;
    double major, minor;
    axis_lengths(s.cnt, s.sy, s.sx, s.syy, s.sxx, s.sxy, &major, &minor);
    atomicMax(&info[i / ids].max_mal, static_cast<int>(ceil(major)));
}

__global__ void lab_window_kernel(CellStats *cs, int ids, int total, int H, int W, CropInfo *info, int radius_override) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    CropInfo *ci = info + i / ids;
    // get_label: search_radius = int(np.ceil(0.75 * max_mal))
    const int R = radius_override >= 0 ? radius_override : static_cast<int>(ceil(0.75 * static_cast<double>(ci->max_mal)));
    if (i % ids == 0) ci->radius = R;
    CellStats *s = cs + i;
    if (s->cnt == 0) return;
    const double n = static_cast<double>(s->cnt);
    const double cy = rint(static_cast<double>(s->sy) / n), cx = rint(static_cast<double>(s->sx) / n);   // np.round
    s->wy0 = static_cast<int>(fmax(cy - R, 0.0));
    s->wy1 = static_cast<int>(fmin(cy + R, static_cast<double>(H)));
    s->wx0 = static_cast<int>(fmax(cx - R, 0.0));
    s->wx1 = static_cast<int>(fmin(cx + R, static_cast<double>(W)));
}

// ---------------------------------------------------------------------------------------------
// one CTA per (instance, crop): exact EDT of the instance and of "everything but the other
// instances" inside the search window (distance_label :280-330)
// ---------------------------------------------------------------------------------------------
// `sm`: 4 * wh * ww unsigned shorts of scratch -- shared memory (lab_cell_kernel) or, for windows that do not fit, a
// per-crop global buffer (lab_cell_big_kernel); every exit is block-uniform, so the function can be called in a loop.
__device__ void cell_edt(const uint16_t *__restrict__ masks, int H, int W, int crop, int id, const CellStats &s, int wh, int ww,
                         unsigned short *sm, float *__restrict__ cell_dist, double *__restrict__ nraw, float cell_clip) {
    unsigned short *lab = sm;                 // [wh][ww] labels
    unsigned short *g1 = sm + wh * ww;        // vertical distance to the nearest pixel with label != id (own pixels)
    unsigned short *g2 = g1 + wh * ww;        // vertical distance to the nearest pixel of ANOTHER instance
    unsigned short *own = g2 + wh * ww;       // compact list of the instance's pixels (window indices)
    // An instance covers ~10-15 % of its search window: the three per-pixel passes below walk the compact list instead of
    // the window (no divergence between own / foreign lanes, ~7x fewer iterations).  Windows of more than 65536 pixels
    // (lab_cell_big_kernel only) do not fit 16-bit indices and walk the window.
    const int n_win = wh * ww;
    const bool use_list = n_win <= 65536;
    __shared__ unsigned int s_max1, s_max2, s_any_other, s_n_own;
    __shared__ int s_bx0, s_bx1, s_by0, s_by1;
    __syncthreads();                          // a previous call's readers of the shared flags are done
    if (threadIdx.x == 0) {
        s_max1 = 0; s_max2 = 0; s_any_other = 0; s_n_own = 0;
        s_bx0 = ww; s_bx1 = -1; s_by0 = wh; s_by1 = -1;
    }
    __syncthreads();
    const uint16_t *m = masks + static_cast<size_t>(crop) * H * W;
    {
        // one window row per warp and step (no index division; warp-uniform trip counts for the full-mask ballots);
        // also the bounding box of the instance inside the window
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        int bx0 = ww, bx1 = -1, by0 = wh, by1 = -1;
        for (int y = wid; y < wh; y += nw) {
            const uint16_t *mrow = m + static_cast<size_t>(s.wy0 + y) * W + s.wx0;
            for (int xb = 0; xb < ww; xb += 32) {
                const int x = xb + lane;
                bool mine = false;
                if (x < ww) {
                    const unsigned short l = mrow[x];
                    lab[y * ww + x] = l;
                    mine = l == id;
                }
                const unsigned int bal = __ballot_sync(0xFFFFFFFFu, mine);
                if (bal) {
                    const int lo = xb + __ffs(bal) - 1, hi = xb + 31 - __clz(bal);
                    bx0 = lo < bx0 ? lo : bx0;
                    bx1 = hi > bx1 ? hi : bx1;
                    by0 = y < by0 ? y : by0;
                    by1 = y;
                    unsigned int pos = 0;
                    if (lane == 0) pos = atomicAdd(&s_n_own, __popc(bal));
                    if (use_list) {
                        pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
                        if (mine) own[pos + __popc(bal & ((1u << lane) - 1u))] = static_cast<unsigned short>(y * ww + x);
                    }
                }
            }
        }
        if (lane == 0 && bx1 >= 0) {          // warp-uniform values (derived from ballots)
            atomicMin(&s_bx0, bx0);
            atomicMax(&s_bx1, bx1);
            atomicMin(&s_by0, by0);
            atomicMax(&s_by1, by1);
        }
    }
    __syncthreads();
    const int n_own = static_cast<int>(s_n_own);
    if (n_own == 0) return;                   // instance has no pixel inside its window: max EDT == 0 -> `continue`
    const bool any_bg = n_own < n_win;
    const int bx0 = s_bx0, bx1 = s_bx1, by0 = s_by0, by1 = s_by1;
    // Column scans (one thread per column): vertical distance to the nearest site above / below, only where it can matter.
    // sites = pixels that are not the instance (g) or pixels of other instances (which = 1); rows r0..r1, columns c0..c1.
    auto scan_columns = [&](unsigned short *g, bool others, int c0, int c1, int r0, int r1) {
        bool any_site = false;
        for (int x = c0 + static_cast<int>(threadIdx.x); x <= c1; x += blockDim.x) {
            int d = -1;                       // distance to the last site seen going down (-1: none yet)
            for (int y = r0; y <= r1; ++y) {
                const int l = lab[y * ww + x];
                const bool site = others ? (l != 0 && l != id) : (l != id);
                any_site |= site;
                d = site ? 0 : (d < 0 ? -1 : d + 1);
                g[y * ww + x] = d < 0 ? kInf16 : static_cast<unsigned short>(d);
            }
            d = -1;
            for (int y = r1; y >= r0; --y) {
                const int l = lab[y * ww + x];
                const bool site = others ? (l != 0 && l != id) : (l != id);
                d = site ? 0 : (d < 0 ? -1 : d + 1);
                if (d >= 0 && d < g[y * ww + x]) g[y * ww + x] = static_cast<unsigned short>(d);
            }
        }
        return any_site;
    };
    // exact minimum over the columns c0..c1 of row y, visited outwards from x: once dx^2 alone reaches the best squared
    // distance no farther column can improve it
    auto row_min = [&](const unsigned short *r, int x, int c0, int c1, unsigned int b) {
        const int kmax = x - c0 > c1 - x ? x - c0 : c1 - x;
        for (int k = 0; k <= kmax; ++k) {
            const unsigned int dx2 = static_cast<unsigned int>(k * k);
            if (dx2 >= b) break;
            // branch-free step: clamped (always readable) loads, out-of-range sides masked afterwards; a column without a
            // site holds kInf16, whose square exceeds every real squared distance (and the capped start value of pass 2)
            const int xl = x - k, xr = x + k;
            const unsigned int gl = r[xl >= c0 ? xl : c0], gr = r[xr <= c1 ? xr : c1];
            const unsigned int dl = xl >= c0 ? dx2 + gl * gl : 0xFFFFFFFFu;
            const unsigned int dr = xr <= c1 ? dx2 + gr * gr : 0xFFFFFFFFu;
            const unsigned int dm = dl < dr ? dl : dr;
            b = dm < b ? dm : b;
        }
        return b;
    };
    // Pass 1: distance to "not my id".  The nearest such pixel of an instance pixel lies inside the instance's bounding box
    // grown by one pixel (a site farther out has a closer site between it and the pixel), so only that part of the window is
    // scanned and searched.
    const int c0a = bx0 > 0 ? bx0 - 1 : 0, c1a = bx1 + 1 < ww ? bx1 + 1 : ww - 1;
    const int r0a = by0 > 0 ? by0 - 1 : 0, r1a = by1 + 1 < wh ? by1 + 1 : wh - 1;
    if (any_bg) scan_columns(g1, false, c0a, c1a, r0a, r1a);
    __syncthreads();
    unsigned int lmax1 = 0;
    for (int j = threadIdx.x; j < (use_list ? n_own : n_win); j += blockDim.x) {
        const int i = use_list ? own[j] : j;
        if (!use_list && lab[i] != id) continue;
        const int y = i / ww, x = i - y * ww;
        // all-foreground window: scipy's feature transform points at (-1, 0)
        const unsigned int b1 = any_bg ? row_min(g1 + y * ww, x, c0a, c1a, 0xFFFFFFFFu) : static_cast<unsigned int>((y + 1) * (y + 1) + x * x);
        lmax1 = b1 > lmax1 ? b1 : lmax1;
        // park the exact squared distance in the output array (bit pattern); the same thread converts it below
        const size_t o = (static_cast<size_t>(crop) * H + s.wy0 + y) * W + s.wx0 + x;
        cell_dist[o] = __uint_as_float(b1);
    }
    atomicMax(&s_max1, lmax1);
    __syncthreads();
    // Pass 2: distance to the other instances.  It only matters below den = min(max1 + 3, max2) (:321, larger values are
    // clipped to 1), so the outward search starts from the bound (floor(max1 + 3) + 1)^2: a pixel without a site inside it
    // has d2 >= max1 + 3, hence max2 >= max1 + 3, den = max1 + 3 and its value clips -- exact, and the search is short.
    // For the same reason only the bounding box grown by `cap` is scanned; other instances farther away than that change
    // nothing (every value clips, as if there were none).
    const unsigned int cap = static_cast<unsigned int>(floor(sqrt(static_cast<double>(s_max1)) + 3.0)) + 1u;
    const unsigned int cap2 = cap < 65535u ? cap * cap : 0xFFFFFFFFu;
    const int capi = cap < 65535u ? static_cast<int>(cap) : 65535;
    const int c0b = bx0 > capi ? bx0 - capi : 0, c1b = bx1 + capi < ww ? bx1 + capi : ww - 1;
    const int r0b = by0 > capi ? by0 - capi : 0, r1b = by1 + capi < wh ? by1 + capi : wh - 1;
    if (scan_columns(g2, true, c0b, c1b, r0b, r1b)) s_any_other = 1;
    __syncthreads();
    const bool any_other = s_any_other != 0;
    unsigned int lmax2 = 0;
    if (any_other) {
        for (int j = threadIdx.x; j < (use_list ? n_own : n_win); j += blockDim.x) {
            const int i = use_list ? own[j] : j;
            if (!use_list && lab[i] != id) continue;
            const int y = i / ww, x = i - y * ww;
            const unsigned int b2 = row_min(g2 + y * ww, x, c0b, c1b, cap2);
            lmax2 = b2 > lmax2 ? b2 : lmax2;        // a capped value (>= cap2) keeps max2 >= (max1 + 3)^2, which is all den needs
            const size_t o = (static_cast<size_t>(crop) * H + s.wy0 + y) * W + s.wx0 + x;
            nraw[o] = __longlong_as_double(static_cast<long long>(b2));
        }
    }
    atomicMax(&s_max2, lmax2);
    __syncthreads();
    const double max1 = sqrt(static_cast<double>(s_max1));               // np.max(EDT) (> 0: own pixels exist)
    const double max2 = sqrt(static_cast<double>(s_max2));
    const double den = fmin(max1 + 3.0, max2);                           // :321
    for (int j = threadIdx.x; j < (use_list ? n_own : n_win); j += blockDim.x) {
        const int i = use_list ? own[j] : j;
        if (!use_list && lab[i] != id) continue;
        const int y = i / ww, x = i - y * ww;
        const size_t o = (static_cast<size_t>(crop) * H + s.wy0 + y) * W + s.wx0 + x;
        const double d1 = sqrt(static_cast<double>(__float_as_uint(cell_dist[o])));
        // :292, cast :361 -- or cell_distance_label(apply_clipping=True) (:246-256): the raw EDT clipped to clip_val, / clip_val
        cell_dist[o] = cell_clip > 0.0f ? static_cast<float>(fmin(d1, static_cast<double>(cell_clip)) / static_cast<double>(cell_clip))
                                        : static_cast<float>(d1 / max1);
        double v = 0.0;
        if (any_other) {
            const double d2 = sqrt(static_cast<double>(static_cast<unsigned int>(__double_as_longlong(nraw[o]))));
            double q = d2 / den;
            q = q < 0.0 ? 0.0 : (q > 1.0 ? 1.0 : q);
            v = 1.0 - q;                                                  // :327 (own mask == 1 here)
        }
        nraw[o] = v;
    }
}

__global__ void __launch_bounds__(256)
lab_cell_kernel(const uint16_t *__restrict__ masks, int H, int W, int ids, const CellStats *cs, CropInfo *info,
                float *__restrict__ cell_dist, double *__restrict__ nraw, int smem_cap_elems, float cell_clip) {
    const int crop = blockIdx.y;
    const int id = blockIdx.x + 1;
    if (id >= ids) return;
    const CellStats s = cs[static_cast<size_t>(crop) * ids + id];
    if (s.cnt == 0) return;
    const int wh = s.wy1 - s.wy0, ww = s.wx1 - s.wx0;
    if (wh <= 0 || ww <= 0) return;           // empty crop: np.max(...) of an empty EDT -> skipped
    if (wh * ww > smem_cap_elems) return;     // window larger than shared memory: lab_cell_big_kernel takes it
    extern __shared__ unsigned short sm[];
    cell_edt(masks, H, W, crop, id, s, wh, ww, sm, cell_dist, nraw, cell_clip);
}

// Instances whose search window exceeds shared memory (max_mal >~ 113 px): one CTA per crop walks them one after the
// other with the window state in a per-crop global buffer (3 * H * W unsigned shorts).  Rare and slow, but exact -- the
// reference has no size limit (train_data_representations.py:280-330).
__global__ void __launch_bounds__(256)
lab_cell_big_kernel(const uint16_t *__restrict__ masks, int H, int W, int ids, const CellStats *cs, unsigned short *scratch,
                    size_t scratch_stride, float *__restrict__ cell_dist, double *__restrict__ nraw, int smem_cap_elems, float cell_clip) {
    const int crop = blockIdx.x;
    unsigned short *sm = scratch + static_cast<size_t>(crop) * scratch_stride;
    // the usual case is "none": all threads look for oversized windows in parallel and the CTA leaves at once
    __shared__ int s_list[256];
    __shared__ int s_n;
    for (int id0 = 1; id0 < ids; id0 += 256 * 64) {           // chunks of ids (a chunk never yields more than 256 * 64 candidates)
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        for (int id = id0 + threadIdx.x; id < ids && id < id0 + 256 * 64; id += 256) {
            const CellStats s = cs[static_cast<size_t>(crop) * ids + id];
            const int wh = s.wy1 - s.wy0, ww = s.wx1 - s.wx0;
            if (s.cnt != 0 && wh > 0 && ww > 0 && wh * ww > smem_cap_elems) {
                const int pos = atomicAdd(&s_n, 1);
                if (pos < 256) s_list[pos] = id;
            }
        }
        __syncthreads();
        const int n = s_n;
        if (n == 0) continue;
        if (n <= 256) {
            for (int k = 0; k < n; ++k) {
                const int id = s_list[k];
                const CellStats s = cs[static_cast<size_t>(crop) * ids + id];
                cell_edt(masks, H, W, crop, id, s, s.wy1 - s.wy0, s.wx1 - s.wx0, sm, cell_dist, nraw, cell_clip);
            }
        } else {                                              // more than 256 oversized instances in the chunk: plain walk
            for (int id = id0; id < ids && id < id0 + 256 * 64; ++id) {
                const CellStats s = cs[static_cast<size_t>(crop) * ids + id];
                if (s.cnt == 0) continue;
                const int wh = s.wy1 - s.wy0, ww = s.wx1 - s.wx0;
                if (wh <= 0 || ww <= 0 || wh * ww <= smem_cap_elems) continue;
                cell_edt(masks, H, W, crop, id, s, wh, ww, sm, cell_dist, nraw, cell_clip);
            }
        }
        __syncthreads();
    }
}

constexpr int CB_MAX_ROWS = 96;      // window rows (bbox + 6) handled by the warp-per-instance bit-parallel closing

// per-instance binary_closing(nucleus, disk(3)) OR-ed into the label_bin bit image (bottom_hat_closing :50-55) for
// instances too large for lab_close_cell_bits_kernel;
// scipy's erosion uses border_value=0, i.e. pixels whose disk leaves the image are removed.
__device__ __forceinline__ bool in_disk3(int dy, int dx) { return dy * dy + dx * dx <= 9; }

__global__ void __launch_bounds__(256)
lab_close_cell_kernel(const uint16_t *__restrict__ masks, int H, int W, int WW, int ids, const CellStats *cs, CropInfo *info,
                      unsigned long long *__restrict__ lbits, int smem_cap_bytes, const int *__restrict__ big_list) {
  // big_list[0] = number of oversized instances, then their (crop * ids + id) keys (appended by lab_close_cell_bits_kernel)
  for (int k = blockIdx.x; k < big_list[0]; k += gridDim.x) {
    __syncthreads();                                    // smem reuse between instances
    const int key = big_list[1 + k];
    const int crop = key / ids, id = key - crop * ids;
    const CellStats s = cs[key];
    const int bh = s.y1 - s.y0 + 1, bw = s.x1 - s.x0 + 1;
    const int dh = bh + 6, dw = bw + 6;
    if (bh * bw + dh * dw > smem_cap_bytes) {
        if (threadIdx.x == 0) atomicOr(&info[crop].error, 1);
        continue;
    }
    extern __shared__ unsigned short sm[];
    uint8_t *X = reinterpret_cast<uint8_t *>(sm);      // [bh][bw]
    uint8_t *D = X + bh * bw;                          // [dh][dw], origin (y0-3, x0-3)
    const uint16_t *m = masks + static_cast<size_t>(crop) * H * W;
    for (int i = threadIdx.x; i < bh * bw; i += blockDim.x) {
        const int y = i / bw, x = i - y * bw;
        X[i] = m[static_cast<size_t>(s.y0 + y) * W + s.x0 + x] == id;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < dh * dw; i += blockDim.x) {
        const int y = i / dw - 3, x = i % dw - 3;      // bbox coordinates
        const int gy = s.y0 + y, gx = s.x0 + x;
        bool v = false;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            for (int dy = -3; dy <= 3 && !v; ++dy)
                for (int dx = -3; dx <= 3; ++dx) {
                    if (!in_disk3(dy, dx)) continue;
                    const int yy = y + dy, xx = x + dx;
                    if (yy >= 0 && yy < bh && xx >= 0 && xx < bw && X[yy * bw + xx]) { v = true; break; }
                }
        }
        D[i] = v;
    }
    __syncthreads();
    unsigned long long *out = lbits + static_cast<size_t>(crop) * H * WW;
    for (int i = threadIdx.x; i < bh * bw; i += blockDim.x) {
        const int y = i / bw, x = i - y * bw;
        bool v = true;
        for (int dy = -3; dy <= 3 && v; ++dy)
            for (int dx = -3; dx <= 3; ++dx) {
                if (!in_disk3(dy, dx)) continue;
                const int gy = s.y0 + y + dy, gx = s.x0 + x + dx;
                if (gy < 0 || gy >= H || gx < 0 || gx >= W || !D[(y + dy + 3) * dw + (x + dx + 3)]) { v = false; break; }
            }
        if (v) atomicOr(out + static_cast<size_t>(s.y0 + y) * WW + ((s.x0 + x) >> 6), 1ull << ((s.x0 + x) & 63));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Bit-parallel morphology with disk(3) = {|dx|<=3, dy=0} U {|dx|<=2, |dy|<=2} U {dx=0, |dy|=3} (x^2+y^2 <= 9):
// image rows are 64-bit words (bit b of word w = pixel x = 64w+b), a dilation / erosion of 64 pixels is ~40 word
// operations instead of 64 x 29 byte tests.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long or_spread(unsigned long long c, int k) {       // single-word window
    unsigned long long r = c;
    for (int s = 1; s <= k; ++s) r |= (c << s) | (c >> s);
    return r;
}
__device__ __forceinline__ unsigned long long and_spread(unsigned long long c, int k) {      // zeros shift in at the ends
    unsigned long long r = c;
    for (int s = 1; s <= k; ++s) r &= (c << s) & (c >> s);
    return r;
}
// OR / AND over horizontal shifts -k..k of a row given as (left word, centre word, right word)
__device__ __forceinline__ unsigned long long or_spread3(unsigned long long l, unsigned long long c, unsigned long long r, int k) {
    unsigned long long o = c;
    for (int s = 1; s <= k; ++s) o |= ((c << s) | (l >> (64 - s))) | ((c >> s) | (r << (64 - s)));
    return o;
}
__device__ __forceinline__ unsigned long long and_spread3(unsigned long long l, unsigned long long c, unsigned long long r, int k) {
    unsigned long long o = c;
    for (int s = 1; s <= k; ++s) o &= ((c << s) | (l >> (64 - s))) & ((c >> s) | (r << (64 - s)));
    return o;
}


// per-instance binary_closing(nucleus, disk(3)) OR-ed into the label_bin BIT image (bottom_hat_closing :50-55); one
// warp per instance, the window (bbox + 3 px) is at most 64 columns x 96 rows (larger instances: lab_close_cell_kernel).
// Window bit b = image column x0 - 3 + b.  scipy's erosion uses border_value = 0.
__global__ void __launch_bounds__(256)
lab_close_cell_bits_kernel(const uint16_t *__restrict__ masks, int H, int W, int WW, int ids, const CellStats *cs,
                           unsigned long long *__restrict__ lbits, int *__restrict__ big_list) {
    __shared__ unsigned long long sX[8][CB_MAX_ROWS + 6], sD[8][CB_MAX_ROWS + 6];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crop = blockIdx.y;
    const int id = blockIdx.x * 8 + wib + 1;
    if (id >= ids) return;
    const CellStats s = cs[static_cast<size_t>(crop) * ids + id];
    if (s.cnt == 0) return;
    const int bh = s.y1 - s.y0 + 1, bw = s.x1 - s.x0 + 1;
    const int dh = bh + 6, dw = bw + 6;
    if (dw > 64 || dh > CB_MAX_ROWS) {                  // queue for the byte-wise kernel
        if (lane == 0) big_list[1 + atomicAdd(&big_list[0], 1)] = crop * ids + id;
        return;
    }
    unsigned long long *X = sX[wib] + 3, *D = sD[wib] + 3;       // rows -3 .. dh+2 addressable
    const uint16_t *m = masks + static_cast<size_t>(crop) * H * W;
    const int gx0 = s.x0 - 3, gy0 = s.y0 - 3;
    // columns of the window that lie inside the image
    unsigned long long colmask = 0;
    for (int b = 0; b < dw; ++b)
        if (gx0 + b >= 0 && gx0 + b < W) colmask |= 1ull << b;
    if (lane < 3) { X[-1 - lane] = 0; X[dh + lane] = 0; D[-1 - lane] = 0; D[dh + lane] = 0; }
    // rows as ballots: lane = column (two halves)
    for (int r = 0; r < dh; ++r) {
        const int gy = gy0 + r;
        unsigned lo = 0, hi = 0;
        if (gy >= 0 && gy < H) {             // warp-uniform
            const int xa = gx0 + lane, xb = gx0 + 32 + lane;
            const bool a = lane < dw && xa >= 0 && xa < W && m[static_cast<size_t>(gy) * W + xa] == id;
            const bool b = 32 + lane < dw && xb >= 0 && xb < W && m[static_cast<size_t>(gy) * W + xb] == id;
            lo = __ballot_sync(0xffffffffu, a);
            hi = __ballot_sync(0xffffffffu, b);
        }
        if (lane == 0) X[r] = (static_cast<unsigned long long>(hi) << 32) | lo;
    }
    __syncwarp();
    for (int r = lane; r < dh; r += 32) {    // dilation, defined inside the image only
        const int gy = gy0 + r;
        unsigned long long d = 0;
        if (gy >= 0 && gy < H)
            d = (or_spread(X[r], 3) | or_spread(X[r - 1] | X[r + 1] | X[r - 2] | X[r + 2], 2) | X[r - 3] | X[r + 3]) & colmask;
        D[r] = d;
    }
    __syncwarp();
    unsigned long long *out = lbits + static_cast<size_t>(crop) * H * WW;
    for (int r = 3 + lane; r < 3 + bh; r += 32) {      // erosion at the bounding-box rows
        unsigned long long e = and_spread(D[r], 3) & and_spread(D[r - 1] & D[r + 1] & D[r - 2] & D[r + 2], 2) & D[r - 3] & D[r + 3];
        if (!e) continue;
        const int gy = gy0 + r;
        int gx = gx0;
        if (gx < 0) { e >>= -gx; gx = 0; }
        const int w = gx >> 6, sh = gx & 63;
        unsigned long long *row = out + static_cast<size_t>(gy) * WW;
        atomicOr(row + w, e << sh);
        if (sh && w + 1 < WW && (e >> (64 - sh))) atomicOr(row + w + 1, e >> (64 - sh));
    }
}

__device__ __forceinline__ void load3(const unsigned long long *img, int H, int WW, int y, int w, unsigned long long &l,
                                      unsigned long long &c, unsigned long long &r) {
    l = c = r = 0;
    if (y < 0 || y >= H) return;
    const unsigned long long *row = img + static_cast<size_t>(y) * WW;
    c = row[w];
    if (w > 0) l = row[w - 1];
    if (w + 1 < WW) r = row[w + 1];
}
// dil = binary_dilation(label_bin, disk(3)) on bit images (bottom_hat_closing :58)
__global__ void lab_dilate_bits_kernel(const unsigned long long *__restrict__ in, int H, int W, int WW, int n_crops,
                                       unsigned long long *__restrict__ out) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<long long>(n_crops) * H * WW) return;
    const int w = static_cast<int>(t % WW), y = static_cast<int>((t / WW) % H);
    const unsigned long long *img = in + (t / (static_cast<long long>(WW) * H)) * H * WW;
    unsigned long long d = 0, l, c, r;
    load3(img, H, WW, y, w, l, c, r);
    d |= or_spread3(l, c, r, 3);
#pragma unroll
    for (int dy = 1; dy <= 2; ++dy) {
        load3(img, H, WW, y - dy, w, l, c, r);
        d |= or_spread3(l, c, r, 2);
        load3(img, H, WW, y + dy, w, l, c, r);
        d |= or_spread3(l, c, r, 2);
    }
    load3(img, H, WW, y - 3, w, l, c, r);
    d |= c;
    load3(img, H, WW, y + 3, w, l, c, r);
    d |= c;
    const int valid = W - 64 * w;                      // columns of this word inside the image
    if (valid < 64) d &= (1ull << valid) - 1;
    out[t] = d;
}
// gap = ~label_bin & (erode(dil) ^ label_bin) as BYTES for the component labelling (bottom_hat_closing :58-59)
// ... and L[i] = i at the gap pixels: the union-find arrays are only ever touched where gap != 0 (gaps are ~1 % of the
// pixels, so the labelling passes below read one byte per pixel instead of 4-byte ids).
// lbits: in = closed-label bits of word t, out = the gap bits of word t (each thread reads and writes only its own word;
// gap_stats_kernel uses the bit image to reject pixels without a gap in their 3x3 neighbourhood)
__global__ void lab_erode_gap_bits_kernel(const unsigned long long *__restrict__ dil, unsigned long long *lbits,
                                          int H, int W, int WW, int n_crops, uint8_t *__restrict__ gap, int *__restrict__ L) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= static_cast<long long>(n_crops) * H * WW) return;
    const int w = static_cast<int>(t % WW), y = static_cast<int>((t / WW) % H);
    const long long crop = t / (static_cast<long long>(WW) * H);
    const unsigned long long *img = dil + crop * H * WW;
    unsigned long long l, c, r;
    load3(img, H, WW, y, w, l, c, r);
    unsigned long long e = and_spread3(l, c, r, 3);
#pragma unroll
    for (int dy = 1; dy <= 2; ++dy) {
        load3(img, H, WW, y - dy, w, l, c, r);
        e &= and_spread3(l, c, r, 2);
        load3(img, H, WW, y + dy, w, l, c, r);
        e &= and_spread3(l, c, r, 2);
    }
    load3(img, H, WW, y - 3, w, l, c, r);
    e &= c;
    load3(img, H, WW, y + 3, w, l, c, r);
    e &= c;
    const unsigned long long lb = lbits[t];
    const unsigned long long g = ~lb & (e ^ lb);
    const int valid = min(64, W - 64 * w);
    uint8_t *dst = gap + (crop * H + y) * W + 64 * w;
    {
        const int pix0 = static_cast<int>((crop * H + y) * W + 64 * w);
        unsigned long long rest = valid < 64 ? (g & ((1ull << valid) - 1)) : g;
        lbits[t] = rest;
        while (rest) {
            const int b = __ffsll(static_cast<long long>(rest)) - 1;
            L[pix0 + b] = pix0 + b;
            rest &= rest - 1;
        }
    }
    if (valid == 64 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        // 8 bits -> 8 bytes: isolate bit i in byte i, then normalise every non-zero byte to 1
        auto expand8 = [](unsigned long long x) {
            unsigned long long v = (x * 0x0101010101010101ull) & 0x8040201008040201ull;
            return ((v + 0x7F7F7F7F7F7F7F7Full) >> 7) & 0x0101010101010101ull;
        };
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned long long a = expand8((g >> (16 * q)) & 0xFFull), b = expand8((g >> (16 * q + 8)) & 0xFFull);
            *reinterpret_cast<uint4 *>(dst + 16 * q) = make_uint4(static_cast<unsigned>(a), static_cast<unsigned>(a >> 32),
                                                                 static_cast<unsigned>(b), static_cast<unsigned>(b >> 32));
        }
    } else {
        for (int b = 0; b < valid; ++b) dst[b] = static_cast<uint8_t>((g >> b) & 1ull);
    }
}

// border_label(label) == 2  <=>  foreground pixel with a different positive id in its 3x3 neighbourhood
// eight pixels of a row per thread: three 16-byte row loads (+ the two edge pixels) instead of 72 two-byte loads, one 8-byte store
__global__ void __launch_bounds__(256) lab_border_kernel(const uint16_t *__restrict__ masks, int H, int W, uint8_t *__restrict__ border) {
    const int crop = blockIdx.z;
    // groups of eight pixels numbered row-major over the crop: consecutive lanes take consecutive groups across row ends,
    // so every lane has work whatever W is (W = 320: 40 groups per row)
    const int ngx = (W + 7) >> 3;
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int y = t / ngx, x0 = (t - y * ngx) * 8;
    if (y >= H) return;
    const uint16_t *m = masks + static_cast<size_t>(crop) * H * W;
    uint8_t *out = border + (static_cast<size_t>(crop) * H + y) * W + x0;
    const bool vec = (W & 7) == 0;            // rows are 16-byte aligned and whole groups of eight
    unsigned short r[3][10];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int yy = y + dy - 1;
        if (yy < 0 || yy >= H) {
#pragma unroll
            for (int i = 0; i < 10; ++i) r[dy][i] = 0;
            continue;
        }
        const uint16_t *row = m + static_cast<size_t>(yy) * W;
        r[dy][0] = x0 > 0 ? row[x0 - 1] : 0;
        if (vec) {
            const uint4 v = *reinterpret_cast<const uint4 *>(row + x0);
            const unsigned w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                r[dy][1 + 2 * i] = static_cast<unsigned short>(w4[i] & 0xFFFFu);
                r[dy][2 + 2 * i] = static_cast<unsigned short>(w4[i] >> 16);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) r[dy][1 + i] = x0 + i < W ? row[x0 + i] : 0;
        }
        r[dy][9] = x0 + 8 < W ? row[x0 + 8] : 0;
    }
    unsigned long long packed = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned l = r[1][i + 1];
        bool b = false;
        if (l) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const unsigned o = r[dy][i + dx];       // out-of-image neighbours read as 0 (skipped by the reference)
                    b |= (o != 0 && o != l);
                }
        }
        packed |= static_cast<unsigned long long>(b ? 1u : 0u) << (8 * i);
    }
    if (vec) {
        *reinterpret_cast<unsigned long long *>(out) = packed;
    } else {
        for (int i = 0; i < 8 && x0 + i < W; ++i) out[i] = static_cast<uint8_t>((packed >> (8 * i)) & 1u);
    }
}

// ---------------------------------------------------------------------------------------------
// gaps: 8-connected components per crop (union-find on global pixel indices), dense ids, statistics
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(volatile int *L, int a) {
    int p = L[a];
    while (p != a) { a = p; p = L[a]; }
    return a;
}
__device__ __forceinline__ void uf_union(int *L, int a, int b) {
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) { int old = atomicMin(&L[b], a); done = (old == b); b = old; }
        else if (b < a) { int old = atomicMin(&L[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}
// The gap passes walk the gap BYTES sixteen at a time (one 16-byte load per thread; gaps are ~1 % of the pixels, so almost
// every thread leaves after that load) and call f(i) for every gap pixel i.
template <typename F>
__device__ __forceinline__ void for_gap_pixels(const uint8_t *__restrict__ gap, long long n, F f) {
    const long long i0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 16;
    if (i0 >= n) return;
    if (i0 + 16 <= n && (reinterpret_cast<uintptr_t>(gap + i0) & 15) == 0) {
        const uint4 v = *reinterpret_cast<const uint4 *>(gap + i0);
        if (!(v.x | v.y | v.z | v.w)) return;
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if ((w[k] >> (8 * b)) & 0xFFu) f(i0 + 4 * k + b);
    } else {
        for (long long i = i0; i < n && i < i0 + 16; ++i)
            if (gap[i]) f(i);
    }
}
// 8-connected union of the gap pixels.  Fast path (W a multiple of 16: a thread's 16 bytes never straddle a row): the
// thread works on RUNS -- pixels of a run link to the run's first pixel with one atomicMin each (no find chains), and a
// run is united once with every run of the row above that touches columns [a-1, b+1] instead of up to three unions per
// pixel.  The resulting partition is the same as that of the per-pixel path below.
__global__ void gap_merge_kernel(const uint8_t *__restrict__ gap, long long n, int H, int W, int *L) {
    if ((W & 15) == 0 && (reinterpret_cast<uintptr_t>(gap) & 15) == 0) {
        const long long i0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 16;
        if (i0 >= n) return;
        const uint4 v = *reinterpret_cast<const uint4 *>(gap + i0);
        if (!(v.x | v.y | v.z | v.w)) return;
        auto bits4 = [](unsigned w) { return (w & 1u) | ((w >> 7) & 2u) | ((w >> 14) & 4u) | ((w >> 21) & 8u); };   // bytes are 0 / 1
        auto bits16 = [&](const uint4 &q) { return bits4(q.x) | (bits4(q.y) << 4) | (bits4(q.z) << 8) | (bits4(q.w) << 12); };
        const unsigned cur = bits16(v);
        const int HW = H * W;
        const int pix = static_cast<int>(i0 % HW);
        const int y = pix / W, x0 = pix - y * W;
        const int g0 = static_cast<int>(i0);
        unsigned above = 0;                       // bit p + 1 <-> column x0 + p of the row above, p = -1 .. 16
        if (y > 0) {
            above = bits16(*reinterpret_cast<const uint4 *>(gap + i0 - W)) << 1;
            if (x0 > 0 && gap[i0 - W - 1]) above |= 1u;
            if (x0 + 16 < W && gap[i0 - W + 16]) above |= 1u << 17;
        }
        const bool left = x0 > 0 && gap[i0 - 1];
        unsigned rest = cur;
        while (rest) {
            const int a = __ffs(rest) - 1;
            const int len = __ffs(~(cur >> a)) - 1;                 // run [a, a + len)
            const int ia = g0 + a;
            for (int k = a + 1; k < a + len; ++k) {
                const int old = atomicMin(&L[g0 + k], ia);
                if (old != g0 + k && old != ia) uf_union(L, old, ia);   // a pixel below had linked it already
            }
            if (a == 0 && left) uf_union(L, ia, ia - 1);
            unsigned am = above & (((1u << (len + 2)) - 1u) << a);   // columns a - 1 .. a + len of the row above
            while (am) {
                const int q = __ffs(am) - 1;
                uf_union(L, ia, g0 - W + q - 1);
                const int l2 = __ffs(~(above >> q)) - 1;            // skip the rest of that run
                am &= ~(((1u << l2) - 1u) << q);
            }
            rest &= ~(((1u << len) - 1u) << a);
        }
        return;
    }
    for_gap_pixels(gap, n, [&](long long gi) {
        const int HW = H * W;
        const int base = static_cast<int>(gi / HW) * HW;
        const int i = static_cast<int>(gi - base);
        const int y = i / W, x = i - y * W;
        const uint8_t *g = gap + base;
        if (x > 0 && g[i - 1]) uf_union(L, base + i, base + i - 1);
        if (y > 0) {
            if (g[i - W]) uf_union(L, base + i, base + i - W);
            else {
                if (x > 0 && g[i - W - 1]) uf_union(L, base + i, base + i - W - 1);
                if (x + 1 < W && g[i - W + 1]) uf_union(L, base + i, base + i - W + 1);
            }
        }
    });
}
__global__ void gap_compress_ids_kernel(const uint8_t *__restrict__ gap, int *L, long long n, int HW, int max_gaps, CropInfo *info,
                                        int *gid) {
    for_gap_pixels(gap, n, [&](long long i) {               // gid is only defined at gap pixels
        if (L[i] == i) {          // root claims a dense id inside its crop
            const int crop = static_cast<int>(i / HW);
            const int k = atomicAdd(&info[crop].n_gaps, 1);
            if (k >= max_gaps) { atomicOr(&info[crop].error, 2); gid[i] = -1; }
            else gid[i] = k;
        }
    });
}
__global__ void gap_resolve_kernel(const uint8_t *__restrict__ gap, int *L, long long n) {
    for_gap_pixels(gap, n, [&](long long i) { L[i] = uf_find(L, static_cast<int>(i)); });
}
__global__ void gap_assign_kernel(const uint8_t *__restrict__ gap, const int *__restrict__ L, long long n, int *gid) {
    for_gap_pixels(gap, n, [&](long long i) {
        if (L[i] != i) gid[i] = gid[L[i]];
    });
}
__global__ void gap_stats_kernel(const uint8_t *__restrict__ gap, const int *__restrict__ gid, const double *__restrict__ nraw,
                                 const unsigned long long *__restrict__ gbits, int WW, int H, int W, int max_gaps, GapStats *gs) {
    const int crop = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y_end = min(H, static_cast<int>(blockIdx.y + 1) * kRowsPerBlock);
  for (int y = blockIdx.y * kRowsPerBlock + threadIdx.y; y < y_end; y += blockDim.y) {      // warp-uniform trip count
    const bool inside = x < W;                // no early exit of single lanes: the warp aggregates below (a warp is a row piece)
    // ~97 % of the pixels have no gap pixel in their 3x3 neighbourhood: they are rejected on the gap BIT image (three
    // words that the whole warp shares) before any per-pixel load
    const unsigned long long *rows = gbits + static_cast<size_t>(crop) * H * WW;
    const int w = x >> 6, b = x & 63;
    auto near3 = [&](int yy) -> unsigned {       // bits {x-1, x, x+1} of row yy (bits beyond W are zero)
        const unsigned long long *r = rows + static_cast<size_t>(yy) * WW;
        const unsigned long long c = r[w];
        unsigned m = static_cast<unsigned>((c >> b) & 1ull) << 1;
        m |= b > 0 ? static_cast<unsigned>((c >> (b - 1)) & 1ull) : (w > 0 ? static_cast<unsigned>(r[w - 1] >> 63) : 0u);
        m |= (b < 63 ? static_cast<unsigned>((c >> (b + 1)) & 1ull) : (w + 1 < WW ? static_cast<unsigned>(r[w + 1] & 1ull) : 0u)) << 2;
        return m;
    };
    unsigned mid = 0, any = 0;        // any: bit 3 * (dy + 1) + (dx + 1) <-> neighbour (dy, dx) is a gap pixel
    if (inside) {
        mid = near3(y);
        any = mid << 3;
        if (y > 0) any |= near3(y - 1);
        if (y + 1 < H) any |= near3(y + 1) << 6;
    }
    if (!__any_sync(0xffffffffu, any != 0)) continue;
    const size_t base = static_cast<size_t>(crop) * H * W;
    const int *g = gid + base;
    GapStats *S = gs + static_cast<size_t>(crop) * max_gaps;
    const int me = (mid & 2u) ? g[y * W + x] : -1;
    {
        // all pixels of a gap hit the same six addresses: group the warp's lanes by gap id, reduce each group with REDUX,
        // one lane per group issues the atomics (as lab_accum_kernel)
        const int lane = threadIdx.x & 31;
        unsigned remaining = __ballot_sync(0xffffffffu, me >= 0);
        while (remaining) {
            const int leader = __ffs(remaining) - 1;
            const int cur = __shfl_sync(0xffffffffu, me, leader);
            const bool mine = me == cur;
            const unsigned grp = __ballot_sync(0xffffffffu, mine);
            if (mine) {
                const unsigned ux = static_cast<unsigned>(x), uy = static_cast<unsigned>(y);
                const unsigned sx = __reduce_add_sync(grp, ux), sxx = __reduce_add_sync(grp, ux * ux);
                if (lane == leader) {
                    const unsigned long long c = __popc(grp);       // same row: y is common to the group
                    atomicAdd(&S[cur].cnt, c);
                    atomicAdd(&S[cur].sy, c * uy);
                    atomicAdd(&S[cur].sx, static_cast<unsigned long long>(sx));
                    atomicAdd(&S[cur].syy, c * uy * uy);
                    atomicAdd(&S[cur].sxx, static_cast<unsigned long long>(sxx));
                    atomicAdd(&S[cur].sxy, static_cast<unsigned long long>(sx) * uy);
                }
            }
            remaining &= ~grp;
        }
    }
    if (!any) continue;
    // obj_boundary = dilate3x3(obj) ^ obj : this pixel belongs to the boundary of every *other* gap it touches
    const unsigned nb = any & ~(1u << 4);
    if (!nb) continue;
    const double v = nraw[base + y * W + x];
    // the ids of the neighbouring gap pixels: independent (predicated) loads issued together -- the bit image says where
    // they are, so there is no load -> branch -> load chain per neighbour
    int o[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = (nb >> k) & 1u ? g[(y + k / 3 - 1) * W + x + k % 3 - 1] : -1;
    if (v == 0.0) continue;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (o[k] < 0 || o[k] == me) continue;
        bool dup = false;
#pragma unroll
        for (int j = 0; j < k; ++j) dup |= (o[j] == o[k]);
        if (!dup) atomicAdd(&S[o[k]].bsum, v);
    }
  }
}

// rescale + clip (:355-358)
__device__ __forceinline__ double rescale_clip(double v) {
    v = 1.0 / sqrt(0.65 + 0.5 * exp(-11.0 * (v - 0.75))) - 0.19;
    return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
}
// gap map (label_closed_corr) + max with raw neighbour map and touching borders + rescale + clip (:352-358) at one pixel
__device__ __forceinline__ double compose_at(const uint8_t *__restrict__ gp, const int *__restrict__ g, const GapStats *__restrict__ gsc,
                                             const uint8_t *__restrict__ border, const double *__restrict__ nraw, int H, int W, int y, int x,
                                             double at_zero) {
    const int me = gp[y * W + x] ? g[y * W + x] : -1;
    float corr = 0.0f;
    if (me >= 0) {
        const GapStats s = gsc[me];
        const double th = s.cnt <= 20 ? 5.0 : (s.cnt <= 30 ? 8.0 : (s.cnt <= 50 ? 10.0 : 20.0));   // :342-349
        if (s.bsum < th) {
            corr = 0.0f;                                    // artefact: completely in the background
        } else {
            corr = 1.0f;
            double major, minor;
            axis_lengths(s.cnt, s.sy, s.sx, s.syy, s.sxx, s.sxy, &major, &minor);
            if (minor >= 3.0) {
                // ring = gap ^ binary_erosion(gap, cross), border_value = 0
                auto same = [&](int i) { return gp[i] && g[i] == me; };
                const bool inner = y > 0 && y + 1 < H && x > 0 && x + 1 < W && same((y - 1) * W + x) && same((y + 1) * W + x) &&
                                   same(y * W + x - 1) && same(y * W + x + 1);
                if (!inner) corr = 0.8f;
            }
        }
    }
    double v = nraw[y * W + x];
    v = fmax(v, static_cast<double>(corr));
    v = fmax(v, border[y * W + x] ? 1.0 : 0.0);
    // ~85 % of the pixels are plain background (v == 0): they take the value the same instructions gave for 0 (computed
    // once per block from a run-time zero, so the compiler cannot fold it with host arithmetic) and skip ~100 FP64 ops
    if (v == 0.0) return at_zero;
    return rescale_clip(v);
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
    while (i < 0 || i >= n) {
        if (i < 0) i = -i - 1;
        if (i >= n) i = 2 * n - 1 - i;
    }
    return i;
}
// compose + grey_closing(size=(3,3)) in ONE pass (round 2: the float64 `scaled` map is never written; it cost 16 B/px of
// traffic and a launch).  grey closing = 3x3 max filter then 3x3 min filter, mode 'reflect' (edge sample duplicated).
// S holds compose(reflect(raw)) for the raw coordinates of the 32x32 tile with a 2-pixel apron (27 % recomputed apron);
// the dilated map D at a raw position h (1-pixel apron) is the 3x3 maximum around q = reflect(h), i.e. the value the min
// filter would read there; the output is the 3x3 minimum of D.
constexpr int GT = 32;
__global__ void __launch_bounds__(256)
lab_compose_closing_kernel(const uint8_t *__restrict__ gap, const int *__restrict__ gid, const GapStats *__restrict__ gs,
                           const uint8_t *__restrict__ border, const double *__restrict__ nraw, int H, int W, int max_gaps,
                           float *__restrict__ out, double zero) {
    // float32 rounding is monotone, so float(min(max(doubles))) == min(max(float(doubles))): the composed value is rounded
    // once and the two 3x3 filters run on floats (one FMNMX per comparison instead of ~10 instructions per fp64 max)
    __shared__ float S[GT + 4][GT + 4 + 1];
    __shared__ float D[GT + 2][GT + 2 + 1];
    __shared__ double s_at_zero;
    if (threadIdx.x == 0) s_at_zero = rescale_clip(zero);
    __syncthreads();
    const double at_zero = s_at_zero;
    const int crop = blockIdx.z;
    const int x0 = blockIdx.x * GT, y0 = blockIdx.y * GT;
    const int tid = threadIdx.x;
    const size_t base = static_cast<size_t>(crop) * H * W;
    const uint8_t *gp = gap + base, *bp = border + base;
    const int *g = gid + base;
    const double *np_ = nraw + base;
    const GapStats *gsc = gs + static_cast<size_t>(crop) * max_gaps;
    for (int i = tid; i < (GT + 4) * (GT + 4); i += 256) {
        const int r = i / (GT + 4), c = i - r * (GT + 4);
        S[r][c] = static_cast<float>(compose_at(gp, g, gsc, bp, np_, H, W, reflect_idx(y0 - 2 + r, H), reflect_idx(x0 - 2 + c, W), at_zero));
    }
    __syncthreads();
    for (int i = tid; i < (GT + 2) * (GT + 2); i += 256) {
        const int r = i / (GT + 2), c = i - r * (GT + 2);
        if (y0 - 1 + r > H || x0 - 1 + c > W) {      // beyond the 1-pixel apron of the image (ragged last tiles): never read
            D[r][c] = 0.0f;
            continue;
        }
        const int qy = reflect_idx(y0 - 1 + r, H) - (y0 - 2), qx = reflect_idx(x0 - 1 + c, W) - (x0 - 2);   // indices into S
        float v = S[qy][qx];
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) v = fmaxf(v, S[qy + dy][qx + dx]);
        D[r][c] = v;
    }
    __syncthreads();
    for (int i = tid; i < GT * GT; i += 256) {
        const int r = i / GT, c = i - r * GT;
        const int x = x0 + c, y = y0 + r;
        if (x >= W || y >= H) continue;
        float v = D[r + 1][c + 1];
#pragma unroll
        for (int dy = 0; dy <= 2; ++dy)
#pragma unroll
            for (int dx = 0; dx <= 2; ++dx) v = fminf(v, D[r + dy][c + dx]);
        out[base + static_cast<size_t>(y) * W + x] = v;
    }
}

inline size_t r256(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }
constexpr int kMaxGaps = 4096;

}  // namespace

namespace {
// boundary_label / border_label (train_data_representations.py:75-99 / :102-126) as ONE stencil pass instead of a
// full-image binary_dilation per nucleus: `boundary` = pixels with an 8-neighbour inside a nucleus they do not belong to
// (dilation(nucleus) ^ nucleus, OR-ed over the ids; binary_dilation's border_value = 0: outside the image is empty);
// border = boundary ^ (dilation(label > 0) ^ (label > 0)) = nucleus pixels touching a DIFFERENT nucleus.
// out = max(label > 0, 2 * boundary | border) as uint8.
__global__ void simple_label_kernel(const uint16_t *__restrict__ masks, int H, int W, int mode, uint8_t *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t base = static_cast<size_t>(blockIdx.z) * H * W;
    const uint16_t me = masks[base + static_cast<size_t>(y) * W + x];
    bool other = false;                 // an 8-neighbour belongs to a nucleus that is not mine
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            const uint16_t q = masks[base + static_cast<size_t>(yy) * W + xx];
            other |= (q != 0 && q != me);
        }
    }
    uint8_t v = me ? 1 : 0;
    if (mode == 0) {                    // boundary: background pixels next to a nucleus count too
        if (other) v = 2;
    } else {                            // border: only nucleus pixels (the outer ring dilation(bin) ^ bin cancels)
        if (other && me) v = 2;
    }
    out[base + static_cast<size_t>(y) * W + x] = v;
}
}  // namespace

namespace {
// j4_label (train_data_representations.py:158-190): 0 background, 1 cell, 2 touching, 3 gap.
// pass 1: binary_dilation(label > 0, disk(r)) (border_value = 0: outside the image is empty)
__global__ void j4_dilate_kernel(const uint16_t *__restrict__ masks, int H, int W, int r, uint8_t *__restrict__ dil) {
    const int crop = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const uint16_t *m = masks + static_cast<size_t>(crop) * H * W;
    bool any = false;
    for (int dy = -r; dy <= r && !any; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -r; dx <= r; ++dx) {
            if (dx * dx + dy * dy > r * r) continue;          // skimage.morphology.disk
            const int xx = x + dx;
            if (xx >= 0 && xx < W && m[yy * W + xx]) { any = true; break; }
        }
    }
    dil[(static_cast<size_t>(crop) * H + y) * W + x] = any ? 1 : 0;
}
// pass 2: binary_erosion of the dilation (border_value = 0: a structuring-element pixel outside the image erodes, scipy's
// binary_closing passes the same border value to both steps) -> bottom hat at background pixels; number of distinct
// instances in the (2k+1)^2 window (zero padded) > 1 -> touching
__global__ void j4_compose_kernel(const uint16_t *__restrict__ masks, const uint8_t *__restrict__ dil, int H, int W, int r, int k,
                                  uint8_t *__restrict__ out) {
    const int crop = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const uint16_t *m = masks + static_cast<size_t>(crop) * H * W;
    const uint8_t *d = dil + static_cast<size_t>(crop) * H * W;
    const int l = m[y * W + x];
    uint8_t v;
    if (l) {
        bool touching = false;                                // a second positive id in the window (it always holds l itself)
        for (int dy = -k; dy <= k && !touching; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            for (int dx = -k; dx <= k; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                const int o = m[yy * W + xx];
                if (o && o != l) { touching = true; break; }
            }
        }
        v = touching ? 2 : 1;
    } else {
        bool closed = true;
        for (int dy = -r; dy <= r && closed; ++dy)
            for (int dx = -r; dx <= r; ++dx) {
                if (dx * dx + dy * dy > r * r) continue;
                const int yy = y + dy, xx = x + dx;
                if (yy < 0 || yy >= H || xx < 0 || xx >= W || !d[yy * W + xx]) { closed = false; break; }
            }
        v = closed ? 3 : 0;
    }
    out[(static_cast<size_t>(crop) * H + y) * W + x] = v;
}
}  // namespace

extern "C" int mbs_j4_labels(const uint16_t *masks, int n_crops, int H, int W, int k_neighbors, int se_radius, uint8_t *out,
                             uint8_t *tmp, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(masks && out && tmp && n_crops > 0 && n_crops <= 65535 && H > 0 && W > 0 && k_neighbors >= 0 && k_neighbors <= 16 &&
                    se_radius >= 0 && se_radius <= 16,
                "j4_labels: bad arguments");
    dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8, n_crops);
    j4_dilate_kernel<<<grid, block, 0, stream>>>(masks, H, W, se_radius, tmp);
    MBS_CHECK_LAUNCH();
    j4_compose_kernel<<<grid, block, 0, stream>>>(masks, tmp, H, W, se_radius, k_neighbors, out);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_boundary_border_labels(const uint16_t *masks, int n_crops, int H, int W, int mode, uint8_t *out,
                                          void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(masks && out && n_crops > 0 && n_crops <= 65535 && H > 0 && W > 0 && (mode == 0 || mode == 1),
                "boundary_border_labels: bad arguments");
    dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8, n_crops);
    simple_label_kernel<<<grid, block, 0, stream>>>(masks, H, W, mode, out);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t mbs_labels_workspace_bytes(int n_crops, int H, int W, int max_id) {
    const size_t px = static_cast<size_t>(n_crops) * H * W;
    const size_t ids = static_cast<size_t>(max_id) + 1;
    return r256(n_crops * ids * sizeof(CellStats)) + r256(n_crops * sizeof(CropInfo)) +
           r256(static_cast<size_t>(n_crops) * kMaxGaps * sizeof(GapStats)) + 2 * r256(px * 8) /*nraw, scaled*/ +
           2 * r256(px) /*gap, border*/ + 2 * r256(px * 4) /*L, gid*/ +
           2 * r256(static_cast<size_t>(n_crops) * H * ((W + 63) / 64) * 8) /*label_bin, dil as bit images*/ +
           r256((n_crops * ids + 1) * sizeof(int)) /*oversized-instance list*/ + 4096;
}

extern "C" int mbs_labels_max_mal(const uint16_t *masks, int n_crops, int H, int W, int max_id, int32_t *max_mal_out,
                                  void *workspace, size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(n_crops > 0 && H > 0 && W > 0 && max_id >= 0 && max_id <= 65535 && max_mal_out, "labels_max_mal: bad arguments");
    const int ids = max_id + 1;
    const size_t need = r256(static_cast<size_t>(n_crops) * ids * sizeof(CellStats)) + r256(n_crops * sizeof(CropInfo));
    MBS_REQUIRE(workspace_bytes >= need, "labels_max_mal: workspace too small");
    CellStats *cs = reinterpret_cast<CellStats *>(workspace);
    CropInfo *info = reinterpret_cast<CropInfo *>(static_cast<char *>(workspace) +
                                                  r256(static_cast<size_t>(n_crops) * ids * sizeof(CellStats)));
    const int total = n_crops * ids;
    dim3 b2(32, 8), g3(mbs::cdiv(W, 32), mbs::cdiv(H, kRowsPerBlock), n_crops);
    MBS_CHECK_CUDA(cudaMemsetAsync(info, 0, n_crops * sizeof(CropInfo), stream));
    lab_init_stats_kernel<<<mbs::cdiv(total, 256), 256, 0, stream>>>(cs, total);
    MBS_CHECK_LAUNCH();
    lab_accum_kernel<<<g3, b2, 0, stream>>>(masks, H, W, ids, cs);
    MBS_CHECK_LAUNCH();
    lab_axes_kernel<<<mbs::cdiv(total, 256), 256, 0, stream>>>(cs, ids, total, info);
    MBS_CHECK_LAUNCH();
    MBS_CHECK_CUDA(cudaMemcpy2DAsync(max_mal_out, sizeof(int), &info[0].max_mal, sizeof(CropInfo), sizeof(int), n_crops,
                                     cudaMemcpyDeviceToDevice, stream));
    return 0;
}

// per-frame analysis (src/inference/analysis.py:141-170): area and axis lengths of every instance
namespace {
__global__ void lab_export_stats_kernel(const CellStats *cs, int total, int32_t *area, double *major, double *minor) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const CellStats s = cs[i];
    double ma = 0.0, mi = 0.0;
    if (s.cnt) axis_lengths(s.cnt, s.sy, s.sx, s.syy, s.sxx, s.sxy, &ma, &mi);
    area[i] = static_cast<int32_t>(s.cnt);
    major[i] = ma;
    minor[i] = mi;
}
}  // namespace

extern "C" int mbs_instance_stats(const uint16_t *masks, int n_frames, int H, int W, int max_id, int32_t *area, double *major,
                                  double *minor, void *workspace, size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(n_frames > 0 && H > 0 && W > 0 && max_id >= 0 && max_id <= 65535 && area && major && minor,
                "instance_stats: bad arguments");
    const int ids = max_id + 1;
    const size_t need = r256(static_cast<size_t>(n_frames) * ids * sizeof(CellStats));
    MBS_REQUIRE(workspace_bytes >= need, "instance_stats: workspace too small (%zu < %zu)", workspace_bytes, need);
    MBS_REQUIRE(static_cast<long long>(n_frames) * ids < (1ll << 31), "instance_stats: too many (frame, id) pairs");
    CellStats *cs = reinterpret_cast<CellStats *>(workspace);
    const int total = n_frames * ids;
    dim3 b2(32, 8), g3(mbs::cdiv(W, 32), mbs::cdiv(H, kRowsPerBlock), n_frames);
    lab_init_stats_kernel<<<mbs::cdiv(total, 256), 256, 0, stream>>>(cs, total);
    MBS_CHECK_LAUNCH();
    lab_accum_kernel<<<g3, b2, 0, stream>>>(masks, H, W, ids, cs);
    MBS_CHECK_LAUNCH();
    lab_export_stats_kernel<<<mbs::cdiv(total, 256), 256, 0, stream>>>(cs, total, area, major, minor);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_distance_labels(const uint16_t *masks, int n_crops, int H, int W, int max_id, int search_radius,
                                   int radius_hint, float *cell_dist, float *neighbor_dist, int32_t *max_mal_out, int32_t *error_out,
                                   void *workspace, size_t workspace_bytes, void *stream_) {
    return mbs_distance_labels_ex(masks, n_crops, H, W, max_id, search_radius, radius_hint, 0.0f, cell_dist, neighbor_dist, max_mal_out,
                                  error_out, workspace, workspace_bytes, stream_);
}

extern "C" int mbs_distance_labels_ex(const uint16_t *masks, int n_crops, int H, int W, int max_id, int search_radius,
                                      int radius_hint, float cell_clip, float *cell_dist, float *neighbor_dist, int32_t *max_mal_out,
                                      int32_t *error_out, void *workspace, size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(n_crops > 0 && H > 0 && W > 0 && H < 32768 && W < 32768 && max_id >= 0 && max_id <= 65535, "distance_labels: bad arguments");
    const size_t px = static_cast<size_t>(n_crops) * H * W;
    MBS_REQUIRE(px < (1ull << 31), "distance_labels: batch too large (use fewer crops per call)");
    MBS_REQUIRE(workspace_bytes >= mbs_labels_workspace_bytes(n_crops, H, W, max_id), "distance_labels: workspace too small");
    const int ids = max_id + 1;
    char *p = static_cast<char *>(workspace);
    auto take = [&](size_t bytes) { char *r = p; p += r256(bytes); return r; };
    CellStats *cs = reinterpret_cast<CellStats *>(take(static_cast<size_t>(n_crops) * ids * sizeof(CellStats)));
    CropInfo *info = reinterpret_cast<CropInfo *>(take(n_crops * sizeof(CropInfo)));
    GapStats *gs = reinterpret_cast<GapStats *>(take(static_cast<size_t>(n_crops) * kMaxGaps * sizeof(GapStats)));
    double *nraw = reinterpret_cast<double *>(take(px * 8));
    double *scaled = reinterpret_cast<double *>(take(px * 8));
    uint8_t *gap = reinterpret_cast<uint8_t *>(take(px));
    uint8_t *border = reinterpret_cast<uint8_t *>(take(px));
    int *L = reinterpret_cast<int *>(take(px * 4));
    int *gid = reinterpret_cast<int *>(take(px * 4));
    const int WW = (W + 63) / 64;
    const size_t bit_bytes = static_cast<size_t>(n_crops) * H * WW * 8;
    unsigned long long *lbits = reinterpret_cast<unsigned long long *>(take(bit_bytes));
    unsigned long long *dbits = reinterpret_cast<unsigned long long *>(take(bit_bytes));
    int *big_list = reinterpret_cast<int *>(take((static_cast<size_t>(n_crops) * ids + 1) * sizeof(int)));

    const int total = n_crops * ids;
    dim3 b2(32, 8), g3(mbs::cdiv(W, 32), mbs::cdiv(H, kRowsPerBlock), n_crops);
    MBS_CHECK_CUDA(cudaMemsetAsync(info, 0, n_crops * sizeof(CropInfo), stream));
    MBS_CHECK_CUDA(cudaMemsetAsync(gs, 0, static_cast<size_t>(n_crops) * kMaxGaps * sizeof(GapStats), stream));
    MBS_CHECK_CUDA(cudaMemsetAsync(nraw, 0, px * 8, stream));
    MBS_CHECK_CUDA(cudaMemsetAsync(cell_dist, 0, px * 4, stream));
    MBS_CHECK_CUDA(cudaMemsetAsync(lbits, 0, bit_bytes, stream));
    MBS_CHECK_CUDA(cudaMemsetAsync(big_list, 0, sizeof(int), stream));
    lab_init_stats_kernel<<<mbs::cdiv(total, 256), 256, 0, stream>>>(cs, total);
    MBS_CHECK_LAUNCH();
    lab_accum_kernel<<<g3, b2, 0, stream>>>(masks, H, W, ids, cs);
    MBS_CHECK_LAUNCH();
    lab_axes_kernel<<<mbs::cdiv(total, 256), 256, 0, stream>>>(cs, ids, total, info);
    MBS_CHECK_LAUNCH();
    lab_window_kernel<<<mbs::cdiv(total, 256), 256, 0, stream>>>(cs, ids, total, H, W, info, search_radius);
    MBS_CHECK_LAUNCH();
    if (max_id > 0) {
        // dynamic shared memory: as much as the device allows (opt-in), the kernels flag windows that do not fit
        static int smem_opt_dev[mbs::kMaxDevices] = {0};
        const int dev = mbs::current_device();
        if (!smem_opt_dev[dev]) {
            int v = 0;
            cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
            v -= 1024;
            MBS_CHECK_CUDA(cudaFuncSetAttribute(lab_cell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
            MBS_CHECK_CUDA(cudaFuncSetAttribute(lab_close_cell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
            smem_opt_dev[dev] = v;
        }
        const int smem_opt = smem_opt_dev[dev];
        // size the launch for the largest window the radius can produce, capped by the device limit
        const int R = search_radius >= 0 ? search_radius : (radius_hint >= 0 ? radius_hint : (H > W ? H : W));
        long long want = 4ll * 2 * (2ll * R) * (2ll * R);
        const long long full = 4ll * 2 * H * W;
        if (want > full) want = full;
        int smem = want > smem_opt ? smem_opt : static_cast<int>(want);
        if (smem < 4096) smem = 4096;
        dim3 gc(max_id, n_crops);
        // latency-bound CTAs (serial column scans, ~300 own pixels per instance): small blocks, many resident per SM
        static int cell_threads = 0;
        if (!cell_threads) {
            const char *e = getenv("MBS_LAB_CELL_THREADS");       // A/B knob
            cell_threads = e ? atoi(e) : 128;
            if (cell_threads < 32 || cell_threads > 256 || cell_threads % 32) cell_threads = 128;   // whole warps (full-mask ballots)
        }
        lab_cell_kernel<<<gc, cell_threads, smem, stream>>>(masks, H, W, ids, cs, info, cell_dist, nraw, smem / 8, cell_clip);
        MBS_CHECK_LAUNCH();
        if (full > smem)         // windows larger than shared memory are possible: the global-memory walk picks them up
            lab_cell_big_kernel<<<n_crops, 256, 0, stream>>>(masks, H, W, ids, cs, reinterpret_cast<unsigned short *>(scaled),
                                                             static_cast<size_t>(H) * W * 4, cell_dist, nraw, smem / 8, cell_clip);
        MBS_CHECK_LAUNCH();
        long long wantc = 2ll * (static_cast<long long>(H) + 6) * (W + 6);
        int smemc = wantc > smem_opt ? smem_opt : static_cast<int>(wantc);
        lab_close_cell_bits_kernel<<<dim3(mbs::cdiv(max_id, 8), n_crops), 256, 0, stream>>>(masks, H, W, WW, ids, cs, lbits, big_list);
        MBS_CHECK_LAUNCH();
        // oversized instances only: a small persistent grid walks the queue (usually empty)
        lab_close_cell_kernel<<<2 * 148, 256, smemc, stream>>>(masks, H, W, WW, ids, cs, info, lbits, smemc, big_list);
        MBS_CHECK_LAUNCH();
    }
    lab_border_kernel<<<dim3(mbs::cdiv(mbs::cdiv(W, 8) * H, 256), 1, n_crops), 256, 0, stream>>>(masks, H, W, border);
    MBS_CHECK_LAUNCH();
    const int nwords = static_cast<int>((static_cast<long long>(n_crops) * H * WW + 255) / 256);
    lab_dilate_bits_kernel<<<nwords, 256, 0, stream>>>(lbits, H, W, WW, n_crops, dbits);
    MBS_CHECK_LAUNCH();
    lab_erode_gap_bits_kernel<<<nwords, 256, 0, stream>>>(dbits, lbits, H, W, WW, n_crops, gap, L);
    MBS_CHECK_LAUNCH();
    const int nb16 = static_cast<int>((px + 16 * 256 - 1) / (16 * 256));
    gap_merge_kernel<<<nb16, 256, 0, stream>>>(gap, static_cast<long long>(px), H, W, L);
    MBS_CHECK_LAUNCH();
    gap_resolve_kernel<<<nb16, 256, 0, stream>>>(gap, L, static_cast<long long>(px));
    MBS_CHECK_LAUNCH();
    gap_compress_ids_kernel<<<nb16, 256, 0, stream>>>(gap, L, static_cast<long long>(px), H * W, kMaxGaps, info, gid);
    MBS_CHECK_LAUNCH();
    gap_assign_kernel<<<nb16, 256, 0, stream>>>(gap, L, static_cast<long long>(px), gid);
    MBS_CHECK_LAUNCH();
    gap_stats_kernel<<<g3, b2, 0, stream>>>(gap, gid, nraw, lbits, WW, H, W, kMaxGaps, gs);
    MBS_CHECK_LAUNCH();
    lab_compose_closing_kernel<<<dim3(mbs::cdiv(W, GT), mbs::cdiv(H, GT), n_crops), 256, 0, stream>>>(gap, gid, gs, border, nraw, H, W, kMaxGaps,
                                                                                                  neighbor_dist, 0.0);
    MBS_CHECK_LAUNCH();
    if (max_mal_out)
        MBS_CHECK_CUDA(cudaMemcpy2DAsync(max_mal_out, sizeof(int), &info[0].max_mal, sizeof(CropInfo), sizeof(int), n_crops,
                                         cudaMemcpyDeviceToDevice, stream));
    if (error_out)
        MBS_CHECK_CUDA(cudaMemcpy2DAsync(error_out, sizeof(int), &info[0].error, sizeof(CropInfo), sizeof(int), n_crops,
                                         cudaMemcpyDeviceToDevice, stream));
    return 0;
}
