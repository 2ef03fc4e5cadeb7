// Mask -> polygon step that follows the segmentation path on the OMERO route.
//
// Replaces, per frame, get_indices_pandas + one cv2.findContours call per cell
// (/root/reference/src/utils/hull_polygon.py:8-89, called from src/inference/infer.py:273-289): the outer contour of
// every instance of a uint16 mask in OpenCV's CHAIN_APPROX_NONE point order (Suzuki-Abe border following,
// 8-connectivity, start = first raster pixel of the instance).  One thread follows one instance; instances are
// independent, so a frame with thousands of cells keeps the machine busy.  Two passes over the same walk: count the
// points per instance, (exclusive scan on the host side), write them.
#include <cuda_runtime.h>

#include <climits>
#include <cstdint>

#include "../../include/mbseg.h"
#include "common.cuh"

namespace {

// first pixel (raster order) of every instance id: the point where OpenCV's scan meets the outer border
__global__ void contour_first_kernel(const uint16_t *__restrict__ mask, long long n, int n_labels, int *first) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const int id = mask[i];
    if (id > 0 && id <= n_labels) atomicMin(&first[id - 1], static_cast<int>(i));
}

__global__ void contour_fill_kernel(int *first, int n_labels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_labels) first[i] = INT_MAX;
}

// OpenCV direction codes (x, y): 0 E, 1 NE, 2 N, 3 NW, 4 W, 5 SW, 6 S, 7 SE
__device__ __constant__ int c_dx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
__device__ __constant__ int c_dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};

template <bool WRITE>
__global__ void contour_trace_kernel(const uint16_t *__restrict__ mask, int H, int W, int n_labels, const int *__restrict__ first,
                                     const long long *__restrict__ offsets, int *__restrict__ counts, int *__restrict__ points_yx,
                                     int *overflow) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_labels) return;
    const int f = first[l];
    if (f == INT_MAX) {
        if (!WRITE) counts[l] = 0;
        return;
    }
    const uint16_t id = static_cast<uint16_t>(l + 1);
    auto fg = [&](int x, int y) -> bool { return x >= 0 && x < W && y >= 0 && y < H && mask[static_cast<size_t>(y) * W + x] == id; };
    const int x0 = f % W, y0 = f / W;
    int *out = WRITE ? points_yx + 2 * offsets[l] : nullptr;
    // icvFetchContour (outer border): clockwise search for the last border pixel i1, starting after the west neighbour
    int s = 4;
    const int s_stop = 4;
    int x1 = x0, y1 = y0;
    bool found = false;
    do {
        s = (s - 1) & 7;
        x1 = x0 + c_dx[s];
        y1 = y0 + c_dy[s];
        if (fg(x1, y1)) {
            found = true;
            break;
        }
    } while (s != s_stop);
    int n = 0;
    if (!found) {                                   // single pixel
        if (WRITE) {
            out[0] = y0;
            out[1] = x0;
        }
        n = 1;
    } else {
        int x3 = x0, y3 = y0;
        const long long guard = 8ll * H * W + 16;   // a border is visited at most a few times per pixel
        for (long long it = 0; it < guard; ++it) {
            int x4, y4;
            for (;;) {                              // counter-clockwise search, starts after the direction we came from
                ++s;
                x4 = x3 + c_dx[s & 7];
                y4 = y3 + c_dy[s & 7];
                if (fg(x4, y4)) break;
            }
            s &= 7;
            if (WRITE) {
                out[2 * n] = y3;
                out[2 * n + 1] = x3;
            }
            ++n;
            if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) break;
            x3 = x4;
            y3 = y4;
            s = (s + 4) & 7;
            if (it + 1 == guard) atomicExch(overflow, 1);
        }
    }
    if (!WRITE) counts[l] = n;
}

}  // namespace

extern "C" int mbs_contour_first(const uint16_t *mask, int H, int W, int n_labels, int32_t *first, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long n = static_cast<long long>(H) * W;
    MBS_REQUIRE(H > 0 && W > 0 && n < (1ll << 31) && n_labels >= 0, "contour_first: bad shape H=%d W=%d labels=%d", H, W, n_labels);
    if (n_labels == 0) return 0;
    contour_fill_kernel<<<mbs::cdiv(n_labels, 256), 256, 0, stream>>>(first, n_labels);
    MBS_CHECK_LAUNCH();
    contour_first_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(mask, n, n_labels, first);
    MBS_CHECK_LAUNCH();
    return 0;
}

extern "C" int mbs_contour_trace(const uint16_t *mask, int H, int W, int n_labels, const int32_t *first, const int64_t *offsets,
                                 int32_t *counts, int32_t *points_yx, int32_t *overflow, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    MBS_REQUIRE(H > 0 && W > 0 && n_labels >= 0, "contour_trace: bad shape");
    if (n_labels == 0) return 0;
    const int blocks = mbs::cdiv(n_labels, 64);
    if (offsets == nullptr) {
        MBS_REQUIRE(counts != nullptr, "contour_trace: the counting pass needs `counts`");
        contour_trace_kernel<false><<<blocks, 64, 0, stream>>>(mask, H, W, n_labels, first, nullptr, counts, nullptr, overflow);
    } else {
        MBS_REQUIRE(points_yx != nullptr, "contour_trace: the writing pass needs `points_yx`");
        contour_trace_kernel<true><<<blocks, 64, 0, stream>>>(mask, H, W, n_labels, first,
                                                              reinterpret_cast<const long long *>(offsets), counts, points_yx, overflow);
    }
    MBS_CHECK_LAUNCH();
    return 0;
}
