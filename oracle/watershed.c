/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY. Never imported by the product path.
 *
 * CPU restatement of the marker-controlled watershed that the reference calls at
 *   src/inference/postprocessing.py:57  (distance_postprocessing)
 *   src/inference/postprocessing.py:88  (boundary_postprocessing)
 * i.e. skimage.segmentation.watershed(image, markers, mask=mask, watershed_line=False),
 * connectivity=1, compactness=0, scikit-image 0.19.0 (requirements.yml:24).
 *
 * scikit-image is NOT vendored in /root/reference and is not installable in this image, so
 * this file restates its published algorithm (skimage/segmentation/_watershed.py,
 * _watershed_cy.pyx::watershed_raveled, ../_shared/heap_general.pxi + heap_watershed.pxi):
 *   - image -> float64, markers*mask -> int32, all arrays zero padded by one pixel;
 *   - every marker pixel is pushed in raster order with (value, age=0);
 *   - repeat: pop the smallest (value, then age); for each neighbour in the order
 *     up, left, right, down (the C-order of the connectivity-1 footprint; the reference passes
 *     (H,W,1) arrays, whose z-neighbours always fall into the zero padding and are skipped
 *     before the age counter is touched): skip if outside mask or already labelled, otherwise
 *     age += 1, label the neighbour with the popped pixel's label and push (value, age);
 *   - the heap is an array binary heap: push appends and sifts up while strictly smaller than
 *     the parent; pop moves the last item to the root and sifts down, choosing the left child
 *     if it is strictly smaller than the current node and then the right child if it is
 *     strictly smaller than that choice.
 * PARITY UNPINNED: the reference has no tests/golden vectors for this path and skimage cannot
 * be executed here; the restatement is cross-checked against an independent order-free
 * formulation (oracle/postproc.py::watershed_minimax) on tie-free inputs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double value;
    int64_t age;
    int64_t index;
} item_t;

typedef struct {
    item_t *d;
    int64_t n, cap;
} heap_t;

static inline int smaller(const item_t *a, const item_t *b) {
    if (a->value != b->value) return a->value < b->value;
    return a->age < b->age;
}

static int heap_push(heap_t *h, const item_t *e) {
    if (h->n == h->cap) {
        int64_t nc = h->cap * 2;
        item_t *nd = (item_t *)realloc(h->d, (size_t)nc * sizeof(item_t));
        if (!nd) return -1;
        h->d = nd;
        h->cap = nc;
    }
    int64_t child = h->n++;
    h->d[child] = *e;
    while (child > 0) {
        int64_t parent = (child + 1) / 2 - 1;
        if (smaller(&h->d[child], &h->d[parent])) {
            item_t t = h->d[parent];
            h->d[parent] = h->d[child];
            h->d[child] = t;
            child = parent;
        } else {
            break;
        }
    }
    return 0;
}

static void heap_pop(heap_t *h, item_t *dst) {
    *dst = h->d[0];
    h->n -= 1;
    if (h->n == 0) return;
    h->d[0] = h->d[h->n];
    int64_t i = 0, smallest = 0;
    for (;;) {
        int64_t l = 2 * i + 1, r = 2 * i + 2;
        if (l < h->n) {
            if (smaller(&h->d[l], &h->d[i])) smallest = l;
            if (r < h->n && smaller(&h->d[r], &h->d[smallest])) smallest = r;
        } else {
            break;
        }
        if (smallest == i) break;
        item_t t = h->d[i];
        h->d[i] = h->d[smallest];
        h->d[smallest] = t;
        i = smallest;
    }
}

/*
 * image   : H*W float64 (already negated by the caller where the reference negates)
 * markers : H*W int32 (already multiplied by mask, as _validate_inputs does)
 * mask    : H*W uint8
 * out     : H*W int32
 * marker_order: 0 = true skimage behaviour (equal-valued age-0 markers pop in heap-internal
 *               order); 1 = canonical variant: marker pixels get age = -(N_markers) + rank so
 *               equal-valued markers pop in raster order (used to quantify how often the
 *               implementation-defined order matters).
 * returns 0 on success.
 */
int oracle_watershed(const double *image, const int32_t *markers, const uint8_t *mask,
                     int32_t *out, int64_t H, int64_t W, int marker_order) {
    const int64_t Wp = W + 2, Hp = H + 2;
    double *img = (double *)calloc((size_t)(Hp * Wp), sizeof(double));
    uint8_t *msk = (uint8_t *)calloc((size_t)(Hp * Wp), 1);
    int32_t *lab = (int32_t *)calloc((size_t)(Hp * Wp), sizeof(int32_t));
    if (!img || !msk || !lab) return -1;
    int64_t nmark = 0;
    for (int64_t y = 0; y < H; ++y)
        for (int64_t x = 0; x < W; ++x) {
            int64_t p = (y + 1) * Wp + (x + 1);
            img[p] = image[y * W + x];
            msk[p] = mask[y * W + x] ? 1 : 0;
            lab[p] = markers[y * W + x];
            if (lab[p]) nmark++;
        }
    heap_t h;
    h.cap = nmark > 1024 ? nmark * 2 : 2048;
    h.n = 0;
    h.d = (item_t *)malloc((size_t)h.cap * sizeof(item_t));
    if (!h.d) return -1;
    const int64_t nb[4] = {-Wp, -1, +1, +Wp}; /* up, left, right, down */
    int64_t age = 1;
    item_t e, ne;
    int64_t rank = 0;
    for (int64_t p = 0; p < Hp * Wp; ++p) {
        if (lab[p]) {
            e.value = img[p];
            e.age = marker_order ? (rank - nmark) : 0;
            e.index = p;
            rank++;
            if (heap_push(&h, &e)) return -1;
        }
    }
    while (h.n > 0) {
        heap_pop(&h, &e);
        for (int k = 0; k < 4; ++k) {
            int64_t q = e.index + nb[k];
            if (!msk[q]) continue;
            if (lab[q]) continue;
            age += 1;
            lab[q] = lab[e.index];
            ne.value = img[q];
            ne.age = age;
            ne.index = q;
            if (heap_push(&h, &ne)) return -1;
        }
    }
    for (int64_t y = 0; y < H; ++y)
        for (int64_t x = 0; x < W; ++x) out[y * W + x] = lab[(y + 1) * Wp + (x + 1)];
    free(h.d);
    free(img);
    free(msk);
    free(lab);
    return 0;
}
