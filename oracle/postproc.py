"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement of the reference post-processing operators

  * ``distance_postprocessing``  -- /root/reference/src/inference/postprocessing.py:7-59
  * ``boundary_postprocessing``  -- /root/reference/src/inference/postprocessing.py:62-90

The reference file itself cannot be imported here: it needs scikit-image 0.19.0
(requirements.yml:24) which is neither vendored in /root/reference nor installable offline.
What *is* the reference's own dependency and is present (scipy.ndimage.gaussian_filter,
postprocessing.py:25) is called directly; the scikit-image pieces are restated:

  * ``measure.label(.., background=0)`` on the (H,W,1) arrays the callers pass
    (postprocessing.py:38,54) == 8-connectivity in the plane, ids 1..n in raster order of each
    component's first pixel  ->  ``scipy.ndimage.label(structure=ones((3,3)))``.
  * ``measure.regionprops(..).area`` (postprocessing.py:41-45) -> pixel counts.
  * ``segmentation.watershed`` (postprocessing.py:57) -> oracle/watershed.c (heap flood).

PARITY UNPINNED for the scikit-image pieces: the reference ships no tests, golden vectors or
fixtures for this path and scikit-image cannot run in this image. The
scipy pieces are the real thing.  The heap flood is cross-checked against an independent,
order-free formulation (``watershed_minimax`` below) on tie-free inputs.

``np.tan`` on float32 is host-SIMD dependent.  ``tan_mode='f64'`` (the parity
mode, default) evaluates ``float32(tan(float64(x)))``; ``tan_mode='host'`` is the raw
``np.tan(float32)`` the reference would execute on this host.
"""
import ctypes
import os
import subprocess

import numpy as np
from scipy import ndimage

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# scipy.ndimage._filters._gaussian_kernel1d(sigma=0.5, order=0, radius=2) -- radius is
# int(4.0 * 0.5 + 0.5) = 2 (scipy truncate default 4.0).  Checked against scipy in tests.
GAUSS_SIGMA = 0.5
GAUSS_RADIUS = 2


def gaussian_weights():
    x = np.arange(-GAUSS_RADIUS, GAUSS_RADIUS + 1)
    phi = np.exp(-0.5 / (GAUSS_SIGMA * GAUSS_SIGMA) * x ** 2)
    return phi / phi.sum()


def build_lib(force=False):
    """gcc the C heap flood into oracle/liboracle.so (git-ignored build artefact)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, "watershed.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so] + srcs)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_lib())
        _LIB.oracle_watershed.restype = ctypes.c_int
        _LIB.oracle_watershed.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
    return _LIB


# ------------------------------------------------------------------------------------------
# pieces
# ------------------------------------------------------------------------------------------
def gaussian_smooth_emulated(cell2d):
    """NumPy emulation of scipy.ndimage.gaussian_filter(float32 (H,W,1), 0.5).

    Follows scipy/ndimage/src/ni_filters.c::NI_Correlate1D (symmetric branch): per output
    sample ``tmp = x0*w0; tmp += (x[-2]+x[+2])*w[-2]; tmp += (x[-1]+x[+1])*w[-1]`` in float64,
    rounded to float32 after each axis; axis 0 first, then axis 1; mode 'reflect'
    (d c b a | a b c d).  Used to document the arithmetic the CUDA kernel must reproduce.
    """
    w = gaussian_weights()
    w0, w1, w2 = w[2], w[1], w[0]
    a = np.asarray(cell2d, dtype=np.float32)
    for axis in (0, 1):
        x = np.moveaxis(a, axis, 0).astype(np.float64)
        xp = np.concatenate([x[1::-1], x, x[:-3:-1]], axis=0) if x.shape[0] >= 2 else np.pad(
            x, ((2, 2),) + ((0, 0),) * (x.ndim - 1), mode="symmetric")
        n = x.shape[0]
        c = xp[2:2 + n]
        tmp = c * w0
        tmp = tmp + (xp[0:n] + xp[4:4 + n]) * w2
        tmp = tmp + (xp[1:1 + n] + xp[3:3 + n]) * w1
        a = np.moveaxis(tmp.astype(np.float32), 0, axis)
    return np.ascontiguousarray(a)


def tan_f32(x, tan_mode="f64"):
    x = np.asarray(x, dtype=np.float32)
    if tan_mode == "host":
        return np.tan(x)
    return np.tan(x.astype(np.float64)).astype(np.float32)


def label8(binary2d):
    """measure.label(background=0) on (H,W,1) bool == 8-connectivity, raster-first numbering."""
    lab, n = ndimage.label(binary2d, structure=np.ones((3, 3), dtype=bool))
    return lab.astype(np.int32), int(n)


def watershed(image2d, markers2d, mask2d, marker_order=0):
    """skimage.segmentation.watershed(image, markers, mask=mask) restated (oracle/watershed.c)."""
    image = np.ascontiguousarray(image2d, dtype=np.float64)
    mask = np.ascontiguousarray(mask2d, dtype=bool)
    markers = np.ascontiguousarray(np.asarray(markers2d) * mask, dtype=np.int32)
    mask_u8 = np.ascontiguousarray(mask.astype(np.uint8))
    out = np.zeros(image.shape, dtype=np.int32)
    H, W = image.shape
    rc = _lib().oracle_watershed(image.ctypes.data, markers.ctypes.data, mask_u8.ctypes.data,
                                 out.ctypes.data, H, W, int(marker_order))
    if rc != 0:
        raise MemoryError("oracle_watershed failed")
    return out


def watershed_minimax(image2d, markers2d, mask2d, max_iter=100000):
    """Independent, order-free formulation of the same flood, valid when no two *distinct* pixels
    with equal value compete for a pixel (tie-free inputs).

    L(p) = flood level at which p is popped = minimax path cost from the marker set:
    L(m) = v(m) on markers, L(p) = max(v(p), min_{q in N4(p)} L(q)).  A pixel is labelled by the
    first popped neighbour, which has the minimum L; equal-L neighbours carry the same label when
    values are tie-free.  Returns (labels, ambiguous) where ``ambiguous`` flags pixels whose
    minimal-L neighbours disagree (only possible with exact value ties).
    """
    v = np.asarray(image2d, dtype=np.float64)
    mask = np.asarray(mask2d, dtype=bool)
    lab = (np.asarray(markers2d) * mask).astype(np.int64)
    H, W = v.shape
    INF = np.inf
    L = np.where(lab > 0, v, INF)
    ismark = lab > 0

    def nbr_min(A):
        P = np.pad(A, 1, constant_values=INF)
        return np.minimum(np.minimum(P[:-2, 1:-1], P[2:, 1:-1]), np.minimum(P[1:-1, :-2], P[1:-1, 2:]))

    for _ in range(max_iter):
        cand = np.maximum(v, nbr_min(L))
        newL = np.where(ismark, L, np.where(mask, np.minimum(L, cand), INF))
        if np.array_equal(newL, L):
            break
        L = newL
    # propagate labels: plateau-aware.  Pixels take the label of a strictly-lower-L neighbour if one
    # exists, else they belong to an equal-L plateau and inherit from plateau members that do.
    out = lab.copy()
    ambiguous = np.zeros((H, W), dtype=bool)
    reach = np.isfinite(L) & mask
    order = np.argsort(L, axis=None, kind="stable")
    Lf = L.ravel()
    outf = out.ravel()
    reachf = reach.ravel()
    # process by increasing L; inside one L value run a BFS from already-labelled pixels
    i = 0
    n = order.size
    offs = [(-1, 0), (0, -1), (0, 1), (1, 0)]
    while i < n and np.isfinite(Lf[order[i]]):
        j = i
        while j < n and Lf[order[j]] == Lf[order[i]]:
            j += 1
        group = [p for p in order[i:j] if reachf[p]]
        pending = set(p for p in group if outf[p] == 0)
        # first: pixels with a lower-L labelled neighbour
        frontier = []
        for p in list(pending):
            y, x = divmod(p, W)
            labs = set()
            lmin = INF
            for dy, dx in offs:
                yy, xx = y + dy, x + dx
                if 0 <= yy < H and 0 <= xx < W:
                    q = yy * W + xx
                    if reachf[q] and Lf[q] < Lf[p] and outf[q] > 0:
                        if Lf[q] < lmin:
                            lmin = Lf[q]
                            labs = set()
                        if Lf[q] == lmin:
                            labs.add(outf[q])
            if labs:
                if len(labs) > 1:
                    ambiguous[y, x] = True
                outf[p] = min(labs)
                frontier.append(p)
                pending.discard(p)
        frontier += [p for p in group if outf[p] > 0 and p not in frontier]
        while frontier and pending:
            nxt = []
            for p in frontier:
                y, x = divmod(p, W)
                for dy, dx in offs:
                    yy, xx = y + dy, x + dx
                    if 0 <= yy < H and 0 <= xx < W:
                        q = yy * W + xx
                        if q in pending:
                            outf[q] = outf[p]
                            pending.discard(q)
                            nxt.append(q)
                        elif reachf[q] and Lf[q] == Lf[p] and outf[q] > 0 and outf[q] != outf[p]:
                            ambiguous[y, x] = True
            frontier = nxt
        i = j
    return out.astype(np.int32), ambiguous


# ------------------------------------------------------------------------------------------
# operators
# ------------------------------------------------------------------------------------------
def _squeeze_hw(a):
    a = np.asarray(a)
    if a.ndim == 3 and a.shape[-1] == 1:
        a = a[..., 0]
    return a


def seed_mask_maps(border_prediction, cell_prediction, th_seed, th_cell, tan_mode="f64"):
    """postprocessing.py:25-37 -> (cell_smoothed f32, mask bool, seeds bool)."""
    cell = np.asarray(cell_prediction, dtype=np.float32)
    border = np.asarray(border_prediction, dtype=np.float32)
    cell = ndimage.gaussian_filter(cell, sigma=GAUSS_SIGMA)           # :25 (real scipy, same ndim as caller)
    border = np.clip(border, 0, 1)                                    # :27
    mask = cell > np.float32(th_cell)                                 # :30
    borders = tan_f32(border ** 2, tan_mode)                          # :33
    borders[borders < np.float32(0.05)] = 0                           # :34
    borders = np.clip(borders, 0, 1)                                  # :35
    cleaned = cell - borders                                          # :36
    seeds = cleaned > np.float32(th_seed)                             # :37
    return cell, mask, seeds


def filter_seeds(seeds_bool2d, floor_area=4.0, use_mean=True):
    """postprocessing.py:38-54 -> markers int32 (1..m, raster-first order)."""
    lab, n = label8(seeds_bool2d)                                     # :38
    if n > 0:
        areas = np.bincount(lab.ravel(), minlength=n + 1)[1:].astype(np.int64)   # :41-44
        min_area = 0.10 * np.mean(areas) if use_mean else 0.0         # :46
    else:
        areas = np.zeros(0, dtype=np.int64)
        min_area = 0.0                                                # :48
    min_area = np.maximum(min_area, floor_area)                       # :49
    keep = np.concatenate([[False], areas > min_area])                # :51-53 removes area <= min_area
    kept = keep[lab]
    lab2, m = label8(kept)                                            # :54
    return lab2, m


def distance_postprocessing(border_prediction, cell_prediction, th_seed, th_cell, tan_mode="f64",
                            marker_order=0, return_intermediates=False):
    """Restatement of postprocessing.py:7-59.  Same positional order as the reference."""
    cell, mask, seeds = seed_mask_maps(border_prediction, cell_prediction, th_seed, th_cell, tan_mode)
    cell2, mask2, seeds2 = _squeeze_hw(cell), _squeeze_hw(mask), _squeeze_hw(seeds)
    markers, m = filter_seeds(seeds2)
    inst = watershed(-cell2.astype(np.float64), markers, mask2, marker_order)      # :57
    out = np.squeeze(inst.astype(np.uint16))                                        # :59
    if return_intermediates:
        return out, dict(cell=cell2, mask=mask2, seeds=seeds2, markers=markers, n_markers=m, inst=inst)
    return out


def boundary_postprocessing(prediction, marker_order=0):
    """Restatement of postprocessing.py:62-90."""
    prediction = np.asarray(prediction)
    prediction_bin = np.argmax(prediction, axis=-1).astype(np.uint16)               # :71
    mask = prediction_bin == 1                                                      # :74
    seeds = (prediction[:, :, 1] * (1 - prediction[:, :, 2])) > 0.5                 # :77
    markers, m = filter_seeds(seeds, floor_area=4.0, use_mean=False)                # :78-85
    inst = watershed(mask.astype(np.float64), markers, mask, marker_order)          # :88
    return np.squeeze(inst.astype(np.uint16))
