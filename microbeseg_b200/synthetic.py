"""Synthetic microscopy-like inputs for tests and bench (SURVEY.md section 8(d) generators).

No dataset or checkpoint is reachable offline, so every config in BASELINE.json is exercised on
seeded synthetic data: instance masks of roundish cells, distance maps derived from them
(config 3), and uint16 frames rendered from them (configs 1, 2).
"""
import numpy as np
from scipy import ndimage


def synth_instance_mask(H, W, n_cells, seed, a_rng=(7.0, 14.0), b_rng=(5.0, 10.0)):
    """uint16/int32 instance mask with ~n_cells ellipses on a jittered grid (touching allowed)."""
    rng = np.random.default_rng(seed)
    lab = np.zeros((H, W), dtype=np.int32)
    pitch = np.sqrt(H * W / max(n_cells, 1))
    gy, gx = max(int(round(H / pitch)), 1), max(int(round(W / pitch)), 1)
    cy = (np.arange(gy) + 0.5) * H / gy
    cx = (np.arange(gx) + 0.5) * W / gx
    centers = np.stack(np.meshgrid(cy, cx, indexing="ij"), -1).reshape(-1, 2)
    centers = centers + rng.uniform(-0.35, 0.35, centers.shape) * pitch
    rng.shuffle(centers)
    centers = centers[:n_cells]
    a = rng.uniform(*a_rng, len(centers))
    b = rng.uniform(*b_rng, len(centers))
    th = rng.uniform(0, np.pi, len(centers))
    R = int(np.ceil(max(a_rng[1], b_rng[1]))) + 1
    yy, xx = np.mgrid[-R:R + 1, -R:R + 1].astype(np.float64)
    k = 0
    for (y0, x0), ai, bi, ti in zip(centers, a, b, th):
        iy, ix = int(round(y0)), int(round(x0))
        ys, ye = max(iy - R, 0), min(iy + R + 1, H)
        xs, xe = max(ix - R, 0), min(ix + R + 1, W)
        if ys >= ye or xs >= xe:
            continue
        Y = yy[ys - iy + R:ye - iy + R, xs - ix + R:xe - ix + R]
        X = xx[ys - iy + R:ye - iy + R, xs - ix + R:xe - ix + R]
        c, s = np.cos(ti), np.sin(ti)
        u = (X * c + Y * s) / ai
        v = (-X * s + Y * c) / bi
        inside = (u * u + v * v) <= 1.0
        win = lab[ys:ye, xs:xe]
        put = inside & (win == 0)
        if put.sum() < 12:
            continue
        k += 1
        win[put] = k
    return lab


def synth_distance_maps(mask, seed, noise=0.02):
    """(border, cell) float32 maps of shape (H,W,1) as the network would predict for ``mask``.

    cell  = per-instance Euclidean distance / per-instance max (+ noise, breaks ties);
    border = 0.9 inside cells within 2 px of a *different* cell (+ noise).
    """
    rng = np.random.default_rng(seed)
    lab = np.asarray(mask).astype(np.int32)
    H, W = lab.shape
    P = np.pad(lab, 1, mode="edge")
    same = ((P[:-2, 1:-1] == lab) & (P[2:, 1:-1] == lab) & (P[1:-1, :-2] == lab) & (P[1:-1, 2:] == lab))
    interior = (lab > 0) & same
    d = ndimage.distance_transform_edt(interior)
    n = int(lab.max())
    cell = np.zeros((H, W), dtype=np.float64)
    if n > 0:
        mx = ndimage.maximum(d, lab, index=np.arange(1, n + 1))
        mx = np.concatenate([[1.0], np.maximum(mx, 1e-6)])
        cell = d / mx[lab]
    # neighbour zones: pixels of a cell within 2 px (chebyshev) of another cell
    big = np.int32(np.iinfo(np.int32).max)
    lo = np.where(lab > 0, lab, big)
    mn = ndimage.minimum_filter(lo, size=5, mode="constant", cval=big)
    mxl = ndimage.maximum_filter(lab, size=5, mode="constant", cval=0)
    touch = (lab > 0) & ((mn != lab) | (mxl != lab))
    border = 0.9 * touch.astype(np.float64)
    border = ndimage.gaussian_filter(border, 0.7)
    cell = cell + rng.normal(0, noise, (H, W))
    border = border + rng.normal(0, noise, (H, W))
    return (np.ascontiguousarray(border[..., None], dtype=np.float32),
            np.ascontiguousarray(cell[..., None], dtype=np.float32))


def synth_frame(H, W, seed, n_cells=None):
    """One uint16 frame: noisy background + blurred bright ellipses."""
    rng = np.random.default_rng(seed)
    if n_cells is None:
        n_cells = max(int(H * W * 0.4 / 330.0), 1)
    lab = synth_instance_mask(H, W, n_cells, seed + 1)
    amp = rng.uniform(1500, 4000, int(lab.max()) + 1)
    amp[0] = 0
    img = ndimage.gaussian_filter(amp[lab], 1.2)
    img = img + rng.normal(2000, 60, (H, W))
    img = img + rng.normal(0, 1, (H, W)) * np.sqrt(np.maximum(img, 0)) * 0.5
    return np.clip(img, 0, 65535).astype(np.uint16)


def synth_stack(T, H, W, seed0=2000, distinct=4):
    """[T,H,W] uint16 stack; ``distinct`` rendered frames are cycled (with a per-frame intensity
    offset so that min/max differ) to keep generation time bounded for T=200 x 2048^2."""
    base = [synth_frame(H, W, seed0 + t) for t in range(min(distinct, T))]
    out = np.empty((T, H, W), dtype=np.uint16)
    for t in range(T):
        f = base[t % len(base)]
        off = (t // len(base)) * 3
        out[t] = np.clip(f.astype(np.int32) + off, 0, 65535).astype(np.uint16)
    return out
