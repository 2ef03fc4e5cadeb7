"""CPU restatement of the reference's training augmentations (TEST INFRASTRUCTURE: only tests/ may import this).

Follows src/training/mytransforms.py:13-406 on one sample = dict(image uint16 (H,W,1), border_label / cell_label float32
(H,W,1)) for explicit parameters (the decisions are drawn by microbeseg_b200.augment.draw_params in the reference's call
order).  Pinned pieces: Flip (NumPy itself), Blur (scipy.ndimage.gaussian_filter itself), ToTensor (NumPy).  PARITY
UNPINNED: Scaling / Rotate (imgaug 0.4 Affine -> cv2.warpAffine: restated as an inverse-mapped bilinear / nearest resampling
about the image centre (w/2 - 0.5, h/2 - 0.5) with constant border 0, float arithmetic instead of cv2's 1/32 fixed-point
coefficients), the percentile stretch (skimage 0.19 rescale_intensity restated), the imgaug noise model (round(N(0, s)),
clip).  The CLAHE branch (skimage.exposure.equalize_adapthist) is not restated."""
import numpy as np
import scipy.ndimage


def flip(arr, h):
    """mytransforms.py:150-231"""
    if h == 0:
        return arr
    if h == 1:
        return np.flip(arr, axis=1).copy()
    if h == 2:
        return np.flip(arr, axis=0).copy()
    if h in (3, 4, 5):
        return np.rot90(arr, k=h - 2, axes=(0, 1)).copy()
    if h == 6:
        return np.rot90(np.flip(arr, axis=1).copy(), axes=(0, 1)).copy()
    return np.rot90(np.flip(arr, axis=0).copy(), k=1, axes=(0, 1)).copy()


def affine_matrix(h, w, scale_x=1.0, scale_y=1.0, rotate_deg=0.0):
    """forward matrix of imgaug Affine: to_topleft + AffineTransform(scale, rotation) + to_center"""
    sx, sy = w / 2.0 - 0.5, h / 2.0 - 0.5
    t = np.deg2rad(rotate_deg)
    a = np.array([[scale_x * np.cos(t), -scale_y * np.sin(t), 0.0], [scale_x * np.sin(t), scale_y * np.cos(t), 0.0], [0, 0, 1.0]])
    t0 = np.array([[1, 0, -sx], [0, 1, -sy], [0, 0, 1.0]])
    t1 = np.array([[1, 0, sx], [0, 1, sy], [0, 0, 1.0]])
    return t1 @ a @ t0


def warp(arr, forward, order):
    """cv2.warpAffine(arr, M, dsize=(w, h), flags=INTER_LINEAR / INTER_NEAREST, borderMode=BORDER_CONSTANT, borderValue=0)"""
    a = arr[..., 0].astype(np.float64)
    h, w = a.shape
    inv = np.linalg.inv(forward)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    sx = inv[0, 0] * xx + inv[0, 1] * yy + inv[0, 2]
    sy = inv[1, 0] * xx + inv[1, 1] * yy + inv[1, 2]

    def at(iy, ix):
        ok = (iy >= 0) & (iy < h) & (ix >= 0) & (ix < w)
        return np.where(ok, a[np.clip(iy, 0, h - 1), np.clip(ix, 0, w - 1)], 0.0)

    if order == 0:
        out = at(np.floor(sy + 0.5).astype(np.int64), np.floor(sx + 0.5).astype(np.int64))
    else:
        fx, fy = np.floor(sx), np.floor(sy)
        ix, iy = fx.astype(np.int64), fy.astype(np.int64)
        ax, ay = sx - fx, sy - fy
        out = (1 - ay) * ((1 - ax) * at(iy, ix) + ax * at(iy, ix + 1)) + ay * ((1 - ax) * at(iy + 1, ix) + ax * at(iy + 1, ix + 1))
    if arr.dtype.kind == "u":
        out = np.clip(np.rint(out), 0, np.iinfo(arr.dtype).max)
    return out.astype(arr.dtype)[..., None]


def contrast_stretch(img, q0, q1):
    """mytransforms.py:97-103: np.percentile + skimage.exposure.rescale_intensity(in_range=(p0, p1)) (out_range 'dtype')"""
    p0, p1 = np.percentile(img, (q0, q1))
    out = np.clip(img, p0, p1).astype(np.float64)
    if p0 != p1:
        out = (out - p0) / (p1 - p0)
    return (out * 65535.0 + 0.0).astype(img.dtype)


def contrast_gamma(img, factor, gamma):
    """mytransforms.py:105-124 (float32 arithmetic as NumPy evaluates it)"""
    dtype = img.dtype
    f32 = np.float32
    x = (img.astype(f32) - f32(np.iinfo(dtype).min)) / f32(np.iinfo(dtype).max - np.iinfo(dtype).min)
    img_mean = x.mean()
    x = (x - img_mean) * f32(factor) + img_mean
    img_min, img_max = x.min(), x.max()
    rnge = img_max - img_min
    x = np.power(((x - img_min) / f32(rnge + f32(1e-7))), f32(gamma)) * rnge + img_min
    x = np.clip(x, 0, 1)
    x = x * f32(np.iinfo(dtype).max - np.iinfo(dtype).min) - f32(np.iinfo(dtype).min)
    return x.astype(dtype)


def blur(img, sigma):
    """mytransforms.py:57-59"""
    return scipy.ndimage.gaussian_filter(img, sigma, order=0)


def gaussian_weights(sigma, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius), radius = int(truncate * sigma + 0.5)"""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    return phi / phi.sum(), radius


def to_tensor_image(img, min_value, max_value):
    """ToTensor + min_max_normalization (mytransforms.py:395, utils.py:50-74) -> float32 (1,H,W)"""
    x = np.clip(img, min_value, max_value)
    x = 2 * (x.astype(np.float32) - min_value) / (max_value - min_value) - 1
    return np.transpose(x.astype(np.float32), (2, 0, 1))


def apply(sample, p, min_value=0, max_value=65535):
    """The 'train' Compose of augmentors() (mytransforms.py:25-33) for explicit parameters ``p`` (one entry of
    microbeseg_b200.augment.draw_params), noise excluded (its random field belongs to the generator)."""
    img, bl, cl = sample["image"], sample["border_label"], sample["cell_label"]
    img, bl, cl = flip(img, p["flip"]), flip(bl, p["flip"]), flip(cl, p["flip"])
    if p["contrast"] == 1:
        img = contrast_stretch(img, *p["percentiles"])
    elif p["contrast"] == 2:
        img = contrast_gamma(img, p["factor"], p["gamma"])
    h, w = img.shape[:2]
    if p["scale"] is not None:
        f = affine_matrix(h, w, scale_x=p["scale"][0], scale_y=p["scale"][1])
        img, bl, cl = warp(img, f, 1), warp(bl, f, 1), warp(cl, f, 1)
    if p["rotate"] is not None:
        f = affine_matrix(h, w, rotate_deg=p["rotate"])
        img, bl, cl = warp(img, f, 1), warp(bl, f, 1), warp(cl, f, 1)
    if p["blur_sigma"] is not None:
        img = blur(img, p["blur_sigma"])
    return {"image": img, "border_label": bl, "cell_label": cl,
            "tensor": to_tensor_image(img, min_value, max_value)}
