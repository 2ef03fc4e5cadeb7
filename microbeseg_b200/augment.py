"""Training-time augmentations on the GPU, batched over crops.

Drop-in for the 'train' transform of ``augmentors(label_type, min_value, max_value)`` (src/training/mytransforms.py:13-35:
Flip p=1.0, Contrast p=0.45, Scaling p=0.25, Rotate p=0.25, Blur p=0.3, Noise p=0.3, ToTensor) for the distance method.
``draw_params`` consumes Python's ``random`` / ``numpy.random`` in exactly the reference's call order, so a seeded run
makes the same decisions with the same parameters; ``GpuAugmenter`` applies them to a whole batch that already lives on
the device (no DataLoader workers, no host round trip).  Not built: the CLAHE branch of Contrast (skimage's
equalize_adapthist) -- ``clahe='stretch'`` substitutes the 0.2 / 99.8 percentile stretch, ``clahe='error'`` raises.
"""
import math
import random as _py_random

import numpy as np
import torch

from . import _native as nat

# inverse maps (sx, sy) = M (x, y, 1) of the eight dihedral transforms of Flip (mytransforms.py:150-231), n = side - 1
_DIHEDRAL = {
    0: lambda n: (1, 0, 0, 0, 1, 0),
    1: lambda n: (-1, 0, n, 0, 1, 0),           # flip left-right
    2: lambda n: (1, 0, 0, 0, -1, n),           # flip up-down
    3: lambda n: (0, -1, n, 1, 0, 0),           # rot90:  out[y][x] = in[x][n - y]
    4: lambda n: (-1, 0, n, 0, -1, n),          # rot180
    5: lambda n: (0, 1, 0, -1, 0, n),           # rot270: out[y][x] = in[n - x][y]
    6: lambda n: (0, 1, 0, 1, 0, 0),            # flip lr + rot90 = transpose
    7: lambda n: (0, -1, n, -1, 0, n),          # flip ud + rot90 = anti-transpose
}


def draw_params(n, py_random=_py_random, np_random=np.random, clahe="stretch"):
    """One parameter dict per sample, drawn in the reference's call order (Flip, Contrast, Scaling, Rotate, Blur, Noise)."""
    out = []
    for _ in range(n):
        p = {"flip": 0, "contrast": 0, "percentiles": (0.2, 99.8), "factor": 1.0, "gamma": 1.0, "scale": None, "rotate": None,
             "blur_sigma": None, "noise": 0}
        if py_random.random() < 1.0:                      # Flip(p=1.0)
            p["flip"] = py_random.randint(0, 7)
        if py_random.random() < 0.45:                     # Contrast(p=0.45)
            h = py_random.randint(0, 2)
            if h == 0:
                if clahe == "error":
                    raise NotImplementedError("Contrast: the CLAHE branch (skimage equalize_adapthist) is not built on the GPU path")
                p["contrast"], p["percentiles"] = 1, (0.2, 99.8)
            elif h == 1:
                p["contrast"] = 1
                p["percentiles"] = (0.2, 99.8) if py_random.randint(0, 1) == 0 else (0.1, 99.9)
            else:
                p["contrast"] = 2
                p["factor"] = float(np_random.uniform(0.75, 1.25))
                p["gamma"] = float(np_random.uniform(0.7, 1.3))
        if py_random.random() < 0.25:                     # Scaling(p=0.25)
            p["scale"] = (py_random.uniform(0.85, 1.15), py_random.uniform(0.85, 1.15))
        if py_random.random() < 0.25:                     # Rotate(p=0.25)
            p["rotate"] = py_random.uniform(-45, 45)
        if py_random.random() < 0.3:                      # Blur(p=0.3)
            p["blur_sigma"] = py_random.random() + 1.0
        if py_random.random() < 0.3:                      # Noise(p=0.3)
            p["noise"] = py_random.randint(1, 5)
        out.append(p)
    return out


def _affine_inverse(h, w, scale_x=1.0, scale_y=1.0, rotate_deg=0.0):
    """inverse of imgaug's Affine matrix (to_topleft + AffineTransform(scale, rotation) + to_center), centre (w/2-.5, h/2-.5)"""
    sx, sy = w / 2.0 - 0.5, h / 2.0 - 0.5
    t = math.radians(rotate_deg)
    a = np.array([[scale_x * math.cos(t), -scale_y * math.sin(t), 0.0], [scale_x * math.sin(t), scale_y * math.cos(t), 0.0], [0, 0, 1.0]])
    t0 = np.array([[1, 0, -sx], [0, 1, -sy], [0, 0, 1.0]])
    t1 = np.array([[1, 0, sx], [0, 1, sy], [0, 0, 1.0]])
    inv = np.linalg.inv(t1 @ a @ t0)
    return tuple(inv[0]) + tuple(inv[1])


def _gaussian_weights(sigma):
    """scipy's _gaussian_kernel1d (float64, NumPy arithmetic): the kernels only multiply and add"""
    radius = int(4.0 * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    return phi / phi.sum(), radius


class GpuAugmenter:
    """``aug(images, border_labels, cell_labels, params)`` -> (img float32 [n,1,H,W], border [n,1,H,W], cell [n,1,H,W]).
    images: uint16 [n,H,W] CUDA tensor (square crops: Flip rotates by 90 degrees), labels float32 [n,H,W]."""

    def __init__(self, min_value=0, max_value=65535, seed=0):
        self.min_value, self.max_value = float(min_value), float(max_value)
        self.seed = int(seed)
        self.calls = 0
        self.L = nat.lib()

    def _dev(self, arr, dtype, device):
        return torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(device)

    def _warp(self, tensors, mats, modes, device):
        n, H, W = tensors[0].shape
        m = self._dev(mats, np.float64, device)
        md = self._dev(modes, np.int32, device)
        out = []
        for t in tensors:
            code = {torch.int16: 1, torch.uint16: 1, torch.float32: 2, torch.uint8: 0}[t.dtype]
            dst = torch.empty_like(t)
            nat.check(self.L.mbs_aug_warp(t.data_ptr(), dst.data_ptr(), code, n, H, W, m.data_ptr(), md.data_ptr(), nat.stream_ptr()),
                      "aug_warp")
            out.append(dst)
        return out

    @torch.no_grad()
    def __call__(self, images, border_labels, cell_labels, params, return_image=False):
        if not images.is_cuda:
            raise RuntimeError("microbeseg_b200.augment needs CUDA tensors (no CPU fallback)")
        if images.dtype not in (torch.uint16, torch.int16):
            raise RuntimeError("augment: images must be uint16 (pass a uint16 tensor or its int16 view)")
        n, H, W = images.shape
        if len(params) != n:
            raise ValueError("one parameter dict per sample")
        dev = images.device
        img = images.contiguous()
        bl, cl = border_labels.contiguous().float(), cell_labels.contiguous().float()
        with torch.cuda.device(dev):
            ws = torch.empty(int(self.L.mbs_aug_workspace_bytes(n)), dtype=torch.uint8, device=dev)
            # Flip: always applied (p = 1.0); h = 0 is the identity
            if any(p["flip"] for p in params):
                if H != W and any(p["flip"] >= 3 for p in params):
                    raise RuntimeError("Flip rotates by 90 degrees: crops must be square (mytransforms.py:129)")
                mats = [self._dihedral(p["flip"], H, W) for p in params]
                modes = [1 if p["flip"] else 0 for p in params]
                img, bl, cl = self._warp([img, bl, cl], mats, modes, dev)
            # Contrast
            if any(p["contrast"] for p in params):
                img = img.clone() if img.data_ptr() == images.data_ptr() else img
                modes = self._dev([p["contrast"] for p in params], np.int32, dev)
                pr = self._dev([[p["percentiles"][0], p["percentiles"][1], p["factor"], p["gamma"]] for p in params], np.float64, dev)
                nat.check(self.L.mbs_aug_contrast(img.data_ptr(), n, H, W, modes.data_ptr(), pr.data_ptr(), ws.data_ptr(), ws.numel(),
                                                  nat.stream_ptr()), "aug_contrast")
            # Scaling, then Rotate: two resamplings, as in the reference
            for key in ("scale", "rotate"):
                if any(p[key] is not None for p in params):
                    mats = [(1, 0, 0, 0, 1, 0) if p[key] is None else
                            (_affine_inverse(H, W, scale_x=p[key][0], scale_y=p[key][1]) if key == "scale" else
                             _affine_inverse(H, W, rotate_deg=p[key])) for p in params]
                    modes = [0 if p[key] is None else 2 for p in params]
                    img, bl, cl = self._warp([img, bl, cl], mats, modes, dev)
            # Blur
            if any(p["blur_sigma"] is not None for p in params):
                img = img.clone() if img.data_ptr() == images.data_ptr() else img
                wts = np.zeros((n, 17), np.float64)
                rad = np.zeros(n, np.int32)
                for i, p in enumerate(params):
                    if p["blur_sigma"] is not None:
                        w, r = _gaussian_weights(p["blur_sigma"])
                        wts[i, :2 * r + 1], rad[i] = w, r
                wd, rd = self._dev(wts, np.float64, dev), self._dev(rad, np.int32, dev)
                tmp = torch.empty_like(img)
                nat.check(self.L.mbs_aug_blur(img.data_ptr(), tmp.data_ptr(), n, H, W, wd.data_ptr(), rd.data_ptr(), nat.stream_ptr()), "aug_blur")
            # Noise + ToTensor
            frac = self._dev([p["noise"] / 100.0 for p in params], np.float32, dev)
            out = torch.empty((n, 1, H, W), dtype=torch.float32, device=dev)
            img_out = torch.empty_like(img) if return_image else None
            self.calls += 1
            seed = (self.seed * 0x9E3779B97F4A7C15 + self.calls * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
            nat.check(self.L.mbs_aug_noise_normalize(img.data_ptr(), n, H, W, frac.data_ptr(), seed, self.min_value, self.max_value,
                                                     img_out.data_ptr() if return_image else None, out.data_ptr(), ws.data_ptr(),
                                                     ws.numel(), nat.stream_ptr()), "aug_noise_normalize")
        res = (out, bl[:, None], cl[:, None])
        return res + (img_out,) if return_image else res

    @staticmethod
    def _dihedral(h, H, W):
        """inverse map of Flip's case h for an H x W image (cases >= 3 need H == W)"""
        if h == 0:
            return (1, 0, 0, 0, 1, 0)
        if h == 1:
            return (-1, 0, W - 1, 0, 1, 0)
        if h == 2:
            return (1, 0, 0, 0, -1, H - 1)
        if h == 4:
            return (-1, 0, W - 1, 0, -1, H - 1)
        return _DIHEDRAL[h](H - 1)
