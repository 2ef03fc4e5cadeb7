"""CPU tests of the augmentation oracle (oracle/augment.py) and of the host-side parameter logic (no GPU)."""
import random

import numpy as np
import scipy.ndimage

from microbeseg_b200 import augment as ga
from oracle import augment as oa


def test_flip_cases_are_the_eight_dihedral_maps():
    a = np.arange(49, dtype=np.uint16).reshape(7, 7, 1)
    outs = [oa.flip(a, h)[..., 0] for h in range(8)]
    assert len({o.tobytes() for o in outs}) == 8                       # all distinct
    assert np.array_equal(outs[6], a[..., 0].T)                        # flip lr + rot90 = transpose
    assert np.array_equal(outs[7], a[::-1, ::-1, 0].T)                 # flip ud + rot90 = anti-transpose
    for h in range(8):                                                 # the GPU path's inverse maps agree with NumPy
        m = ga.GpuAugmenter._dihedral(h, 7, 7)
        yy, xx = np.mgrid[0:7, 0:7]
        sx, sy = m[0] * xx + m[1] * yy + m[2], m[3] * xx + m[4] * yy + m[5]
        assert np.array_equal(a[sy, sx, 0], outs[h]), h


def test_affine_restatement_orientation_and_identity():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 60000, (9, 9, 1)).astype(np.uint16)
    assert np.array_equal(oa.warp(a, oa.affine_matrix(9, 9), 1), a)
    # +90 degrees about the centre = clockwise on the screen (skimage / imgaug convention) = np.rot90(k=-1)
    r = oa.warp(a, oa.affine_matrix(9, 9, rotate_deg=90.0), 0)
    assert np.array_equal(r[..., 0], np.rot90(a[..., 0], k=-1))
    f = rng.random((9, 9, 1)).astype(np.float32)
    assert np.allclose(oa.warp(f, oa.affine_matrix(9, 9, rotate_deg=90.0), 1)[..., 0], np.rot90(f[..., 0], k=-1), atol=1e-6)
    # scaling by 2 about the centre: the centre pixel stays, the border is filled from inside
    s = oa.warp(a, oa.affine_matrix(9, 9, scale_x=2.0, scale_y=2.0), 0)
    assert s[4, 4, 0] == a[4, 4, 0]
    # inverse used by the GPU path == inverse of the oracle's forward matrix
    inv = np.array(ga._affine_inverse(9, 9, 1.1, 0.9, 0.0)).reshape(2, 3)
    assert np.allclose(inv, np.linalg.inv(oa.affine_matrix(9, 9, 1.1, 0.9))[:2])


def test_gaussian_weights_match_scipy():
    for sigma in (1.0, 1.37, 1.99):
        w, r = oa.gaussian_weights(sigma)
        w2, r2 = ga._gaussian_weights(sigma)
        assert r == r2 == int(4 * sigma + 0.5) and np.array_equal(w, w2)
        imp = np.zeros(4 * r + 1)
        imp[2 * r] = 1.0
        assert np.allclose(scipy.ndimage.gaussian_filter1d(imp, sigma)[r:3 * r + 1], w, rtol=0, atol=1e-17)


def test_draw_params_call_order_and_rates():
    """the reference consumes random.random() once per transform (even for p = 1.0) and the parameters right after it"""
    r = random.Random(11)
    p = ga.draw_params(1, r, np.random.RandomState(5))[0]
    q = random.Random(11)                                # literal replay of mytransforms.py's __call__ sequence
    q.random(); flip = q.randint(0, 7)
    contrast = q.random() < 0.45
    assert p["flip"] == flip and (p["contrast"] != 0) == contrast
    ps = ga.draw_params(4000, random.Random(1), np.random.RandomState(1))
    frac = lambda f: np.mean([f(x) for x in ps])
    assert abs(frac(lambda x: x["contrast"] != 0) - 0.45) < 0.03 and abs(frac(lambda x: x["scale"] is not None) - 0.25) < 0.03
    assert abs(frac(lambda x: x["rotate"] is not None) - 0.25) < 0.03 and abs(frac(lambda x: x["blur_sigma"] is not None) - 0.3) < 0.03
    assert abs(frac(lambda x: x["noise"] > 0) - 0.3) < 0.03 and {x["flip"] for x in ps} == set(range(8))
    assert all(1.0 <= x["blur_sigma"] < 2.0 for x in ps if x["blur_sigma"] is not None)


def test_contrast_restatements():
    rng = np.random.default_rng(3)
    img = rng.integers(1000, 30000, (40, 40, 1)).astype(np.uint16)
    s = oa.contrast_stretch(img, 0.2, 99.8)
    assert s.dtype == np.uint16 and s.min() == 0 and s.max() == 65535
    g = oa.contrast_gamma(img, 1.0, 1.0)
    assert np.abs(g.astype(np.int64) - img.astype(np.int64)).max() <= 2          # identity parameters up to float32 rounding
    t = oa.to_tensor_image(img, 0, 65535)
    assert t.shape == (1, 40, 40) and t.dtype == np.float32 and -1 <= t.min() and t.max() <= 1


def test_against_the_reference_transform():
    """tests/golden/augment_reference.npz: outputs of the reference's OWN 'train' Compose (mytransforms.py:25-33: Flip, Contrast,
    Scaling, Rotate, Blur, Noise, ToTensor; imported by path in make_golden.py augment) on 24 seeds where neither an imgaug
    transform nor CLAHE is drawn.  draw_params with the same seeded generators + oracle/augment.py::apply must give the very same
    tensors: pins the RNG call order, the probabilities and the Flip / Contrast / Blur / ToTensor bodies to the reference's code."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment_reference.npz"))
    kinds = set()
    for k, seed in enumerate(g["seeds"]):
        random.seed(int(seed))
        np.random.seed(int(seed))
        p = ga.draw_params(1, py_random=random, np_random=np.random, clahe="error")[0]
        assert p["scale"] is None and p["rotate"] is None and p["noise"] == 0
        r = oa.apply({"image": g[f"image{k}"], "border_label": g[f"border{k}"], "cell_label": g[f"cell{k}"]}, p)
        assert np.array_equal(r["tensor"], g[f"t_image{k}"]), seed
        assert np.array_equal(np.transpose(r["border_label"], (2, 0, 1)), g[f"t_border{k}"]), seed
        assert np.array_equal(np.transpose(r["cell_label"], (2, 0, 1)), g[f"t_cell{k}"]), seed
        kinds.add((p["flip"], p["contrast"], p["blur_sigma"] is not None))
    assert {k[0] for k in kinds} == {0, 1, 2, 3, 4, 6, 7} and {k[1] for k in kinds} == {0, 1, 2} and {k[2] for k in kinds} == {False, True}
