"""Generates the committed golden fixtures under tests/golden/ (run in the build container).

  python tests/golden/make_golden.py postproc     # oracle/postproc.py on seeded synthetic maps
  python tests/golden/make_golden.py net          # the REAL reference DUNet (/root/reference) on CPU
  python tests/golden/make_golden.py labels       # oracle/labels.py on seeded synthetic instance masks
  python tests/golden/make_golden.py ranger       # the REAL reference Ranger optimizer (/root/reference) on CPU
  python tests/golden/make_golden.py losses       # the REAL reference ce_dice / CrossEntropyLoss (/root/reference) on CPU
  python tests/golden/make_golden.py simple_labels  # the REAL reference boundary_label / border_label / j4_label on CPU
  python tests/golden/make_golden.py labels_refbody # the REAL reference distance_label body over restated regionprops / label
  python tests/golden/make_golden.py postproc_refbody # the REAL reference post-processing bodies over restated label / regionprops / watershed
  python tests/golden/make_golden.py aji          # the REAL reference get_fast_aji_plus (stats_utils.py) on CPU
  python tests/golden/make_golden.py augment      # the REAL reference 'train' transform on seeds without imgaug / CLAHE draws
  python tests/golden/make_golden.py utils        # the REAL reference zero_pad_model_input / min_max_normalization

The post-processing goldens are produced by the oracle restatement (scikit-image cannot run in
this image -> "parity unpinned" for the skimage pieces, see oracle/postproc.py); the network
goldens are produced by importing the reference's own src/utils/unets.py, so they pin the
network oracle (oracle/net.py) and the CUDA path to the real reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def make_postproc():
    from microbeseg_b200 import synthetic as sy
    from oracle import postproc as op
    cases = [(96, 128, 40, 11), (128, 128, 70, 12), (160, 100, 45, 13), (64, 64, 0, 14)]
    for H, W, n_cells, seed in cases:
        m = sy.synth_instance_mask(H, W, n_cells, seed)
        border, cell = sy.synth_distance_maps(m, seed + 100)
        out, im = op.distance_postprocessing(border, cell, 0.45, 0.10, return_intermediates=True)
        np.savez_compressed(os.path.join(HERE, f"postproc_{H}x{W}_s{seed}.npz"), border=border, cell=cell,
                            th_seed=0.45, th_cell=0.10, mask_u16=out, cell_smooth=im["cell"],
                            n_markers=im["n_markers"])
        print("postproc", H, W, seed, "objects", int(out.max()))


def _real_skimage():
    """scikit-image itself, if this environment has it: the *_refbody generators then import the reference modules WITHOUT stubs, so
    the fixtures pin the restated primitives (measure.label, regionprops, watershed) as well.  Not the case in the build image."""
    try:
        import skimage                                      # noqa: F401
        from skimage import measure, segmentation           # noqa: F401
        return skimage.__version__
    except Exception:
        return None


def _plain_reference_import(name, path):
    import importlib.util
    sys.path.insert(0, "/root/reference")
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove("/root/reference")
    return mod


def _import_reference_postprocessing():
    """The reference's own src/inference/postprocessing.py, imported by path.  scikit-image is not installed here, so
    ``skimage.measure.label`` / ``regionprops`` / ``skimage.segmentation.watershed`` are the oracle's RESTATEMENTS
    (oracle/postproc.py::label8, bincount areas, oracle/watershed.c): the function bodies that run are the reference's own
    (thresholds, np.tan, area filter loop, relabelling, casts), only those three primitives remain restated."""
    import importlib.util
    import types
    from oracle import postproc as op
    if _real_skimage():
        print("   scikit-image", _real_skimage(), "found: importing the reference post-processing without stubs")
        return _plain_reference_import("reference_postprocessing", "/root/reference/src/inference/postprocessing.py")

    def _hw(a):
        a = np.asarray(a)
        return (a[..., 0], True) if a.ndim == 3 and a.shape[-1] == 1 else (a, False)

    def label(img, background=0, **kw):
        a, was3 = _hw(img)
        lab = op.label8(a != background)[0].astype(np.int64)
        return lab[..., None] if was3 else lab

    class _P:
        def __init__(self, lab_id, area):
            self.label, self.area = int(lab_id), int(area)

    def regionprops(lab):
        a, _ = _hw(lab)
        counts = np.bincount(a.ravel())
        return [_P(i, c) for i, c in enumerate(counts) if i > 0 and c > 0]

    def watershed(image, markers=None, mask=None, watershed_line=False, **kw):
        assert not watershed_line and not kw
        im, was3 = _hw(image)
        out = op.watershed(np.asarray(im, dtype=np.float64), _hw(markers)[0], _hw(mask)[0])
        return out[..., None] if was3 else out

    sk, sks, skme = (types.ModuleType(n) for n in ("skimage", "skimage.segmentation", "skimage.measure"))
    sks.watershed, skme.label, skme.regionprops = watershed, label, regionprops
    sk.segmentation, sk.measure = sks, skme
    names = ("skimage", "skimage.segmentation", "skimage.measure")
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update({"skimage": sk, "skimage.segmentation": sks, "skimage.measure": skme})
    try:
        spec = importlib.util.spec_from_file_location("reference_postprocessing", "/root/reference/src/inference/postprocessing.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def make_postproc_refbody():
    """distance_postprocessing / boundary_postprocessing: the reference's own function bodies (postprocessing.py:7-90) over
    the restated label / regionprops / watershed.  Pins everything in oracle/postproc.py except those three primitives (and
    shows whether the host's float32 np.tan, which the reference calls, agrees with the oracle's float32(tan(float64)) policy)."""
    from scipy import ndimage
    from microbeseg_b200 import synthetic as sy
    from oracle import postproc as op
    mod = _import_reference_postprocessing()
    for H, W, n_cells, seed in [(96, 128, 40, 61), (128, 112, 70, 62), (64, 64, 0, 63)]:
        m = sy.synth_instance_mask(H, W, n_cells, seed)
        border, cell = sy.synth_distance_maps(m, seed + 100)            # (H, W, 1) float32, as the frame loop passes them
        ref = mod.distance_postprocessing(border, cell, 0.45, 0.10)
        same = (np.array_equal(ref, op.distance_postprocessing(border, cell, 0.45, 0.10, tan_mode="host")),
                np.array_equal(ref, op.distance_postprocessing(border, cell, 0.45, 0.10)))
        inner = ndimage.binary_erosion(m > 0, iterations=2)
        rng = np.random.default_rng(seed)
        logits = rng.normal(0, 0.3, (H, W, 3)).astype(np.float32)
        logits[..., 0] += np.where(m == 0, 3.0, 0.0)
        logits[..., 1] += np.where(inner, 3.0, 0.0)
        logits[..., 2] += np.where((m > 0) & ~inner, 2.0, 0.0)
        e = np.exp(logits - logits.max(-1, keepdims=True))
        prob = (e / e.sum(-1, keepdims=True)).astype(np.float32)
        refb = mod.boundary_postprocessing(prob)
        sameb = np.array_equal(refb, op.boundary_postprocessing(prob))
        np.savez_compressed(os.path.join(HERE, f"refbody_postproc_{H}x{W}_s{seed}.npz"), border=border, cell=cell, th_seed=0.45,
                            th_cell=0.10, mask_u16=ref, prob=prob, boundary_mask_u16=refb)
        print("postproc refbody", H, W, seed, "objects", int(ref.max()), int(refb.max()),
              "oracle == reference body (host tan, f64 tan policy, boundary):", same, sameb)


def make_aji():
    """AJI+ values from the reference's OWN src/evaluation/stats_utils.py::get_fast_aji_plus (numpy + scipy's
    linear_sum_assignment; the module-level cv2 / matplotlib imports, which that function never touches, are empty stubs)."""
    import importlib.util
    import types
    from microbeseg_b200 import synthetic as sy
    names = ("cv2", "matplotlib", "matplotlib.pyplot")
    saved = {k: sys.modules.get(k) for k in names}
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.update({"cv2": types.ModuleType("cv2"), "matplotlib": mpl, "matplotlib.pyplot": plt})
    try:
        spec = importlib.util.spec_from_file_location("reference_stats_utils", "/root/reference/src/evaluation/stats_utils.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    cases = {}
    for k, (H, W, n, seed) in enumerate([(96, 128, 30, 71), (128, 128, 60, 72), (64, 80, 9, 73)]):
        t = sy.synth_instance_mask(H, W, n, seed).astype(np.int32)
        p = sy.synth_instance_mask(H, W, n, seed).astype(np.int32)
        rng = np.random.default_rng(seed)
        p = np.roll(p, (int(rng.integers(-3, 4)), int(rng.integers(-3, 4))), (0, 1))      # shifted prediction
        drop = rng.choice(np.arange(1, n + 1), size=max(1, n // 6), replace=False)          # missed cells
        p[np.isin(p, drop)] = 0
        if n > 4:
            p[p == 2] = 1                                                                   # a merge (ids stay contiguous below)
        ids = np.unique(p[p > 0])
        remap = np.zeros(int(p.max()) + 1, np.int32)
        remap[ids] = np.arange(1, len(ids) + 1)
        p = remap[p]                       # the metric expects contiguous ids (the reference feeds measure.label output)
        cases[f"true{k}"], cases[f"pred{k}"] = t, p
        cases[f"aji{k}"] = np.float64(mod.get_fast_aji_plus(t, p))
        cases[f"aji_swapped{k}"] = np.float64(mod.get_fast_aji_plus(p, t))
        print("aji+", H, W, seed, float(cases[f"aji{k}"]), float(cases[f"aji_swapped{k}"]))
    cases["n"] = np.int32(3)
    np.savez_compressed(os.path.join(HERE, "aji_plus_reference.npz"), **cases)


def make_augment():
    """The reference's OWN 'train' transform (mytransforms.py: Flip, Contrast, Scaling, Rotate, Blur, Noise, ToTensor in a
    Compose) for seeds on which neither an imgaug transform (Scaling / Rotate / Noise) nor CLAHE is drawn -- those need imgaug /
    scikit-image, which are not installed.  Stubs: ``torchvision.transforms.Compose`` (a loop over callables),
    ``skimage.exposure.rescale_intensity`` (restated: clip to in_range, scale to the dtype range -- parity unpinned),
    ``src.utils.utils`` is the reference's own module.  Pins the RNG call order, the decision thresholds and the bodies of Flip,
    Contrast (stretch / contrast + gamma), Blur and ToTensor of oracle/augment.py + microbeseg_b200.augment.draw_params."""
    import importlib.util
    import random
    import types
    import torch
    from microbeseg_b200 import synthetic as sy
    from microbeseg_b200.augment import draw_params

    def rescale_intensity(img, in_range):
        p0, p1 = in_range
        out = np.clip(img, p0, p1).astype(np.float64)
        if p0 != p1:
            out = (out - p0) / (p1 - p0)
        return (out * float(np.iinfo(img.dtype).max)).astype(img.dtype)

    def unusable(*a, **k):
        raise RuntimeError("imgaug / CLAHE are not available: this seed must be skipped")

    class Compose:
        def __init__(self, ts):
            self.ts = ts

        def __call__(self, x):
            for t in self.ts:
                x = t(x)
            return x

    names = ("imgaug", "imgaug.augmenters", "skimage", "skimage.exposure", "torchvision", "torchvision.transforms", "src", "src.utils",
             "src.utils.utils")
    mods = {n: types.ModuleType(n) for n in names[:6]}
    mods["imgaug"].augmenters = mods["imgaug.augmenters"]
    for attr in ("Sequential", "AdditiveGaussianNoise", "Affine"):
        setattr(mods["imgaug.augmenters"], attr, unusable)
    mods["skimage"].exposure = mods["skimage.exposure"]
    mods["skimage.exposure"].equalize_adapthist = unusable
    mods["skimage.exposure"].rescale_intensity = rescale_intensity
    mods["torchvision"].transforms = mods["torchvision.transforms"]
    mods["torchvision.transforms"].Compose = Compose
    spec_u = importlib.util.spec_from_file_location("src.utils.utils", "/root/reference/src/utils/utils.py")
    ref_utils = importlib.util.module_from_spec(spec_u)
    spec_u.loader.exec_module(ref_utils)
    src, srcu = types.ModuleType("src"), types.ModuleType("src.utils")
    src.utils, srcu.utils = srcu, ref_utils
    mods.update({"src": src, "src.utils": srcu, "src.utils.utils": ref_utils})
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update(mods)
    try:
        spec = importlib.util.spec_from_file_location("reference_mytransforms", "/root/reference/src/training/mytransforms.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    tf = mod.augmentors("distance", 0, 65535)["train"]
    H = W = 64
    out, kept, seen = {}, [], set()
    for seed in range(400):
        random.seed(seed)
        np.random.seed(seed)
        try:
            p = draw_params(1, py_random=random, np_random=np.random, clahe="error")[0]
        except NotImplementedError:
            continue
        if p["scale"] is not None or p["rotate"] is not None or p["noise"]:
            continue
        kind = (p["flip"], p["contrast"], p["percentiles"] if p["contrast"] == 1 else None, p["blur_sigma"] is not None)
        if kind in seen:
            continue
        seen.add(kind)
        m = sy.synth_instance_mask(H, W, 10, 900 + seed)
        img = sy.synth_frame(H, W, 900 + seed)[..., None].astype(np.uint16)
        rng = np.random.default_rng(seed)
        bl, cl = rng.random((H, W, 1)).astype(np.float32), rng.random((H, W, 1)).astype(np.float32)
        random.seed(seed)
        np.random.seed(seed)
        ti, tb, tc = tf({"image": img.copy(), "border_label": bl.copy(), "cell_label": cl.copy(), "id": "x"})
        k = len(kept)
        out[f"image{k}"], out[f"border{k}"], out[f"cell{k}"] = img, bl, cl
        out[f"t_image{k}"], out[f"t_border{k}"], out[f"t_cell{k}"] = ti.numpy(), tb.numpy(), tc.numpy()
        kept.append(seed)
        if len(kept) == 24:
            break
    out["seeds"] = np.array(kept, np.int64)
    np.savez_compressed(os.path.join(HERE, "augment_reference.npz"), **out)
    print("augment: seeds", kept)
    print("   kinds (flip, contrast, percentiles, blur):", sorted(seen, key=str))


def make_utils():
    """zero_pad_model_input / min_max_normalization from the reference's OWN src/utils/utils.py (json + numpy only)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_utils", "/root/reference/src/utils/utils.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    shapes = [(50, 70), (64, 64), (65, 8192), (1000, 1000), (2048, 2048), (2049, 100), (321, 4097), (8192, 8192), (1, 1), (6080, 6081),
              (33, 40, 3)]
    rng = np.random.default_rng(5)
    for k, shp in enumerate(shapes):
        small = tuple(min(d, 96) for d in shp[:2]) + tuple(shp[2:])      # content check on a small array of the same rank
        out[f"shape{k}"] = np.array(shp, np.int64)
        padded, pads = mod.zero_pad_model_input(np.zeros(shp, np.uint8), pad_val=3)
        out[f"pads{k}"], out[f"padded_shape{k}"] = np.array(pads, np.int64), np.array(padded.shape, np.int64)
        img = rng.integers(0, 200, small).astype(np.uint16)
        padded, pads = mod.zero_pad_model_input(img, pad_val=7)
        out[f"small{k}"], out[f"small_padded{k}"], out[f"small_pads{k}"] = img, padded, np.array(pads, np.int64)
    out["n_shapes"] = np.int64(len(shapes))
    imgs = [rng.integers(0, 65535, (9, 11)).astype(np.uint16), rng.integers(0, 255, (7, 5)).astype(np.uint8),
            rng.integers(100, 4000, (6, 6, 1)).astype(np.uint16)]
    k = 0
    for img in imgs:
        for lo, hi in [(None, None), (0, 65535), (200, 3000), (int(img.min()), int(img.max()))]:
            out[f"mm_in{k}"], out[f"mm_lo{k}"], out[f"mm_hi{k}"] = img, np.int64(-1 if lo is None else lo), np.int64(-1 if hi is None else hi)
            out[f"mm_out{k}"] = mod.min_max_normalization(img.copy(), min_value=lo, max_value=hi)
            k += 1
    out["n_mm"] = np.int64(k)
    try:
        mod.zero_pad_model_input(np.zeros((9000, 9000), np.uint8))
        out["too_big_raises"] = np.int64(0)
    except Exception:
        out["too_big_raises"] = np.int64(1)
    padded, pads = mod.zero_pad_model_input(np.zeros((9000, 100), np.uint8)) if False else (None, None)
    np.savez_compressed(os.path.join(HERE, "utils_reference.npz"), **out)
    print("utils: shapes", len(shapes), "normalisations", k, "too big raises", int(out["too_big_raises"]))


def make_net(only=None):
    import importlib.util
    import torch
    # the real reference module, loaded by path (this repo also has a drop-in `src` package)
    spec = importlib.util.spec_from_file_location("reference_unets", "/root/reference/src/utils/unets.py")
    reference_unets = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(reference_unets)
    build_unet = reference_unets.build_unet
    from oracle import net as onet
    torch.set_grad_enabled(False)
    for tag, filters, act, H, W, seed, pool, norm in [("f64-128_relu", (64, 128), "relu", 64, 64, 21, "conv", "bn"),
                                                      ("f64-256_mish", (64, 256), "mish", 48, 80, 22, "conv", "bn"),
                                                      ("f64-1024_relu", (64, 1024), "relu", 64, 64, 23, "conv", "bn"),
                                                      # the reference's low-memory fallback architecture (train.py:283-285)
                                                      ("f32-512_relu", (32, 512), "relu", 64, 96, 24, "conv", "bn"),
                                                      ("f64-256_elu_maxpool", (64, 256), "elu", 64, 48, 25, "max", "bn"),
                                                      ("f64-256_relu_gn", (64, 256), "relu", 48, 64, 26, "conv", "gn"),
                                                      ("f64-128_leakyrelu_in", (64, 128), "leakyrelu", 64, 48, 27, "conv", "in")]:
        if only and tag not in only:
            continue
        ref = build_unet("DU", act, pool, norm, torch.device("cpu"), 1, ch_in=1, ch_out=1, filters=list(filters))
        sd = onet.seeded_state_dict(ref.state_dict(), seed)
        ref.load_state_dict(sd)
        ref.eval()
        rng = np.random.default_rng(seed)
        img = rng.integers(100, 4000, (H, W)).astype(np.uint16)
        lo, hi = img.min(), img.max()
        x = 2 * (img.astype(np.float32) - lo) / (hi - lo) - 1
        border, cell = ref(torch.from_numpy(x[None, None]))
        np.savez_compressed(os.path.join(HERE, f"net_{tag}_{H}x{W}_s{seed}.npz"), img=img, filters=np.array(filters),
                            act=act, seed=seed, pool=pool, norm=norm, border=border[0, 0].numpy(), cell=cell[0, 0].numpy())
        print("net", tag, float(border.abs().max()), float(cell.abs().max()))


def make_labels():
    from microbeseg_b200 import synthetic as sy
    from oracle import labels as ol
    for H, W, n, seed in [(128, 128, 22, 31), (96, 160, 30, 32)]:
        m = sy.synth_instance_mask(H, W, n, seed, (9.0, 16.0), (7.0, 12.0)).astype(np.uint16)
        (cd, nd), mal = ol.create_labels(m)
        np.savez_compressed(os.path.join(HERE, f"labels_{H}x{W}_s{seed}.npz"), mask=m, cell_dist=cd, neighbor_dist=nd,
                            max_mal=mal)
        print("labels", H, W, seed, "max_mal", mal, "cells", int(m.max()))


def _import_reference_label_module(restated_measure=False):
    """The reference's own train_data_representations.py, imported by path, with the reference's own src/utils/utils.py
    (json + numpy only) behind ``from src.utils.utils import get_nucleus_ids``.  scikit-image and cv2 are not installed here:
    the module-level imports are satisfied with stubs.
      * ``skimage.morphology.disk`` is the published one-line definition ``x^2 + y^2 <= r^2`` on the (2r+1)^2 grid;
      * restated_measure=False: regionprops / label / cv2 stay unusable (None) -- boundary_label, border_label and j4_label never
        touch them (scipy + numpy otherwise);
      * restated_measure=True: ``skimage.measure.label`` / ``regionprops`` are the oracle's RESTATEMENTS (oracle/labels.py::label8,
        regionprops), so distance_label / bottom_hat_closing / cell_distance_label run the reference's own function bodies and
        only those two scikit-image primitives remain restated."""
    import importlib.util
    import types
    from oracle import labels as ol
    if _real_skimage():
        print("   scikit-image", _real_skimage(), "found: importing the reference label module without skimage stubs")
        try:
            import cv2                                       # noqa: F401
        except Exception:
            sys.modules["cv2"] = types.ModuleType("cv2")   # only adapted_border_label uses it
        return _plain_reference_import("reference_tdr", "/root/reference/src/training/train_data_representations.py")
    names = ("skimage", "skimage.morphology", "skimage.measure", "cv2", "src", "src.utils", "src.utils.utils")
    sk, skm, skme, cv2, src, srcu = (types.ModuleType(n) for n in names[:6])
    skm.disk = ol.disk
    sk.morphology, sk.measure = skm, skme
    skme.regionprops = skme.label = None
    if restated_measure:
        class _Props:
            def __init__(self, r):
                self._r = r
                self.label, self.area, self.centroid = r.label, r.area, r.centroid
                self.equivalent_diameter = float(np.sqrt(4.0 * r.area / np.pi))

            minor_axis_length = property(lambda self: self._r.minor_axis_length)
            major_axis_length = property(lambda self: self._r.major_axis_length)

        skme.regionprops = lambda lab: [_Props(r) for r in ol.regionprops(lab)]
        skme.label = lambda img, *a, **k: ol.label8(np.asarray(img) > 0)[0]
    spec_u = importlib.util.spec_from_file_location("src.utils.utils", "/root/reference/src/utils/utils.py")
    ref_utils = importlib.util.module_from_spec(spec_u)
    spec_u.loader.exec_module(ref_utils)
    src.utils, srcu.utils = srcu, ref_utils
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update({"skimage": sk, "skimage.morphology": skm, "skimage.measure": skme, "cv2": cv2, "src": src, "src.utils": srcu,
                        "src.utils.utils": ref_utils})
    try:
        spec = importlib.util.spec_from_file_location("reference_tdr", "/root/reference/src/training/train_data_representations.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def make_labels_refbody():
    """distance_label / cell_distance_label(clipped) / bottom_hat_closing: the reference's own function bodies
    (train_data_representations.py:40-72, 220-361) on top of the restated regionprops / label (see above).  The fixtures pin
    everything in oracle/labels.py except those two primitives."""
    from microbeseg_b200 import synthetic as sy
    from oracle import labels as ol
    mod = _import_reference_label_module(restated_measure=True)
    np.float = float          # the reference uses the removed NumPy alias np.float (train_data_representations.py:231, 276)
    for H, W, n, seed in [(112, 144, 26, 51), (96, 96, 40, 52)]:
        m = sy.synth_instance_mask(H, W, n, seed, (9.0, 16.0), (7.0, 12.0)).astype(np.uint16)
        mal = ol.max_major_axis_length(m)
        R = int(np.ceil(0.75 * mal))
        cd, nd = mod.distance_label(m, R)
        clipped = mod.cell_distance_label(m, R, apply_clipping=True)
        gaps, gap_map = mod.bottom_hat_closing(m)
        ocd, ond = ol.distance_label(m, R)
        same = (np.array_equal(cd, ocd), np.array_equal(nd, ond), np.array_equal(clipped, ol.get_label(m, "cell_dist_clipped", mal)))
        np.savez_compressed(os.path.join(HERE, f"refbody_labels_{H}x{W}_s{seed}.npz"), mask=m, max_mal=mal, cell_dist=cd,
                            neighbor_dist=nd, cell_dist_clipped=clipped, gaps=gaps.astype(np.int32), gap_map=gap_map)
        print("labels refbody", H, W, seed, "max_mal", mal, "gaps", int(gaps.max()), "oracle == reference body:", same)


def make_simple_labels():
    """boundary / border / j4 label images from the reference's OWN functions (train_data_representations.py:75-190)"""
    from microbeseg_b200 import synthetic as sy
    mod = _import_reference_label_module()
    for H, W, n, seed in [(72, 96, 18, 41), (64, 64, 30, 42)]:
        m = sy.synth_instance_mask(H, W, n, seed, (7.0, 12.0), (5.0, 9.0)).astype(np.uint16)
        m[0:5, 0:6] = 900                  # touches the image corner
        m[0:5, 6:11] = 901                 # and a neighbour
        out = {"mask": m, "boundary": mod.boundary_label(m), "border": mod.border_label(m), "j4": mod.j4_label(m)}
        np.savez_compressed(os.path.join(HERE, f"simple_labels_{H}x{W}_s{seed}.npz"), **out)
        print("simple labels", H, W, seed, {k: np.bincount(v.ravel()).tolist() for k, v in out.items() if k != "mask"})


RANGER_SHAPES = [(8, 4, 3, 3), (8,), (6, 8, 2, 2), (5, 7), (3,), (2, 1100)]
RANGER_CASES = {"default": dict(lr=0.05), "wd_convonly": dict(lr=0.02, weight_decay=0.01, gc_conv_only=True, k=4),
                "nogc": dict(lr=0.05, use_gc=False, betas=(0.9, 0.99))}
RANGER_STEPS = 14
RANGER_SNAPSHOTS = (1, 5, 6, 7, 12, 14)


def ranger_inputs(case_idx):
    """seeded initial parameters and per-step gradients (regenerated identically by the tests)"""
    rng = np.random.default_rng(7000 + case_idx)
    params = [rng.standard_normal(s).astype(np.float32) for s in RANGER_SHAPES]
    grads = [[(rng.standard_normal(s) * 0.3 + 0.05).astype(np.float32) for s in RANGER_SHAPES] for _ in range(RANGER_STEPS)]
    return params, grads


def make_ranger():
    import contextlib
    import importlib.util
    import io
    import torch
    spec = importlib.util.spec_from_file_location("reference_ranger", "/root/reference/src/training/ranger2020.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for ci, (name, kw) in enumerate(RANGER_CASES.items()):
        p0, grads = ranger_inputs(ci)
        params = [torch.nn.Parameter(torch.from_numpy(a.copy())) for a in p0]
        with contextlib.redirect_stdout(io.StringIO()):
            opt = mod.Ranger(params, **kw)
        out = {}
        for step in range(1, RANGER_STEPS + 1):
            for p, g in zip(params, grads[step - 1]):
                p.grad = torch.from_numpy(g.copy())
            opt.step()
            if step in RANGER_SNAPSHOTS:
                for i, p in enumerate(params):
                    st = opt.state[p]
                    out[f"s{step}_p{i}"] = p.detach().numpy().copy()
                    out[f"s{step}_m{i}"] = st["exp_avg"].numpy().copy()
                    out[f"s{step}_v{i}"] = st["exp_avg_sq"].numpy().copy()
                    out[f"s{step}_slow{i}"] = st["slow_buffer"].numpy().copy()
                    out[f"s{step}_g{i}"] = p.grad.numpy().copy()      # the reference centralises p.grad in place
        np.savez_compressed(os.path.join(HERE, f"ranger_{name}.npz"), **out)
        print("ranger", name, "final |p0|", float(np.abs(out[f"s{RANGER_STEPS}_p0"]).mean()))


def make_losses():
    """loss value and dloss/dlogits of the boundary-method criteria from the reference's own losses.py (torch autograd, fp32)"""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("reference_losses", "/root/reference/src/training/losses.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for seed, (n, h, w) in enumerate([(2, 24, 40), (3, 33, 17)]):
        rng = np.random.default_rng(8100 + seed)
        logits = (rng.standard_normal((n, 3, h, w)) * 2.0).astype(np.float32)
        labels = rng.integers(0, 3, (n, h, w)).astype(np.int64)
        labels[0, :4] = 0                                         # a stretch of pure background
        out = {"logits": logits, "labels": labels}
        for kind in ("ce_dice", "ce"):
            crit = mod.get_loss(kind, "boundary")
            z = torch.from_numpy(logits).clone().requires_grad_(True)
            loss = crit(z, torch.from_numpy(labels))
            loss.backward()
            out[kind + "_loss"] = np.float32(loss.item())
            out[kind + "_grad"] = z.grad.numpy().copy()
        np.savez_compressed(os.path.join(HERE, f"ce_dice_s{seed}.npz"), **out)
        print("losses", seed, float(out["ce_dice_loss"]), float(out["ce_loss"]))


if __name__ == "__main__":
    what = sys.argv[1:] or ["postproc", "net", "labels", "ranger", "losses", "simple_labels", "labels_refbody", "postproc_refbody", "aji", "augment", "utils"]
    if "simple_labels" in what:
        make_simple_labels()
    if "labels_refbody" in what:
        make_labels_refbody()
    if "postproc_refbody" in what:
        make_postproc_refbody()
    if "aji" in what:
        make_aji()
    if "augment" in what:
        make_augment()
    if "utils" in what:
        make_utils()
    if "losses" in what:
        make_losses()
    if "ranger" in what:
        make_ranger()
    if "postproc" in what:
        make_postproc()
    if "net" in what:
        make_net([a[4:] for a in what if a.startswith("net:")] or None)
    if "labels" in what:
        make_labels()
