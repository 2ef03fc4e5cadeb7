"""CPU tests: the post-processing oracle against scipy, its golden fixtures and an independent
formulation; and the C-ABI library surface (loads, exports every declared symbol)."""
import glob
import os
import re

import numpy as np
import pytest
from scipy import ndimage

from oracle import postproc as op
from microbeseg_b200 import synthetic as sy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_gaussian_weights_match_scipy():
    imp = np.zeros(5, dtype=np.float64)
    imp[2] = 1.0
    w = ndimage.gaussian_filter1d(imp, 0.5, mode="constant")
    assert np.array_equal(w, op.gaussian_weights())
    assert [float(x).hex() for x in op.gaussian_weights()[2:]] == [
        "0x1.92b965ef5aaeep-1", "0x1.b405b9842b206p-4", "0x1.14aebe6a24088p-12"]


@pytest.mark.parametrize("shape", [(1, 1), (2, 5), (3, 3), (17, 40), (64, 33)])
def test_gaussian_emulation_bit_exact_vs_scipy(shape):
    rng = np.random.default_rng(sum(shape))
    a = rng.normal(size=shape).astype(np.float32)
    ref = ndimage.gaussian_filter(a[..., None], 0.5)[..., 0]   # (H,W,1) as the reference passes it
    assert np.array_equal(ref, op.gaussian_smooth_emulated(a))
    assert np.array_equal(ref, ndimage.gaussian_filter(a, 0.5))


def test_label8_is_raster_first_order():
    b = np.zeros((6, 7), bool)
    b[0, 5] = b[1, 0] = b[1, 1] = b[2, 2] = b[4, 4] = True      # (1,1)-(2,2) touch diagonally
    lab, n = op.label8(b)
    assert n == 3 and lab[0, 5] == 1 and lab[1, 0] == 2 and lab[2, 2] == 2 and lab[4, 4] == 3


def test_golden_postproc():
    files = sorted(glob.glob(os.path.join(HERE, "golden", "postproc_*.npz")))
    assert len(files) >= 4
    for f in files:
        g = np.load(f)
        out, im = op.distance_postprocessing(g["border"], g["cell"], float(g["th_seed"]), float(g["th_cell"]),
                                             return_intermediates=True)
        assert out.dtype == np.uint16 and out.shape == g["cell"].shape[:2]
        assert np.array_equal(out, g["mask_u16"]), f
        assert np.array_equal(im["cell"], g["cell_smooth"]), f
        assert im["n_markers"] == int(g["n_markers"])


def test_against_the_reference_function_bodies():
    """tests/golden/refbody_postproc_*.npz: outputs of the reference's OWN distance_postprocessing / boundary_postprocessing
    (postprocessing.py:7-90, imported by path in make_golden.py postproc_refbody) running on top of the restated
    measure.label / regionprops / watershed -- everything in oracle/postproc.py except those three scikit-image primitives is
    pinned to the reference's code (the reference calls the host's float32 np.tan: both tan policies give these masks)"""
    files = sorted(glob.glob(os.path.join(HERE, "golden", "refbody_postproc_*.npz")))
    assert len(files) >= 3
    for f in files:
        g = np.load(f)
        for mode in ("f64", "host"):
            out = op.distance_postprocessing(g["border"], g["cell"], float(g["th_seed"]), float(g["th_cell"]), tan_mode=mode)
            assert np.array_equal(out, g["mask_u16"]), (f, mode)
        assert np.array_equal(op.boundary_postprocessing(g["prob"]), g["boundary_mask_u16"]), f


def test_heap_flood_equals_order_free_formulation():
    rng = np.random.default_rng(5)
    n_amb = 0
    for trial in range(120):
        H, W = int(rng.integers(3, 36)), int(rng.integers(3, 36))
        q = int(rng.choice([0, 0, 0, 4, 16]))
        v = rng.normal(size=(H, W))
        if rng.random() < 0.5:
            v = ndimage.gaussian_filter(v, rng.uniform(0.5, 3))
        if q:
            v = np.round(v / v.std() * q) / q
        v = v.astype(np.float32).astype(np.float64)
        mask = rng.random((H, W)) < rng.uniform(0.5, 1.0)
        mk = np.zeros((H, W), np.int32)
        for k in range(int(rng.integers(1, 8))):
            y, x = int(rng.integers(0, H)), int(rng.integers(0, W))
            mk[y:y + int(rng.integers(1, 3)), x:x + int(rng.integers(1, 3))] = k + 1
        a = op.watershed(v, mk, mask)
        b, amb = op.watershed_minimax(v, mk, mask)
        if amb.any():
            n_amb += 1
            assert q != 0, "tie-free input flagged ambiguous"
        else:
            assert np.array_equal(a, b), trial
        assert np.array_equal(a > 0, b > 0)
    assert n_amb > 0   # the quantised inputs must exercise the ambiguity detector


def test_flood_semantics_small_cases():
    # (iv) masked pixels not 4-connected to a marker stay 0; markers outside the mask are dropped
    v = np.zeros((3, 5))
    mask = np.array([[1, 1, 0, 1, 1]] * 3, bool)
    mk = np.zeros((3, 5), np.int32)
    mk[1, 0] = 7
    out = op.watershed(v, mk, mask)
    assert (out[:, :2] == 7).all() and (out[:, 2:] == 0).all()
    mk2 = np.zeros((3, 5), np.int32)
    mk2[1, 2] = 3
    assert (op.watershed(v, mk2, mask) == 0).all()
    # FIFO on a flat image: equidistant pixel goes to the marker whose wave front was pushed first
    v = np.zeros((1, 5))
    mk = np.array([[1, 0, 0, 0, 2]], np.int32)
    assert op.watershed(v, mk, np.ones((1, 5), bool)).tolist() == [[1, 1, 1, 2, 2]]


def test_uint16_wrap_and_empty():
    z = np.zeros((16, 16, 1), np.float32)
    assert (op.distance_postprocessing(z, z, 0.45, 0.10) == 0).all()
    inst = np.array([[65535, 65536, 65537]], np.int32)
    assert inst.astype(np.uint16).tolist() == [[65535, 0, 1]]


def test_small_seed_removal_rule():
    # two seeds: areas 30 and 3 -> mean 16.5 -> min_area = max(1.65, 4) = 4 -> the 3-px seed is dropped
    s = np.zeros((20, 20), bool)
    s[2:7, 2:8] = True
    s[15, 15:18] = True
    lab, m = op.filter_seeds(s)
    assert m == 1 and lab[15, 16] == 0 and lab[3, 3] == 1
    s[15:17, 15:18] = True  # 6 px survives (6 > 4)
    assert op.filter_seeds(s)[1] == 2


def test_tan_policy_flip_rate_is_negligible():
    m = sy.synth_instance_mask(256, 256, 120, 3)
    border, cell = sy.synth_distance_maps(m, 4)
    a = op.seed_mask_maps(border, cell, 0.45, 0.10, tan_mode="f64")[2]
    b = op.seed_mask_maps(border, cell, 0.45, 0.10, tan_mode="host")[2]
    assert (a != b).mean() < 1e-4


def test_synthetic_is_deterministic():
    a = sy.synth_instance_mask(128, 128, 40, 9)
    assert np.array_equal(a, sy.synth_instance_mask(128, 128, 40, 9))
    assert a.max() > 20
    f = sy.synth_frame(64, 96, 5)
    assert f.dtype == np.uint16 and f.shape == (64, 96) and np.array_equal(f, sy.synth_frame(64, 96, 5))


def test_c_abi_exports_every_declared_symbol(native_lib):
    hdr = open(os.path.join(ROOT, "include", "mbseg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(mbs_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 12
    for n in names:
        assert hasattr(native_lib, n), f"libmbseg.so does not export {n}"
    assert native_lib.mbs_version() >= 1
    assert native_lib.mbs_postproc_workspace_bytes(64, 64) > 64 * 64 * 40


def test_ctypes_signatures_match_the_header():
    """every prototype in include/mbseg.h has a ctypes signature in the host mirror with the same number of parameters
    (an ABI drift between header, library and mirror would otherwise corrupt the stack silently)"""
    from microbeseg_b200 import _native as nat
    hdr = open(os.path.join(ROOT, "include", "mbseg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"\b(?:int|int64_t|size_t|long long|const char \*)\s*\*?\s*(mbs_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) >= 30
    seen = set()
    for name, params in protos:
        seen.add(name)
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert name in nat._SIGS, f"no ctypes signature for {name}"
        assert len(nat._SIGS[name][1]) == n, (name, n, len(nat._SIGS[name][1]))
    assert set(nat._SIGS) <= seen, sorted(set(nat._SIGS) - seen)
