"""GPU parity tests: label generation (distance method) through the C ABI vs oracle/labels.py.

cell_dist involves only integer squared distances, IEEE sqrt and division -> bit exact.
neighbor_dist additionally goes through exp() (numpy's and CUDA's float64 exp may differ in the last
bit) -> after the float32 cast at most 1 float32 ulp (tolerance 1.2e-7 absolute on [0,1] maps)."""
import os

import numpy as np
import pytest
import torch

from oracle import labels as ol
from microbeseg_b200 import synthetic as sy

pytestmark = pytest.mark.gpu
NEIGH_TOL = 1.2e-7


@pytest.fixture(scope="module")
def lab(native_lib):
    from microbeseg_b200 import labels
    return labels


def _mask(H, W, n, seed):
    return sy.synth_instance_mask(H, W, n, seed, (9.0, 16.0), (7.0, 12.0)).astype(np.uint16)


def _check(got, ref, what=""):
    (gc, gn), (rc, rn) = got, ref
    assert gc.dtype == np.float32 and gn.dtype == np.float32 and gc.shape == rc.shape
    assert np.array_equal(gc, rc), (what, np.abs(gc - rc).max())
    assert np.abs(gn.astype(np.float64) - rn).max() <= NEIGH_TOL, (what, np.abs(gn - rn).max())


@pytest.mark.parametrize("H,W,n,seed", [(320, 320, 120, 10000), (320, 320, 60, 10001), (128, 200, 40, 10002),
                                        (64, 64, 6, 10003)])
def test_distance_label_vs_oracle(lab, H, W, n, seed):
    m = _mask(H, W, n, seed)
    mal = ol.max_major_axis_length(m)
    assert int(lab.max_major_axis_lengths(m)[0]) == mal
    ref = ol.get_label(m, 'distance', mal)
    got = lab.get_label(m, 'distance', mal)
    _check(got, ref, (H, W, n))
    assert (ref[1] > 0).mean() > 0.01          # the neighbour map is exercised


def test_touching_cells_gaps_and_borders(lab):
    # densely packed: many touching borders, bottom-hat gaps with rings, artefact gaps
    m = _mask(256, 256, 150, 7)
    (ref, im) = ol.distance_label(m, 20, return_intermediates=True)
    assert im["label_border"].sum() > 50 and im["gaps"].max() > 5
    _check(lab.distance_label(m, 20), ref, "dense")


def test_instance_window_larger_than_shared_memory(lab):
    """max_mal ~ 300 px -> search radius 225 -> a 450 x 450 window (202 500 elements) cannot live in shared memory
    (cap ~38 k): those instances go through the global-memory walk (lab_cell_big_kernel) -- same exact result."""
    H, W = 512, 512
    yy, xx = np.mgrid[0:H, 0:W]
    m = np.zeros((H, W), np.uint16)
    m[((yy - 250) / 60.0) ** 2 + ((xx - 256) / 150.0) ** 2 <= 1] = 1         # major axis ~300 px
    m[((yy - 80) / 30.0) ** 2 + ((xx - 100) / 40.0) ** 2 <= 1] = 2
    m[((yy - 420) / 35.0) ** 2 + ((xx - 380) / 60.0) ** 2 <= 1] = 3
    m[(np.abs(yy - 330) <= 12) & (np.abs(xx - 120) <= 12)] = 4
    mal = ol.max_major_axis_length(m)
    assert mal > 250 and int(lab.max_major_axis_lengths(m)[0]) == mal
    _check(lab.get_label(m, 'distance', mal), ol.get_label(m, 'distance', mal), "big window")


def test_large_and_word_crossing_instances(lab):
    """instances wider than the 64-column bit window (byte-wise closing fallback OR-ing into the bit image), instances
    straddling the 64-pixel word boundaries of the bit rows, cells at the image border, width not a multiple of 64"""
    H, W = 200, 330
    yy, xx = np.mgrid[0:H, 0:W]
    m = np.zeros((H, W), np.uint16)
    m[((yy - 100) / 45.0) ** 2 + ((xx - 120) / 75.0) ** 2 <= 1] = 3          # 150 px wide
    m[((yy - 30) / 12.0) ** 2 + ((xx - 64) / 9.0) ** 2 <= 1] = 7             # straddles the word boundary x = 64
    m[((yy - 170) / 10.0) ** 2 + ((xx - 256) / 14.0) ** 2 <= 1] = 8          # straddles x = 256
    m[((yy - 186) / 14.0) ** 2 + ((xx - 205) / 11.0) ** 2 <= 1] = 9          # touches the bottom edge
    m[90:112, 318:330] = 11                                                  # touches the right edge (last partial word)
    m[((yy - 60) / 9.0) ** 2 + ((xx - 215) / 16.0) ** 2 <= 1] = 12           # close to the big one: gaps
    m[((yy - 100) / 12.0) ** 2 + ((xx - 206) / 8.0) ** 2 <= 1] = 13          # touching the big one
    ref, im = ol.distance_label(m, 40, return_intermediates=True)
    _check(lab.distance_label(m, 40), ref, "large / word crossing")
    assert (ref[1] > 0).sum() > 10 and im["gaps"].max() >= 1


def test_edge_cases(lab):
    z = np.zeros((64, 80), np.uint16)
    _check(lab.distance_label(z, 10), ol.distance_label(z, 10), "empty")
    one = np.zeros((64, 64), np.uint16)
    one[20:40, 10:50] = 5                      # single instance, ids not starting at 1, window clips the cell
    _check(lab.distance_label(one, 8), ol.distance_label(one, 8), "single clipped")
    two = np.zeros((40, 40), np.uint16)
    two[5:20, 5:20] = 1
    two[5:20, 20:35] = 2                       # touching pair
    two[0:3, 0:40] = 3                         # instance on the image edge (closing strips the rim)
    _check(lab.distance_label(two, 12), ol.distance_label(two, 12), "touching")
    full = np.full((30, 30), 9, np.uint16)     # all-foreground window: scipy's degenerate EDT
    _check(lab.distance_label(full, 6), ol.distance_label(full, 6), "all foreground")
    split = np.zeros((64, 64), np.uint16)      # one id in two blobs: centroid between them, window misses both
    split[2:8, 2:8] = 4
    split[56:62, 56:62] = 4
    split[30:34, 30:34] = 2
    _check(lab.distance_label(split, 5), ol.distance_label(split, 5), "split id")


def test_batch_create_labels_matches_per_crop(lab):
    masks = np.stack([_mask(160, 160, 30 + 5 * i, 20000 + i) for i in range(6)])
    cells, neighs, mals = lab.create_labels(masks)
    for i in range(len(masks)):
        (rc, rn), mal = ol.create_labels(masks[i])
        assert int(mals[i]) == mal
        _check((cells[i], neighs[i]), (rc, rn), f"crop {i}")


def test_golden_fixtures(lab):
    import glob
    import os
    for f in sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "labels_*.npz"))):
        g = np.load(f)
        cells, neighs, mals = lab.create_labels(g["mask"])
        assert int(mals[0]) == int(g["max_mal"])
        _check((cells[0], neighs[0]), (g["cell_dist"], g["neighbor_dist"]), f)


def test_boundary_border_and_cell_dist_label_types():
    """the other label types of get_label (train_data_representations.py:11-37) that are built: bit-exact vs the oracle"""
    from microbeseg_b200 import labels as lab, synthetic as sy
    from oracle import labels as ol
    for seed, (H, W, n) in enumerate([(96, 128, 30), (64, 64, 12), (80, 50, 0)]):
        m = sy.synth_instance_mask(H, W, n, 50 + seed, (7.0, 12.0), (5.0, 9.0)).astype(np.uint16)
        if n:
            m[0:5, 0:6] = 900            # touches the image corner
            m[0:5, 6:11] = 901           # and a neighbour
        for lt in ("boundary", "border", "j4"):
            got, want = lab.get_label(m, lt, 0), ol.get_label(m, lt, 0)
            assert got.dtype == np.uint8 and got.shape == m.shape and np.array_equal(got, want), lt
        mal = ol.max_major_axis_length(m) if n else 1
        for lt in ("cell_dist", "cell_dist_clipped"):      # per-instance normalised EDT / EDT clipped to 5 px and scaled
            got, want = lab.get_label(m, lt, mal), ol.get_label(m, lt, mal)
            assert got.dtype == np.float32 and np.array_equal(got, want), lt
    with pytest.raises(NotImplementedError):
        lab.get_label(m, "adapted_border", 10)
    with pytest.raises(Exception):
        lab.get_label(m, "nonsense", 10)


def test_simple_label_types_against_the_reference_fixtures():
    """boundary / border / j4 through the CUDA path vs fixtures produced by the reference's own functions
    (tests/golden/simple_labels_*.npz, see make_golden.py simple_labels): bit-exact"""
    import glob
    from microbeseg_b200 import labels as lab
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "simple_labels_*.npz")))
    assert len(files) >= 2
    for f in files:
        g = np.load(f)
        for lt in ("boundary", "border", "j4"):
            assert np.array_equal(lab.get_label(g["mask"], lt, 0), g[lt]), (f, lt)


def test_distance_labels_against_the_reference_function_bodies():
    """CUDA path vs tests/golden/refbody_labels_*.npz (the reference's own distance_label / cell_distance_label bodies over
    the restated regionprops / label, see make_golden.py labels_refbody): cell maps bit-exact, neighbor map within 1 ulp"""
    import glob
    from microbeseg_b200 import labels as lab
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refbody_labels_*.npz")))
    assert len(files) >= 2
    for f in files:
        g = np.load(f)
        m, mal = g["mask"], int(g["max_mal"])
        cd, nd = lab.get_label(m, "distance", mal)
        assert np.array_equal(cd, g["cell_dist"]), f
        assert np.abs(nd.astype(np.float64) - g["neighbor_dist"]).max() <= 1.2e-7, f
        assert np.array_equal(lab.get_label(m, "cell_dist_clipped", mal), g["cell_dist_clipped"]), f
        assert int(lab.max_major_axis_lengths(m[None])[0]) == mal


def test_create_labels_host_paths_agree(native_lib):
    """host path of config 4: pipelined pinned staging into fresh arrays, into caller-provided NumPy arrays and straight
    into caller-provided PINNED tensors (no host copy) all give the bytes of the device-resident entry"""
    from microbeseg_b200 import labels as lab, synthetic as sy
    masks = np.stack([sy.synth_instance_mask(160, 192, 18 + i, 900 + i).astype(np.uint16) for i in range(9)])
    c0, n0, m0 = lab.create_labels(masks)
    dev = torch.from_numpy(masks.view(np.int16)).cuda()
    cd, nd, md = lab.create_labels_device(dev, int(masks.max()))
    assert np.array_equal(c0, cd.cpu().numpy()) and np.array_equal(n0, nd.cpu().numpy()) and np.array_equal(m0, md.cpu().numpy())
    c1, n1 = np.zeros_like(c0), np.zeros_like(n0)
    lab.create_labels(masks, out=(c1, n1))
    assert np.array_equal(c1, c0) and np.array_equal(n1, n0)
    pc, pn = torch.zeros(c0.shape).pin_memory(), torch.zeros(n0.shape).pin_memory()
    lab.create_labels(masks, out=(pc, pn))
    assert np.array_equal(pc.numpy(), c0) and np.array_equal(pn.numpy(), n0)


def test_staging_roundtrip(native_lib):
    from microbeseg_b200 import staging
    rng = np.random.default_rng(5)
    for shape, dt in (((3000, 3001), np.float32), ((7, 5), np.uint16), ((2, 4096, 4096), np.uint16)):
        a = (rng.random(shape) * 60000).astype(dt)
        tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.uint16): torch.int16}[np.dtype(dt)]
        d = torch.empty(shape, dtype=tdt, device="cuda")
        staging.upload(a.view(np.int16) if dt == np.uint16 else a, d)
        back = np.zeros(shape, dt)
        staging.download(d, back.view(np.int16) if dt == np.uint16 else back)
        assert np.array_equal(a, back)
