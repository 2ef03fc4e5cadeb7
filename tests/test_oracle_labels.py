"""CPU tests: label-generation oracle (oracle/labels.py) against its golden fixtures and scipy invariants."""
import glob
import os

import numpy as np
from scipy import ndimage

from oracle import labels as ol

HERE = os.path.dirname(os.path.abspath(__file__))


def test_golden_labels():
    files = sorted(glob.glob(os.path.join(HERE, "golden", "labels_*.npz")))
    assert len(files) >= 2
    for f in files:
        g = np.load(f)
        (cd, nd), mal = ol.create_labels(g["mask"])
        assert mal == int(g["max_mal"])
        assert np.array_equal(cd, g["cell_dist"]) and np.array_equal(nd, g["neighbor_dist"]), f


def test_simple_label_types_against_the_reference_functions():
    """boundary / border / j4 label images: the fixtures were produced by the reference's OWN functions
    (train_data_representations.py:75-190, imported by path in tests/golden/make_golden.py simple_labels; scipy + numpy,
    skimage's one-line disk() stubbed) -- the oracle must reproduce them exactly"""
    files = sorted(glob.glob(os.path.join(HERE, "golden", "simple_labels_*.npz")))
    assert len(files) >= 2
    for f in files:
        g = np.load(f)
        m = g["mask"]
        assert np.array_equal(ol.boundary_label(m), g["boundary"]), f
        assert np.array_equal(ol.border_label(m), g["border"]), f
        assert np.array_equal(ol.j4_label(m), g["j4"]), f


def test_distance_label_against_the_reference_function_bodies():
    """tests/golden/refbody_labels_*.npz: outputs of the reference's OWN distance_label / cell_distance_label(clipping) /
    bottom_hat_closing (train_data_representations.py:40-72, 220-361, imported by path) running on top of the restated
    regionprops / measure.label -- so everything in oracle/labels.py except those two scikit-image primitives is pinned to the
    reference's code"""
    files = sorted(glob.glob(os.path.join(HERE, "golden", "refbody_labels_*.npz")))
    assert len(files) >= 2
    for f in files:
        g = np.load(f)
        m, mal = g["mask"], int(g["max_mal"])
        assert ol.max_major_axis_length(m) == mal
        cd, nd = ol.get_label(m, "distance", mal)
        assert np.array_equal(cd, g["cell_dist"]) and np.array_equal(nd, g["neighbor_dist"]), f
        assert np.array_equal(ol.get_label(m, "cell_dist_clipped", mal), g["cell_dist_clipped"]), f
        gaps, gap_map = ol.bottom_hat_closing(m)
        assert np.array_equal(gaps, g["gaps"]) and np.array_equal(gap_map, g["gap_map"]), f


def test_restated_skimage_pieces():
    d = ol.disk(3)
    assert d.shape == (7, 7) and d.sum() == 29 and d[0, 3] == 1 and d[0, 2] == 0
    m = np.zeros((40, 60), np.uint16)
    m[10:20, 5:45] = 3                       # 10 x 40 rectangle: axis lengths 4*sqrt((n^2-1)/12)
    r = ol.regionprops(m)[0]
    assert r.label == 3 and r.area == 400 and r.centroid == (14.5, 24.5)
    assert abs(r.major_axis_length - 4 * np.sqrt((40 ** 2 - 1) / 12)) < 1e-9
    assert abs(r.minor_axis_length - 4 * np.sqrt((10 ** 2 - 1) / 12)) < 1e-9
    assert ol.max_major_axis_length(m) == int(np.ceil(r.major_axis_length))


def test_border_label_and_window_semantics():
    m = np.zeros((12, 12), np.uint16)
    m[2:8, 2:6] = 1
    m[2:8, 6:10] = 2
    b = ol.border_label(m)
    assert set(np.unique(b)) == {0, 1, 2}
    assert (b[2:8, 5] == 2).all() and (b[2:8, 6] == 2).all() and (b[2:8, 2] == 1).all()
    # window = [round(c) - R, round(c) + R) clipped; np.round is half-to-even
    wy, wx = ol._window((4.5, 3.5), 3, (12, 12))
    assert (wy.start, wy.stop, wx.start, wx.stop) == (1, 7, 1, 7)
    # scipy's EDT of an all-foreground array measures the distance to (-1, 0): relied upon by the GPU path
    e = ndimage.distance_transform_edt(np.ones((3, 4), bool))
    assert np.allclose(e[0], np.sqrt(1 + np.arange(4) ** 2)) and np.isclose(e[2, 0], 3.0)


def test_boundary_and_border_label_semantics():
    from oracle import labels as ol
    m = np.zeros((8, 10), np.uint16)
    m[2:5, 1:4] = 3
    m[2:5, 4:7] = 5               # touches nucleus 3
    m[6:8, 8:10] = 9              # isolated, at the image corner
    b = ol.boundary_label(m)
    assert b[1, 1] == 2 and b[3, 3] == 2 and b[3, 4] == 2 and b[3, 2] == 1 and b[0, 9] == 0 and b[5, 8] == 2
    r = ol.border_label(m)
    assert r[1, 1] == 0 and r[3, 3] == 2 and r[3, 4] == 2 and r[3, 2] == 1 and r[7, 9] == 1 and r[5, 8] == 0
    assert (r == 2).sum() == 6
    cd = ol.cell_distance_label(m, 6)
    assert cd.dtype == np.float32 and cd.max() == 1.0 and np.array_equal(cd, ol.distance_label(m, 6)[0])
