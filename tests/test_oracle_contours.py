"""Oracle of the mask -> polygon step (oracle/contours.py) on hand-checkable shapes.  cv2 is not installable here, so
these known answers are worked out from OpenCV's documented border following (outer border, CHAIN_APPROX_NONE: start
at the first raster pixel, the walk goes down the west side first and returns along the north side)."""
import numpy as np

from oracle import contours as oc


def _img(rows):
    return np.array([[1 if ch == "#" else 0 for ch in r] for r in rows], np.uint8)


def test_square_walks_west_side_first():
    pts = oc.find_outer_contour(_img([".....", ".###.", ".###.", ".###.", "....."]))
    assert pts.tolist() == [[1, 1], [1, 2], [1, 3], [2, 3], [3, 3], [3, 2], [3, 1], [2, 1]]


def test_thin_shapes_are_walked_forth_and_back():
    assert oc.find_outer_contour(_img(["......", ".####.", "......"])).tolist() == [[1, 1], [2, 1], [3, 1], [4, 1], [3, 1], [2, 1]]
    assert oc.find_outer_contour(_img(["...", ".#.", "..."])).tolist() == [[1, 1]]
    assert oc.find_outer_contour(_img([".....", ".#...", "..#..", "...#.", "....."])).tolist() == [[1, 1], [2, 2], [3, 3], [2, 2]]


def test_hole_is_ignored_and_offsets_are_restored():
    mask = np.zeros((12, 14), np.uint16)
    mask[3:8, 5:10] = 7
    mask[5, 7] = 0                                             # a hole: the reference ends up with the outer contour
    polys = oc.mask_to_polygons(mask)
    assert list(polys) == [7] and len(polys[7]) == 1
    p = polys[7][0]
    assert p.shape == (2, 16) and p[:, 0].tolist() == [3, 5]  # (y, x) of the first raster pixel, absolute coordinates
    assert p[0].min() == 3 and p[0].max() == 7 and p[1].min() == 5 and p[1].max() == 9
    assert oc.points_string(p).startswith("5,3 5,4 ")         # "x,y " as infer.py:281-284 writes it


def test_contour_is_closed_and_8_connected_on_random_blobs():
    from microbeseg_b200 import synthetic as sy
    m = sy.synth_instance_mask(96, 96, 30, 3)
    for i, polys in oc.mask_to_polygons(m).items():
        p = polys[0].T
        assert (m[p[:, 0], p[:, 1]] == i).all()               # only pixels of the instance
        d = np.abs(np.diff(np.vstack([p, p[:1]]), axis=0)).max(1)
        assert len(p) == 1 or (d == 1).all()                  # consecutive points (and last -> first) are 8-neighbours
