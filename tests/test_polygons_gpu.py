"""GPU parity: mask -> polygon contours (libmbseg, through the C ABI) vs oracle/contours.py, exact."""
import numpy as np
import pytest
import torch

from oracle import contours as oc

pytestmark = pytest.mark.gpu


def _same(got, want):
    assert sorted(got) == sorted(want)
    for i in want:
        assert len(got[i]) == 1 and np.array_equal(got[i][0], want[i][0]), i


def test_synthetic_masks_match_oracle(native_lib):
    from microbeseg_b200 import polygons as pg, synthetic as sy
    for H, W, cells, seed in [(64, 64, 10, 1), (200, 333, 150, 2), (512, 512, 700, 3)]:
        m = sy.synth_instance_mask(H, W, cells, seed).astype(np.uint16)
        _same(pg.mask_to_polygons(m), oc.mask_to_polygons(m))


def test_edge_cases(native_lib):
    from microbeseg_b200 import polygons as pg
    m = np.zeros((9, 11), np.uint16)
    assert pg.mask_to_polygons(m) == {}
    m[0, 0] = 1                      # single pixel in the corner
    m[0, 3:9] = 2                    # a line on the top border
    m[3:9, 10] = 3                   # a line on the right border
    m[4:8, 2:6] = 5                  # id 4 is absent
    m[5, 3] = 0                      # hole
    m[8, 0] = 6
    m[7, 1] = 6                      # diagonal pair in the bottom-left corner
    _same(pg.mask_to_polygons(m), oc.mask_to_polygons(m))
    got = pg.mask_to_polygons(torch.from_numpy(m.view(np.int16)).cuda())
    assert pg.points_string(got[2][0]) == oc.points_string(oc.mask_to_polygons(m)[2][0])


def test_polygons_of_a_segmented_frame(native_lib):
    """the consumer's view: segment a frame on the CUDA path, then polygons of the predicted mask == oracle's"""
    from microbeseg_b200 import polygons as pg, postprocessing as pp, synthetic as sy
    m = sy.synth_instance_mask(384, 384, 330, 9)
    border, cell = sy.synth_distance_maps(m, 10)
    pred = pp.distance_postprocessing(border, cell, 0.45, 0.10)
    _same(pg.mask_to_polygons(pred), oc.mask_to_polygons(pred))
