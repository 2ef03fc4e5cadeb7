"""Oracle of the analysis table on a hand-computable mask (analysis.py:141-170 incl. the sum-of-labels quirk)."""
import numpy as np

from oracle import analysis as oa


def test_known_answers():
    m = np.zeros((8, 10), np.uint16)
    m[1:3, 1:5] = 1          # 2 x 4 rectangle
    m[4:7, 6:9] = 2          # 3 x 3 square
    r = oa.frame_statistics(m)
    assert r['counts'] == [2] and r['total_area'] == [8 * 1 + 9 * 2] and r['mean_area'] == [8.5]
    # regionprops axis lengths: 4 * sqrt(variance along the principal axes); square: var = (9-1)/12 -> 4*sqrt(2/3)
    sq = 4 * np.sqrt(2 / 3)
    rect_major, rect_minor = 4 * np.sqrt((16 - 1) / 12), 4 * np.sqrt((4 - 1) / 12)
    assert np.isclose(r['mean_major_axis_length'][0], (rect_major + sq) / 2)
    assert np.isclose(r['mean_minor_axis_length'][0], (rect_minor + sq) / 2)
