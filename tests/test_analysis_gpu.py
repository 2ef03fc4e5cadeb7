"""Per-frame analysis table: CUDA instance statistics vs the oracle's regionprops restatement.
Integer columns exact; axis lengths are float64 closed-form eigenvalues on the GPU vs LAPACK in the oracle (rtol 1e-9)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_frame_statistics_match_oracle(native_lib):
    from microbeseg_b200 import analysis as an, synthetic as sy
    from oracle import analysis as oa
    masks = np.stack([sy.synth_instance_mask(200, 240, 40 + 10 * t, 70 + t).astype(np.uint16) for t in range(3)])
    masks[2][masks[2] == 5] = 0                                   # an absent id in one frame
    got, want = an.frame_statistics(masks), oa.frame_statistics(masks)
    assert got['frame'] == want['frame']
    assert [int(v) for v in got['counts']] == [int(v) for v in want['counts']]
    assert [int(v) for v in got['total_area']] == [int(v) for v in want['total_area']]
    assert got['mean_area'] == want['mean_area']
    for k in ('mean_minor_axis_length', 'mean_major_axis_length'):
        assert np.allclose(got[k], want[k], rtol=1e-9, atol=0), k
    empty = an.frame_statistics(np.zeros((32, 32), np.uint16))
    assert empty['counts'] == [0] and np.isnan(empty['mean_area'][0])
