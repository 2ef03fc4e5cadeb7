"""One-off robustness run: label generation on random crop sizes / densities vs the oracle (bit-exact cell_dist,
neighbor_dist within 1 float32 ulp), plus the boundary / border / cell_dist_clipped / j4 label types."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from microbeseg_b200 import labels as lab, synthetic as sy
from oracle import labels as ol
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
for case in range(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    H, W = int(rng.integers(24, 220)), int(rng.integers(24, 220))
    n = int(rng.integers(0, max(2, H * W // 500)))
    m = sy.synth_instance_mask(H, W, n, 5000 + case, (5.0, 14.0), (4.0, 10.0)).astype(np.uint16)
    if rng.random() < 0.3 and H > 80 and W > 100:                     # one oversized instance
        yy, xx = np.mgrid[0:H, 0:W]
        m[((yy - H // 2) / (H / 4.0)) ** 2 + ((xx - W // 2) / (W / 2.8)) ** 2 <= 1] = 60000
    mal = ol.max_major_axis_length(m) if m.max() else 1
    try:
        gc, gn = lab.get_label(m, "distance", mal)
    except RuntimeError as e:                    # documented device limit: instance window larger than shared memory
        print("device limit", case, H, W, "max_mal", mal, str(e)[:60], flush=True)
        continue
    rc, rn = ol.get_label(m, "distance", mal)
    ok = np.array_equal(gc, rc) and np.abs(gn.astype(np.float64) - rn).max() <= 1.2e-7
    ok &= np.array_equal(lab.get_label(m, "boundary", 0), ol.get_label(m, "boundary", 0))
    ok &= np.array_equal(lab.get_label(m, "border", 0), ol.get_label(m, "border", 0))
    ok &= np.array_equal(lab.get_label(m, "cell_dist_clipped", mal), ol.get_label(m, "cell_dist_clipped", mal))
    if H * W <= 12000:                            # the oracle's window count is a Python loop over the pixels
        ok &= np.array_equal(lab.get_label(m, "j4", 0), ol.get_label(m, "j4", 0))
    if not ok:
        bad += 1
        print("MISMATCH case", case, H, W, n, float(np.abs(gc - rc).max()), float(np.abs(gn - rn).max()), flush=True)
print("cases done, mismatches:", bad)
