"""oracle/losses.py (restated ce_dice / cross entropy of the boundary method) against fixtures produced by the REAL
reference losses.py (tests/golden/make_golden.py losses)."""
import glob
import os

import numpy as np
import torch

from oracle import losses as ol

HERE = os.path.dirname(os.path.abspath(__file__))


def test_restated_criteria_match_reference_fixtures():
    files = sorted(glob.glob(os.path.join(HERE, "golden", "ce_dice_*.npz")))
    assert len(files) >= 2
    for f in files:
        g = np.load(f)
        for kind in ("ce_dice", "ce"):
            z = torch.from_numpy(g["logits"]).clone().requires_grad_(True)
            loss = ol.boundary_loss(z, torch.from_numpy(g["labels"]), kind)
            loss.backward()
            assert abs(float(loss) - float(g[kind + "_loss"])) <= 1e-6 * abs(float(g[kind + "_loss"]))
            assert np.allclose(z.grad.numpy(), g[kind + "_grad"], rtol=1e-5, atol=1e-9)
