"""CPU tests: network oracle vs golden outputs of the REAL reference module; module surface of the
product network (state-dict key grammar, loud failure without CUDA)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import net as onet

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden_cases():
    return sorted(glob.glob(os.path.join(HERE, "golden", "net_*.npz")))


def normalise(img):
    lo, hi = img.min(), img.max()
    return 2 * (img.astype(np.float32) - lo) / (hi - lo) - 1      # infer_script_local.py:130


def test_oracle_matches_reference_goldens():
    files = _golden_cases()
    assert len(files) >= 3
    for f in files:
        g = np.load(f)
        filters, act, seed = tuple(int(v) for v in g["filters"]), str(g["act"]), int(g["seed"])
        pool = str(g["pool"]) if "pool" in g.files else "conv"
        norm = str(g["norm"]) if "norm" in g.files else "bn"
        sd = onet.seeded_state_dict(onet.reference_layout_template("DU", filters, pool_method=pool, normalization=norm), seed)
        x = torch.from_numpy(normalise(g["img"])[None, None])
        border, cell = onet.dunet_forward(sd, x, act)
        # same math, same library (CPU fp32 conv) -> agreement to rounding noise
        assert np.abs(border[0, 0].numpy() - g["border"]).max() < 2e-4 * max(1.0, np.abs(g["border"]).max()), f
        assert np.abs(cell[0, 0].numpy() - g["cell"]).max() < 2e-4 * max(1.0, np.abs(g["cell"]).max()), f


def test_state_dict_key_grammar_matches_reference_layout():
    from microbeseg_b200.unets import build_unet
    net = build_unet("DU", "relu", "conv", "bn", torch.device("cpu"), 1, filters=[64, 1024])
    sd = net.state_dict()
    tmpl = onet.reference_layout_template("DU", (64, 1024))
    assert list(sd.keys()) == list(tmpl.keys())
    assert len(sd) == 270
    assert all(tuple(sd[k].shape) == tuple(tmpl[k].shape) for k in sd)
    assert sum(v.numel() for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k) == 46374914
    u = build_unet("U", "mish", "conv", "bn", torch.device("cpu"), 1, ch_out=3, filters=[64, 256])
    assert list(u.state_dict().keys()) == list(onet.reference_layout_template("U", (64, 256), ch_out=3).keys())
    # a reference-layout state dict loads strictly
    net2 = build_unet("DU", "relu", "conv", "bn", torch.device("cpu"), 1, filters=[64, 128])
    net2.load_state_dict(onet.seeded_state_dict(onet.reference_layout_template("DU", (64, 128)), 1))


def test_no_cpu_fallback():
    from microbeseg_b200.unets import build_unet
    net = build_unet("DU", "relu", "conv", "bn", torch.device("cpu"), 1, filters=[64, 128]).eval()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 1, 32, 32))
    with pytest.raises(Exception):
        build_unet("X", "relu", "conv", "bn", torch.device("cpu"), 1)
