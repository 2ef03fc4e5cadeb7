"""GPU: evaluation-side consumers (microbeseg_b200/evaluation.py) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import evalmetrics as em
from oracle import postproc as opp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def native_lib():
    from microbeseg_b200 import _native as nat
    return nat.lib()


def _masks(seed, size=160, n=60):
    from microbeseg_b200 import synthetic as sy
    return sy.synth_instance_mask(size, size, n, seed, (6.0, 12.0), (5.0, 9.0))


def test_label_instances_matches_oracle(native_lib):
    from microbeseg_b200 import evaluation as ev
    for seed in (1, 2):
        m = _masks(seed).astype(np.uint16)
        m[5:9, 5:9] = 300
        m[5:9, 9:13] = 301                      # touching instances
        m[100, 3] = 300                         # disconnected second part
        assert np.array_equal(ev.label_instances(m), em.label_instances(m))
    z = np.zeros((33, 47), np.uint16)
    assert ev.label_instances(z).max() == 0


def test_aji_plus_matches_oracle(native_lib):
    from microbeseg_b200 import evaluation as ev
    rng = np.random.default_rng(0)
    t = _masks(11)
    p = np.roll(_masks(11), (2, -1), (0, 1))
    p[rng.random(p.shape) < 0.02] = 0
    got = ev.aji_plus(t, p)
    want = em.aji_plus(em.label_instances(t), em.label_instances(p))
    assert abs(got - want) < 1e-12 and 0.2 < got < 1.0
    assert ev.aji_plus(t, t) == 1.0
    q = np.zeros_like(t)
    q[0, 0] = 1
    assert ev.aji_plus(t, q) == 0.0


def test_aji_plus_against_the_reference_function(native_lib):
    """GPU pair counting + the reference's own assignment call vs values from the reference's get_fast_aji_plus
    (tests/golden/aji_plus_reference.npz); relabel=False: the fixture masks are fed as they are, like the function itself"""
    import os
    from microbeseg_b200 import evaluation as ev
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aji_plus_reference.npz"))
    for k in range(int(g["n"])):
        t, p = g[f"true{k}"].astype(np.uint16), g[f"pred{k}"].astype(np.uint16)
        assert abs(ev.aji_plus(t, p, relabel=False) - float(g[f"aji{k}"])) < 1e-12
        assert abs(ev.aji_plus(p, t, relabel=False) - float(g[f"aji_swapped{k}"])) < 1e-12


def test_threshold_sweep_equals_single_calls_and_oracle(native_lib):
    from microbeseg_b200 import evaluation as ev, postprocessing as pp, synthetic as sy
    m = sy.synth_instance_mask(256, 256, 70, 5)
    border, cell = sy.synth_distance_maps(m, 6)
    res = ev.threshold_sweep(border, cell)
    assert list(res.keys()) == ev.default_thresholds() and len(res) == 8
    for (th_cell, th_seed), mask in res.items():
        single = pp.distance_postprocessing(border_prediction=border, cell_prediction=cell, th_seed=th_seed, th_cell=th_cell)
        assert mask.dtype == np.uint16 and np.array_equal(mask, single)
    for th in [(0.05, 0.35), (0.125, 0.45)]:
        want = opp.distance_postprocessing(border, cell, th_seed=th[1], th_cell=th[0])
        assert np.array_equal(res[th], want)
    # a stricter cell threshold can only shrink the segmented area
    areas = [int((res[(tc, 0.45)] > 0).sum()) for tc in ev.DEFAULT_TH_CELL]
    assert areas == sorted(areas, reverse=True)


def test_sweep_batch_runs_batch_through_the_network(native_lib):
    from microbeseg_b200 import evaluation as ev
    from microbeseg_b200.unets import build_unet
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    net = build_unet("DU", "relu", "conv", "bn", dev, 1, filters=[64, 128]).eval()
    x = torch.rand(3, 1, 64, 64) * 2 - 1
    out = ev.sweep_batch(net, x, [8, 0])
    assert len(out) == 3 and all(len(o) == 8 for o in out)
    assert all(v.shape == (56, 64) and v.dtype == np.uint16 for o in out for v in o.values())
