"""CPU: the Ranger oracle (oracle/ranger.py) against fixtures produced by the REAL reference optimizer
(/root/reference/src/training/ranger2020.py via tests/golden/make_golden.py ranger)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402

from oracle.ranger import RangerOracle  # noqa: E402


@pytest.mark.parametrize("ci,name", list(enumerate(mg.RANGER_CASES)))
def test_oracle_matches_reference_fixture(ci, name):
    gold = np.load(os.path.join(HERE, "golden", f"ranger_{name}.npz"))
    p0, grads = mg.ranger_inputs(ci)
    opt = RangerOracle(p0, **mg.RANGER_CASES[name])
    for step in range(1, mg.RANGER_STEPS + 1):
        g_c = opt.step(grads[step - 1])
        if step in mg.RANGER_SNAPSHOTS:
            for i in range(len(p0)):
                # fp32 rounding only (reduction order of the centralisation mean, fused multiply-adds)
                np.testing.assert_allclose(opt.p[i], gold[f"s{step}_p{i}"], rtol=2e-5, atol=2e-6)
                np.testing.assert_allclose(opt.m[i], gold[f"s{step}_m{i}"], rtol=2e-5, atol=1e-7)
                np.testing.assert_allclose(opt.v[i], gold[f"s{step}_v{i}"], rtol=2e-5, atol=1e-8)
                np.testing.assert_allclose(opt.slow[i], gold[f"s{step}_slow{i}"], rtol=2e-5, atol=2e-6)
                np.testing.assert_allclose(g_c[i], gold[f"s{step}_g{i}"], rtol=1e-5, atol=2e-7)


def test_schedule_switches_to_adaptive_steps_after_the_threshold():
    opt = RangerOracle([np.zeros(3, np.float32)])
    flags = [opt.schedule(s)[0] > opt.thr for s in range(1, 10)]
    assert flags[:5] == [False] * 5 and flags[5:] == [True] * 4      # N_sma crosses 5 at step 6 for beta2 = 0.999


def test_dropin_module_path_exports_ranger():
    import importlib
    m = importlib.import_module("src.training.ranger2020")
    assert hasattr(m, "Ranger")
