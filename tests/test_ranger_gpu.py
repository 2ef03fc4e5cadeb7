"""GPU: the fused Ranger step (mbs_ranger_step via microbeseg_b200.ranger.Ranger) against fixtures produced by the
REAL reference optimizer (/root/reference/src/training/ranger2020.py, tests/golden/make_golden.py ranger)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ci,name", list(enumerate(mg.RANGER_CASES)))
def test_fused_step_matches_reference_fixture(ci, name):
    from microbeseg_b200.ranger import Ranger
    gold = np.load(os.path.join(HERE, "golden", f"ranger_{name}.npz"))
    p0, grads = mg.ranger_inputs(ci)
    dev = torch.device("cuda:0")
    params = [torch.nn.Parameter(torch.from_numpy(a.copy()).to(dev)) for a in p0]
    opt = Ranger(params, **mg.RANGER_CASES[name])
    for step in range(1, mg.RANGER_STEPS + 1):
        for p, g in zip(params, grads[step - 1]):
            p.grad = torch.from_numpy(g.copy()).to(dev)
        opt.step()
        assert opt.launches_last_step == 1                      # ONE kernel for all tensors
        if step in mg.RANGER_SNAPSHOTS:
            for i, p in enumerate(params):
                st = opt.state[p]
                # fp32 rounding only (reduction order of the centralisation mean, fused multiply-adds)
                np.testing.assert_allclose(p.detach().cpu().numpy(), gold[f"s{step}_p{i}"], rtol=2e-5, atol=2e-6)
                np.testing.assert_allclose(st["exp_avg"].cpu().numpy(), gold[f"s{step}_m{i}"], rtol=2e-5, atol=1e-7)
                np.testing.assert_allclose(st["exp_avg_sq"].cpu().numpy(), gold[f"s{step}_v{i}"], rtol=2e-5, atol=1e-8)
                np.testing.assert_allclose(st["slow_buffer"].cpu().numpy(), gold[f"s{step}_slow{i}"], rtol=2e-5, atol=2e-6)
                np.testing.assert_allclose(p.grad.cpu().numpy(), gold[f"s{step}_g{i}"], rtol=1e-5, atol=2e-7)
                assert st["step"] == step
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq", "slow_buffer"}
    assert sd["param_groups"][0]["betas"] == mg.RANGER_CASES[name].get("betas", (.95, 0.999))


def test_rejects_cpu_parameters_and_bad_arguments():
    from microbeseg_b200.ranger import Ranger
    p = torch.nn.Parameter(torch.zeros(4, 4))
    p.grad = torch.ones(4, 4)
    with pytest.raises(RuntimeError):
        Ranger([p]).step()
    with pytest.raises(ValueError):
        Ranger([p], alpha=1.5)
    with pytest.raises(ValueError):
        Ranger([p], k=0)


def test_parameters_without_gradients_are_skipped_and_new_gradient_tensors_are_picked_up():
    from microbeseg_b200.ranger import Ranger
    dev = torch.device("cuda:0")
    a = torch.nn.Parameter(torch.ones(6, 5, device=dev))
    b = torch.nn.Parameter(torch.ones(7, device=dev))
    opt = Ranger([a, b], lr=0.1)
    a.grad = torch.full((6, 5), 0.5, device=dev)
    a.grad[:, 0] = 1.0
    opt.step()
    assert torch.equal(b.detach(), torch.ones(7, device=dev)) and len(opt.state[b]) == 0
    before = a.detach().clone()
    a.grad = torch.randn(6, 5, device=dev)                      # a fresh tensor at a new address
    b.grad = torch.randn(7, device=dev)
    opt.step()
    assert not torch.equal(a.detach(), before) and opt.state[b]["step"] == 1 and opt.state[a]["step"] == 2
