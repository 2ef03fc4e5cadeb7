"""Fused Adam / AMSGrad step vs torch.optim.Adam (the reference's optimizer object, train.py:380-385) on the same
seeded gradients: parameters and moments after 12 steps agree to fp32 rounding (rtol 2e-5; torch's multi-tensor path
fuses some multiplies differently)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("amsgrad,wd", [(True, 0.0), (False, 0.0), (True, 1e-2)])
def test_fused_adam_matches_torch(native_lib, amsgrad, wd):
    from microbeseg_b200.adam import Adam
    torch.manual_seed(0)
    shapes = [(64, 1, 3, 3), (64,), (128, 64, 3, 3), (1, 64, 1, 1), (1,), (5000,), (3, 1025)]
    pa = [torch.randn(s, device="cuda").requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = Adam(pa, lr=8e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd, amsgrad=amsgrad)
    ob = torch.optim.Adam(pb, lr=8e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd, amsgrad=amsgrad)
    for it in range(12):
        for a, b in zip(pa, pb):
            g = torch.randn_like(a) * (1.0 + it)
            a.grad, b.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    assert oa.launches_last_step == 1
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-7), float((a - b).abs().max())
        for k in ("exp_avg", "exp_avg_sq") + (("max_exp_avg_sq",) if amsgrad else ()):
            ref = ob.state[b][k]           # moments are sums with cancellation: absolute tolerance relative to their scale
            assert torch.allclose(oa.state[a][k], ref, rtol=2e-5, atol=2e-6 * float(ref.abs().max())), k
