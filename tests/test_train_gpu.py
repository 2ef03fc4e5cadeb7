"""GPU parity tests for the training step (forward + backward on the CUDA path, no autograd / cuDNN).

Checker: a plain PyTorch fp32 reference of the same graph with autograd (oracle/net.py::dunet_train_loss; this is a
floating-point kernel family, so the reference is torch fp32 on the same seeded inputs).  Tolerance (stated): the
CUDA path keeps activations and activation gradients in bf16 with fp32 accumulation, so the loss must agree within
2e-2 relative and every parameter gradient at least as close to the fp32 reference as torch's own bf16 autocast of the same graph is
(relative L2 error <= max(1.5 x autocast error, 0.05), median error <= 1.25 x autocast median), gradient norms (20 %) within
5 %, cosine similarity > min(0.97, autocast's - 0.05); the transposed-conv bias gradient is exactly zero in theory (BatchNorm follows directly)
and is compared absolutely.  The seeded random labels make the gradients cancellation-heavy, i.e. a hard case."""
import os

import numpy as np
import pytest
import torch

from oracle import net as onet

pytestmark = pytest.mark.gpu


def _setup(filters, seed, n, h, w, graph=False, act="relu"):
    from microbeseg_b200.unets import build_unet
    from microbeseg_b200.training import TrainEngine
    net = build_unet("DU", act, "conv", "bn", torch.device("cuda:0"), 1, filters=list(filters))
    sd = onet.seeded_state_dict(onet.reference_layout_template("DU", filters), seed)
    net.load_state_dict(sd)
    net.train()
    rng = np.random.default_rng(seed)
    img = torch.from_numpy(rng.uniform(-1, 1, (n, 1, h, w)).astype(np.float32)).cuda()
    bl = torch.from_numpy(rng.uniform(0, 1, (n, 1, h, w)).astype(np.float32)).cuda()
    cl = torch.from_numpy(rng.uniform(0, 1, (n, 1, h, w)).astype(np.float32)).cuda()
    return net, sd, TrainEngine(net, use_graph=graph), img, bl, cl


def _reference(sd, img, bl, cl, autocast=False, act="relu"):
    params = {k: v.clone().cuda().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()
              if v.dtype.is_floating_point}
    with torch.enable_grad():      # the inference tests (like the reference, infer.py:343) switch grad off globally
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss = onet.dunet_train_loss(params, img, bl, cl, act)
        loss.backward()
    return float(loss.detach()), {k: p.grad for k, p in params.items() if p.grad is not None}


@pytest.mark.parametrize("filters,n,h,w,act", [((64, 128), 2, 32, 48, "relu"), ((64, 256), 3, 64, 64, "relu"),
                                               ((64, 1024), 2, 128, 96, "relu"), ((64, 256), 2, 64, 48, "mish"),
                                               ((64, 1024), 2, 96, 128, "mish"),
                                               # the reference's out-of-memory fallbacks (train.py:283-288): 32-channel
                                               # levels run zero-padded to 64 channels, gradients sliced back
                                               ((32, 128), 2, 32, 48, "relu"), ((32, 512), 2, 64, 64, "relu"),
                                               ((32, 256), 2, 64, 48, "mish")])
def test_loss_and_gradients_vs_torch_autograd(native_lib, filters, n, h, w, act):
    """relu = the Adam recipe, mish = the Ranger recipe (train.py:174)"""
    net, sd, eng, img, bl, cl = _setup(filters, 7, n, h, w, act=act)
    loss = float(eng.forward_backward(img, bl, cl))
    ref_loss, ref_grads = _reference(sd, img, bl, cl, act=act)
    assert native_lib.mbs_debug_flags(1) == 0
    assert abs(loss - ref_loss) <= 2e-2 * abs(ref_loss), (loss, ref_loss)
    # yardstick: the same graph under torch's own bf16 autocast (the standard mixed-precision policy)
    _, amp_grads = _reference(sd, img, bl, cl, autocast=True, act=act)
    rows = []
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        g, r, a = p.grad.float(), ref_grads[name].float(), amp_grads[name].float()
        assert g.shape == r.shape and torch.isfinite(g).all(), name
        rel = float((g - r).norm() / (r.norm() + 1e-12))
        rel_amp = float((a - r).norm() / (r.norm() + 1e-12))
        cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-20))
        cos_amp = float((a * r).sum() / (a.norm() * r.norm() + 1e-20))
        rows.append((rel, cos, name, float(r.norm()), float(g.norm()), rel_amp, cos_amp))
    if os.environ.get("MBS_PRINT_GRADS"):
        for row in sorted(rows, reverse=True)[:40]:
            print("GRAD %-42s rel=%.4f (torch autocast %.4f) cos=%.5f |ref|=%.3e |got|=%.3e" % (row[2], row[0], row[5], row[1], row[3], row[4]))
    # a bias in front of a BatchNorm without activation (transposed-conv bias) has an exactly zero gradient: BN
    # removes the mean.  Both sides then hold rounding noise only -> compare absolutely, against the weight scale.
    wscale = max(r[3] for r in rows)
    for rel, cos, name, rn, gn, rel_amp, cos_amp in rows:
        if ".up.0.bias" in name:
            assert gn < 1e-3 * wscale and rn < 1e-3 * wscale, (name, rn, gn)
            continue
        assert abs(gn - rn) <= max(0.2, 1.5 * rel_amp) * rn + 1e-6 * wscale, (name, rn, gn, rel_amp)   # magnitudes agree
        assert rel <= max(1.5 * rel_amp, 0.05), (name, rel, rel_amp, cos)         # as accurate as torch bf16 autocast
        assert cos > min(0.97, cos_amp - 0.05), (name, rel, cos, cos_amp)
    assert float(np.median([r[0] for r in rows])) <= 1.25 * float(np.median([r[5] for r in rows])) + 0.01


def test_running_statistics_and_optimizer_step(native_lib):
    from microbeseg_b200.training import train_step
    net, sd, eng, img, bl, cl = _setup((64, 128), 11, 2, 32, 32)
    opt = torch.optim.Adam(net.parameters(), lr=8e-4, betas=(0.9, 0.999), eps=1e-08, weight_decay=0, amsgrad=True)  # train.py:380-385
    rm0 = net.encoderConv[0].conv[2].running_mean.clone()
    losses = [float(train_step(eng, opt, img, bl, cl)) for _ in range(8)]
    assert losses[-1] < losses[0]                                   # the step descends on a fixed batch
    assert not torch.equal(rm0, net.encoderConv[0].conv[2].running_mean)
    assert int(net.encoderConv[0].conv[2].num_batches_tracked) == 100 + 8       # seeded state dict starts at 100
    net.eval()                                                      # and the trained weights run on the inference path
    with torch.no_grad():
        b, c = net(img)
    assert torch.isfinite(b).all() and torch.isfinite(c).all()


def test_narrow_fallback_net_trains(native_lib):
    """filters = [32, 512]-style nets (train.py:283-288): captured steps with the fused Adam; the 32-wide BatchNorm
    buffers get their running statistics back from the padded twins and the trained weights run on the inference path."""
    from microbeseg_b200.adam import Adam
    from microbeseg_b200.training import train_step
    net, sd, eng, img, bl, cl = _setup((32, 128), 5, 2, 32, 32, graph=True)
    opt = Adam(net.parameters(), lr=8e-4, betas=(0.9, 0.999), eps=1e-08, weight_decay=0, amsgrad=True)
    bn0 = net.encoderConv[0].conv[2]
    rm0 = bn0.running_mean.clone()
    losses = [float(train_step(eng, opt, img, bl, cl)) for _ in range(8)]
    assert losses[-1] < losses[0], losses
    assert bn0.running_mean.shape == (32,) and not torch.equal(rm0, bn0.running_mean)
    assert all(p.grad is not None and p.grad.shape == p.shape for p in net.parameters())
    net.eval()
    with torch.no_grad():
        b, c = net(img)
    assert torch.isfinite(b).all() and torch.isfinite(c).all()


def test_cuda_graph_replay_matches_eager(native_lib):
    """The training step is bitwise reproducible: no floating-point atomics on the gradient path (split-K weight gradients
    are stored per split and summed in split order, BatchNorm / bias / head sums go through per-block partials), so the
    captured step (CUDA graphs) and an independent eager engine produce IDENTICAL gradients.  Only the reported loss
    scalar is accumulated with atomics (last-bit differences)."""
    net, sd, eng, img, bl, cl = _setup((64, 128), 13, 2, 32, 32, graph=True)
    net2, _, eng2, _, _, _ = _setup((64, 128), 13, 2, 32, 32, graph=False)
    for it in range(3):                       # call 0 eager, call 1 captures + replays, call 2 replays
        la = float(eng.forward_backward(img, bl, cl))
        lb = float(eng2.forward_backward(img, bl, cl))
        assert abs(la - lb) <= 1e-5 * abs(lb) + 1e-7, (it, la, lb)
        for (n1, p1), (_, p2) in zip(net.named_parameters(), net2.named_parameters()):
            assert torch.equal(p1.grad, p2.grad), (it, n1, float((p1.grad - p2.grad).abs().max()))


@pytest.mark.parametrize("kind,N,Ho,Wo,Cm,Cn", [(0, 2, 32, 48, 64, 64), (0, 2, 16, 20, 128, 128), (1, 2, 20, 20, 128, 64),
                                                 (2, 2, 8, 24, 64, 128), (0, 1, 8, 64, 256, 512)])
def test_conv_wgrad_deterministic_split_k(native_lib, kind, N, Ho, Wo, Cm, Cn):
    """partial = 1: per-split tiles + mbs_wgrad_reduce == the atomics path up to fp32 summation order, and two runs are
    bitwise identical; the reduce writes the reference's parameter layout ([Cout][Cin][3][3] / [Cin][Cout][2][2])."""
    import ctypes
    from microbeseg_b200 import _native as nat
    L = native_lib
    dev = torch.device("cuda:0")
    torch.manual_seed(7 + Cm + Cn)
    s = 2 if kind == 1 else 1
    taps = 4 if kind == 2 else 9
    if kind == 2:
        dz = torch.randn(N, 2 * Ho, 2 * Wo, Cm, device=dev).bfloat16()
        x = torch.randn(N, Ho, Wo, Cn, device=dev).bfloat16()
    else:
        dz = torch.randn(N, Ho, Wo, Cm, device=dev).bfloat16()
        x = torch.randn(N, s * Ho, s * Wo, Cn, device=dev).bfloat16()
    d = nat.WgradDesc()
    d.kind, d.N, d.Ho, d.Wo = kind, N, Ho, Wo
    d.a, d.Cm, d.lda, d.coffa = dz.data_ptr(), Cm, Cm, 0
    d.b, d.Cn, d.ldb, d.coffb = x.data_ptr(), Cn, Cn, 0
    ref = torch.zeros(Cm, taps, Cn, device=dev)
    d.out, d.out_ld, d.out_coff, d.partial = ref.data_ptr(), Cn, 0, 0
    nat.check(L.mbs_conv_wgrad(ctypes.byref(d), nat.stream_ptr()), "wgrad")
    d.partial = 1
    splits = int(L.mbs_conv_wgrad_splits(ctypes.byref(d)))
    assert splits >= 1
    outs = []
    for _ in range(2):
        part = torch.full((splits, Cm, taps, Cn), float("nan"), device=dev)       # every element must be written
        d.out = part.data_ptr()
        nat.check(L.mbs_conv_wgrad(ctypes.byref(d), nat.stream_ptr()), "wgrad partial")
        out = torch.empty((Cn, Cm, 2, 2) if kind == 2 else (Cm, Cn, 3, 3), device=dev)
        nat.check(L.mbs_wgrad_reduce(part.data_ptr(), splits, Cn, None, 0, 0, Cm, 1 if kind == 2 else 0, out.data_ptr(),
                                     nat.stream_ptr()), "reduce")
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    want = ref.permute(2, 0, 1).reshape(Cn, Cm, 2, 2) if kind == 2 else ref.reshape(Cm, 3, 3, Cn).permute(0, 3, 1, 2)
    err = (outs[0] - want).abs().max().item()
    assert err <= 1e-5 * (N * Ho * Wo) ** 0.5 * 4 + 1e-5 * want.abs().max().item(), err
    assert L.mbs_debug_flags(1) == 0


WGRAD_CASES = [(0, 1, 8, 64, 64, 64), (0, 2, 32, 48, 64, 64), (0, 2, 16, 20, 128, 128), (0, 1, 8, 64, 256, 512), (0, 2, 20, 20, 128, 64),
               (0, 1, 6, 4, 64, 64), (1, 2, 16, 24, 64, 64), (1, 2, 20, 20, 128, 128), (2, 2, 8, 24, 64, 128), (2, 2, 20, 20, 256, 512),
               (0, 1, 160, 160, 64, 128)]


@pytest.mark.parametrize("kind,N,Ho,Wo,Cm,Cn", WGRAD_CASES)
def test_conv_wgrad_vs_torch(native_lib, kind, N, Ho, Wo, Cm, Cn):
    """mbs_conv_wgrad (MN-major tcgen05 GEMM straight from NHWC) vs torch autograd in float64 on the same bf16-rounded
    operands: conv3x3 stride 1 (incl. the two-taps-per-tile path for Cm == 64), stride 2, transposed conv; ragged patches."""
    import ctypes
    from microbeseg_b200 import _native as nat
    import torch.nn.functional as F
    L = native_lib
    dev = torch.device("cuda:0")
    torch.manual_seed(Ho * Wo + Cm)
    s = 2 if kind == 1 else 1
    if kind == 2:
        dz = torch.randn(N, 2 * Ho, 2 * Wo, Cm, device=dev).bfloat16()
        x = torch.randn(N, Ho, Wo, Cn, device=dev).bfloat16()
        taps = 4
    else:
        dz = torch.randn(N, Ho, Wo, Cm, device=dev).bfloat16()
        x = torch.randn(N, s * Ho, s * Wo, Cn, device=dev).bfloat16()
        taps = 9
    out = torch.zeros(Cm, taps, Cn, device=dev)
    d = nat.WgradDesc()
    d.kind, d.N, d.Ho, d.Wo = kind, N, Ho, Wo
    d.a, d.Cm, d.lda, d.coffa = dz.data_ptr(), Cm, Cm, 0
    d.b, d.Cn, d.ldb, d.coffb = x.data_ptr(), Cn, Cn, 0
    d.out, d.out_ld, d.out_coff = out.data_ptr(), Cn, 0
    nat.check(L.mbs_conv_wgrad(ctypes.byref(d), nat.stream_ptr()), "wgrad")
    torch.cuda.synchronize()
    with torch.enable_grad():
        xf, gf = x.double().permute(0, 3, 1, 2), dz.double().permute(0, 3, 1, 2)
        if kind == 2:
            w = torch.zeros(Cn, Cm, 2, 2, device=dev, dtype=torch.float64, requires_grad=True)
            F.conv_transpose2d(xf, w, stride=2).backward(gf)
            ref = w.grad.permute(1, 2, 3, 0).reshape(Cm, 4, Cn)
        else:
            w = torch.zeros(Cm, Cn, 3, 3, device=dev, dtype=torch.float64, requires_grad=True)
            F.conv2d(xf, w, stride=s, padding=1).backward(gf)
            ref = w.grad.permute(0, 2, 3, 1).reshape(Cm, 9, Cn)
    err = (out.double() - ref).abs().max().item()
    # exact bf16 products, fp32 accumulation in a different order: tolerance = a few fp32 ulps of the sum of magnitudes
    assert err <= 2e-5 * (N * Ho * Wo) ** 0.5 * 4 + 1e-4 * ref.abs().max().item(), err
    assert L.mbs_debug_flags(1) == 0


def test_ranger_mish_recipe_reduces_the_loss(native_lib):
    """the reference's default recipe (train_script.py --optimizer ranger => mish, train.py:174,399-404): fused Ranger
    step + CUDA training step, a few iterations on one batch must reduce the loss"""
    from microbeseg_b200.ranger import Ranger
    from microbeseg_b200.training import train_step
    net, sd, eng, img, bl, cl = _setup((64, 128), 3, 2, 64, 64, graph=True, act="mish")
    opt = Ranger(net.parameters(), lr=6e-3, alpha=0.5, k=6, N_sma_threshhold=5, betas=(.95, 0.999), eps=1e-6, weight_decay=0,
                 use_gc=True, gc_conv_only=False, gc_loc=True)
    losses = [float(train_step(eng, opt, img, bl, cl)) for _ in range(14)]
    assert all(np.isfinite(losses)) and losses[-1] < 0.7 * losses[0], losses
    assert opt.launches_last_step == 1
    assert native_lib.mbs_debug_flags(1) == 0


@pytest.mark.parametrize("loss", ["l1", "l2"])
def test_other_distance_criteria(native_lib, loss):
    """get_loss(loss_function, 'distance') also offers l1 / l2 (losses.py:24-29): loss value and head-bias gradients"""
    from microbeseg_b200.training import TrainEngine
    net, sd, _, img, bl, cl = _setup((64, 128), 11, 2, 32, 32)
    eng = TrainEngine(net, use_graph=False, loss=loss)
    got = float(eng.forward_backward(img, bl, cl))
    params = {k: v.clone().cuda().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()
              if v.dtype.is_floating_point}
    with torch.enable_grad():
        ref = onet.dunet_train_loss(params, img, bl, cl, "relu", loss)
        ref.backward()
    assert abs(got - float(ref.detach())) <= 2e-2 * abs(float(ref.detach())), (got, float(ref.detach()))
    for name, p in net.named_parameters():
        if name.endswith("Conv.2.bias"):            # the two 1x1 heads: gradient = sum of dloss/dpred
            r = params[name].grad
            assert torch.allclose(p.grad, r, rtol=0.1, atol=0.05 * float(r.abs().max()) + 1e-4), (name, p.grad, r)
    with pytest.raises(Exception):
        TrainEngine(net, loss="ce")                                     # boundary criteria need the 'U' net
    with pytest.raises(Exception):
        TrainEngine(net, loss="huber")


def test_boundary_criteria_kernel_vs_reference_fixtures(native_lib):
    """mbs_ce_dice_loss (loss value + dloss/dlogits) against fixtures produced by the REAL reference losses.py
    (ce_dice :71-96, nn.CrossEntropyLoss :19-20): fp32 with different summation order -> rtol 2e-5."""
    import glob
    from microbeseg_b200 import _native as nat
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ce_dice_*.npz")))
    assert len(files) >= 2
    for f in files:
        g = np.load(f)
        n, _, h, w = g["logits"].shape
        M = n * h * w
        z = torch.from_numpy(np.ascontiguousarray(g["logits"].transpose(1, 0, 2, 3))).cuda()       # planar [3][N*H*W]
        lab = torch.from_numpy(g["labels"].astype(np.uint8)).cuda()
        for kind, with_dice in (("ce_dice", 1), ("ce", 0)):
            loss = torch.zeros(1, device="cuda")
            grad = torch.empty_like(z)
            sums = torch.empty(7, dtype=torch.float64, device="cuda")
            nat.check(native_lib.mbs_ce_dice_loss(z.data_ptr(), lab.data_ptr(), M, with_dice, loss.data_ptr(), grad.data_ptr(),
                                                  sums.data_ptr(), nat.stream_ptr()), "ce_dice_loss")
            ref_l, ref_g = float(g[kind + "_loss"]), g[kind + "_grad"].transpose(1, 0, 2, 3)
            assert abs(float(loss) - ref_l) <= 2e-5 * abs(ref_l), (kind, float(loss), ref_l)
            got = grad.cpu().numpy()
            assert np.allclose(got, ref_g, rtol=2e-4, atol=2e-6 * np.abs(ref_g).max()), (kind, np.abs(got - ref_g).max())


@pytest.mark.parametrize("loss", ["ce_dice", "ce"])
def test_boundary_method_training_step(native_lib, loss):
    """'U' net with three class logits (the boundary method, train.py:468-484): loss within 2e-2 of the fp32 torch-autograd
    reference of the same graph, head gradients close, a few Adam steps reduce the loss"""
    from microbeseg_b200.adam import Adam
    from microbeseg_b200.training import TrainEngine, train_step
    from microbeseg_b200.unets import build_unet
    filters = (64, 128)
    net = build_unet("U", "relu", "conv", "bn", torch.device("cuda:0"), 1, ch_in=1, ch_out=3, filters=list(filters))
    sd = onet.seeded_state_dict(onet.reference_layout_template("U", filters, ch_out=3), 17)
    net.load_state_dict(sd)
    net.train()
    rng = np.random.default_rng(17)
    img = torch.from_numpy(rng.uniform(-1, 1, (2, 1, 32, 48)).astype(np.float32)).cuda()
    lab = torch.from_numpy(rng.integers(0, 3, (2, 32, 48)).astype(np.int64)).cuda()
    eng = TrainEngine(net, use_graph=False, loss=loss)
    got = float(eng.forward_backward(img, lab))
    params = {k: v.clone().cuda().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()
              if v.dtype.is_floating_point}
    with torch.enable_grad():
        ref = onet.unet_train_loss(params, img, lab, "relu", loss)
        ref.backward()
    assert abs(got - float(ref.detach())) <= 2e-2 * abs(float(ref.detach())), (got, float(ref.detach()))
    for name, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        if name.startswith("decoderConv.1."):                          # the 1x1 head (3 x 64 weights, 3 biases)
            r = params[name].grad
            rel = float((p.grad - r).norm() / (r.norm() + 1e-12))
            assert rel < 0.05, (name, rel)
    opt = Adam(net.parameters(), lr=8e-4, amsgrad=True)
    eng2 = TrainEngine(net, use_graph=True, loss=loss)
    with torch.enable_grad():
        losses = [float(train_step(eng2, opt, img, lab)) for _ in range(10)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    with pytest.raises(NotImplementedError):
        TrainEngine(net, loss="smooth_l1")                             # distance criteria need the DU net
