"""GPU probe: tcgen05 weight-gradient kernel (NHWC / MN-major operands) vs torch (fp32 on bf16-rounded operands)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from microbeseg_b200 import _native as nat
L = nat.lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)


def run(kind, N, Ho, Wo, Cm, Cn, bench=0):
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
    xf, gf = x.float().permute(0, 3, 1, 2), dz.float().permute(0, 3, 1, 2)
    if kind == 2:
        w = torch.zeros(Cn, Cm, 2, 2, device=dev, requires_grad=True)
        y = F.conv_transpose2d(xf, w, stride=2)
        y.backward(gf)
        ref = w.grad.permute(1, 2, 3, 0).reshape(Cm, 4, Cn)         # [co][q][ci]
    else:
        w = torch.zeros(Cm, Cn, 3, 3, device=dev, requires_grad=True)
        y = F.conv2d(xf, w, stride=s, padding=1)
        y.backward(gf)
        ref = w.grad.permute(0, 2, 3, 1).reshape(Cm, 9, Cn)         # [co][tap][ci]
    err = (out - ref).abs().max().item()
    res = dict(kind=kind, N=N, Ho=Ho, Wo=Wo, Cm=Cm, Cn=Cn, err=err, ref=ref.abs().max().item(), flag=L.mbs_debug_flags(1))
    if bench:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(bench):
            nat.check(L.mbs_conv_wgrad(ctypes.byref(d), nat.stream_ptr()), "wgrad")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / bench
        res["ms"] = ms
        res["tflops"] = 2.0 * N * Ho * Wo * Cm * Cn * taps / ms / 1e9
    print(res, flush=True)
    return err <= 2e-3 * max(1.0, res["ref"])


cases = [(0, 1, 8, 64, 64, 64), (0, 1, 8, 48, 64, 64), (0, 2, 32, 48, 64, 64), (0, 2, 16, 20, 128, 128), (0, 1, 8, 64, 256, 512),
         (0, 2, 20, 20, 128, 64), (0, 1, 6, 4, 64, 64), (1, 2, 16, 24, 64, 64), (1, 2, 20, 20, 128, 128), (2, 2, 8, 24, 64, 128),
         (2, 2, 20, 20, 256, 512), (0, 1, 320, 320, 64, 128)]
ok = True
for c in cases:
    try:
        ok &= run(*c)
    except Exception as e:
        print("EXC", c, repr(e)[:300], flush=True)
        ok = False
        break
print("ALL_OK", ok)
if ok:
    # the layers of one training step (batch 8 of 320^2)
    for c in [(0, 8, 320, 320, 64, 64), (0, 8, 320, 320, 64, 128), (1, 8, 160, 160, 64, 64), (0, 8, 160, 160, 128, 128),
              (0, 8, 160, 160, 128, 256), (0, 8, 80, 80, 256, 256), (0, 8, 80, 80, 256, 512), (0, 8, 40, 40, 512, 512),
              (0, 8, 40, 40, 512, 1024), (0, 8, 20, 20, 1024, 1024), (2, 8, 160, 160, 64, 128), (2, 8, 20, 20, 512, 1024)]:
        run(*c, bench=5)
