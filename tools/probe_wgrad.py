"""GPU probe: tcgen05 weight-gradient kernel vs torch (fp32 on bf16-rounded operands)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from microbeseg_b200 import _native as nat
L = nat.lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)

def chw(t, shift=0, step=1):
    n, h, w, c = t.shape
    p = ((w + step - 1) // step + 7) // 8 * 8
    out = torch.empty((n, c, h, p), dtype=torch.bfloat16, device=dev)
    nat.check(L.mbs_nhwc_to_chw(t.data_ptr(), n, h, w, c, p, shift, step, out.data_ptr(), nat.stream_ptr()))
    return out

def run(kind, N, Ho, Wo, Cm, Cn):
    s = 2 if kind == 1 else 1
    if kind == 2:
        dz = torch.randn(N, 2 * Ho, 2 * Wo, Cm, device=dev).bfloat16()
        x = torch.randn(N, Ho, Wo, Cn, device=dev).bfloat16()
        taps = 4
    else:
        dz = torch.randn(N, Ho, Wo, Cm, device=dev).bfloat16()
        x = torch.randn(N, s * Ho, s * Wo, Cn, device=dev).bfloat16()
        taps = 9
    ats = {sh: chw(dz, sh, 2 if kind == 2 else 1) for sh in ((0, 1) if kind == 2 else (0,))}
    bts = {sh: chw(x, sh, s) for sh in ((0,) if kind == 2 else (-1, 0, 1))}
    torch.cuda.synchronize()
    out = torch.zeros(Cm, taps, Cn, device=dev)
    d = nat.WgradDesc()
    d.kind, d.N, d.Ho, d.Wo = kind, N, Ho, Wo
    for sh in (-1, 0, 1):
        d.At[sh + 1] = ats[sh].data_ptr() if sh in ats else None
        d.Bt[sh + 1] = bts[sh].data_ptr() if sh in bts else None
    d.Cm, d.pitchA = Cm, ats[0].shape[-1]
    d.Cn, d.pitchB = Cn, bts[0].shape[-1]
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
    print(dict(kind=kind, N=N, Ho=Ho, Wo=Wo, Cm=Cm, Cn=Cn, err=err, ref=ref.abs().max().item(), flag=L.mbs_debug_flags(1)), flush=True)

cases = [(0, 1, 8, 64, 64, 64), (0, 1, 8, 48, 64, 64), (0, 2, 32, 48, 64, 64), (0, 2, 16, 20, 128, 128), (0, 1, 8, 64, 256, 512),
         (1, 2, 16, 24, 64, 64), (2, 2, 8, 24, 64, 128), (0, 1, 320, 320, 64, 128)]
for c in cases:
    try:
        run(*c)
    except Exception as e:
        print("EXC", c, repr(e)[:300], flush=True)
        break
