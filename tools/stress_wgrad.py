"""One-off robustness run: mbs_conv_wgrad on random shapes (generic MN-major kernel, tap-pair tiles, halo kernel, stride 2,
transposed conv) vs torch autograd in float64 on the same bf16-rounded operands."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from microbeseg_b200 import _native as nat
L = nat.lib()
dev = torch.device("cuda:0")
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
cases = int(sys.argv[2]) if len(sys.argv) > 2 else 60
bad = 0
for case in range(cases):
    kind = int(rng.choice([0, 0, 0, 1, 2]))
    Cm = int(rng.choice([64, 64, 128, 256])); Cn = int(rng.choice([64, 128, 256, 512]))
    N = int(rng.choice([1, 2, 3])); Ho = int(rng.integers(2, 70)); Wo = int(rng.integers(2, 90))
    s = 2 if kind == 1 else 1
    torch.manual_seed(case)
    if kind == 2:
        dz = torch.randn(N, 2 * Ho, 2 * Wo, Cm, device=dev).bfloat16(); x = torch.randn(N, Ho, Wo, Cn, device=dev).bfloat16(); taps = 4
    else:
        dz = torch.randn(N, Ho, Wo, Cm, device=dev).bfloat16(); x = torch.randn(N, s * Ho, s * Wo, Cn, device=dev).bfloat16(); taps = 9
    out = torch.zeros(Cm, taps, Cn, device=dev)
    d = nat.WgradDesc()
    d.kind, d.N, d.Ho, d.Wo = kind, N, Ho, Wo
    d.a, d.Cm, d.lda, d.coffa = dz.data_ptr(), Cm, Cm, 0
    d.b, d.Cn, d.ldb, d.coffb = x.data_ptr(), Cn, Cn, 0
    d.out, d.out_ld, d.out_coff = out.data_ptr(), Cn, 0
    nat.check(L.mbs_conv_wgrad(ctypes.byref(d), nat.stream_ptr()), "wgrad")
    torch.cuda.synchronize()
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
    tol = 2e-5 * (N * Ho * Wo) ** 0.5 * 4 + 1e-4 * ref.abs().max().item()
    if not (err <= tol and L.mbs_debug_flags(1) == 0):
        bad += 1
        print("MISMATCH", dict(kind=kind, N=N, Ho=Ho, Wo=Wo, Cm=Cm, Cn=Cn, err=err, tol=tol), flush=True)
print("cases", cases, "mismatches", bad)
