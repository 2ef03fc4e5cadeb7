"""One-off robustness run: mbs_conv_gemm on random shapes (all dispatch paths: generic, paired tiles, halo, transposed
with TMA-store / direct epilogue, stride 2, dual source) vs torch on the same bf16-rounded operands."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from microbeseg_b200 import _native as nat
L = nat.lib()
dev = torch.device("cuda:0")
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
cases = int(sys.argv[2]) if len(sys.argv) > 2 else 80
for case in range(cases):
    mode = int(rng.choice([0, 0, 0, 1, 2]))
    C0 = int(rng.choice([64, 128, 256, 512]))
    C1 = int(rng.choice([0, 0, C0])) if mode == 0 else 0
    Cout = int(rng.choice([64, 128, 256])) if mode != 2 else C0 // 2 if C0 >= 128 else 64
    if mode == 2 and C0 == 64:
        C0 = 128
    N = int(rng.choice([1, 1, 2, 3]))
    big = rng.random() < 0.3
    H = int(rng.integers(4, 260 if big else 70)); W = int(rng.integers(4, 260 if big else 90))
    if mode == 1:
        H, W = 2 * max(H // 2, 2), 2 * max(W // 2, 2)
    torch.manual_seed(case)
    x0 = torch.randn(N, H, W, C0, device=dev).bfloat16()
    x1 = torch.randn(N, H, W, C1, device=dev).bfloat16() if C1 else None
    Cin = C0 + C1
    if mode == 2:
        w = torch.randn(Cin, Cout, 2, 2, device=dev) / Cin ** 0.5
        packed = torch.empty(4 * Cout, Cin, device=dev, dtype=torch.bfloat16)
        nat.check(L.mbs_pack_convT2x2_weight(w.data_ptr(), Cin, Cout, packed.data_ptr(), nat.stream_ptr()))
    else:
        w = torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5
        packed = torch.empty(Cout, 9, Cin, device=dev, dtype=torch.bfloat16)
        nat.check(L.mbs_pack_conv3x3_weight(w.data_ptr(), Cout, Cin, packed.data_ptr(), nat.stream_ptr()))
    bias, scale, shift = torch.randn(Cout, device=dev) * 0.1, torch.rand(Cout, device=dev) + 0.5, torch.randn(Cout, device=dev) * 0.1
    Ho, Wo = (H // 2, W // 2) if mode == 1 else ((2 * H, 2 * W) if mode == 2 else (H, W))
    out = torch.full((N, Ho, Wo, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
    d = nat.ConvDesc()
    d.mode, d.N, d.H, d.W = mode, N, H, W
    d.src0, d.C0, d.ld0, d.coff0 = x0.data_ptr(), C0, C0, 0
    d.src1, d.C1, d.ld1, d.coff1 = (x1.data_ptr() if C1 else None), C1, C1, 0
    d.weight, d.Cout = packed.data_ptr(), Cout
    d.bias, d.scale, d.shift, d.act = bias.data_ptr(), scale.data_ptr(), shift.data_ptr(), 1
    d.dst, d.ldd, d.coffd = out.data_ptr(), Cout, 0
    d.head_w, d.head_n, d.head_out = None, 0, None
    nat.check(L.mbs_conv_gemm(ctypes.byref(d), nat.stream_ptr()))
    torch.cuda.synchronize()
    xin = (torch.cat([x0, x1], -1) if C1 else x0).float().permute(0, 3, 1, 2)
    wf = w.bfloat16().float()
    y = (F.conv2d(xin, wf, bias, padding=1) if mode == 0 else F.conv2d(xin, wf, bias, stride=2, padding=1)
         if mode == 1 else F.conv_transpose2d(xin, wf, bias, stride=2))
    ref = (F.relu(y) * scale[None, :, None, None] + shift[None, :, None, None]).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs().max().item()
    ok = (not torch.isnan(out.float()).any().item()) and err <= 0.01 * ref.abs().max().item() + 1e-3 and L.mbs_debug_flags(1) == 0
    if not ok:
        bad += 1
        print("MISMATCH", dict(mode=mode, N=N, H=H, W=W, C0=C0, C1=C1, Cout=Cout, err=err), flush=True)
print("cases", cases, "mismatches", bad)
