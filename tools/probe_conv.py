"""GPU probe: tcgen05 conv kernel vs torch conv2d (fp32 on bf16-rounded operands). Writes gpurun_out/probe_conv.log"""
import ctypes, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from microbeseg_b200 import _native as nat

if os.environ.get('PROBE_PARTIAL'):
    nat._SIGS = {k: v for k, v in nat._SIGS.items() if not k.startswith(('mbs_pp', 'mbs_post', 'mbs_dist'))}
L = nat.lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)
log = []

def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    log.append(s)

def run_conv(mode, N, H, W, C0, C1, Cout, act=1, head=False, bench=0):
    x0 = torch.randn(N, H, W, C0, device=dev).bfloat16()
    x1 = torch.randn(N, H, W, C1, device=dev).bfloat16() if C1 else None
    Cin = C0 + C1
    if mode == 2:
        w = torch.randn(Cin, Cout, 2, 2, device=dev) / (Cin ** 0.5)
        packed = torch.empty(4 * Cout, Cin, device=dev, dtype=torch.bfloat16)
        nat.check(L.mbs_pack_convT2x2_weight(w.data_ptr(), Cin, Cout, packed.data_ptr(), nat.stream_ptr()))
    else:
        w = torch.randn(Cout, Cin, 3, 3, device=dev) / ((9 * Cin) ** 0.5)
        packed = torch.empty(Cout, 9, Cin, device=dev, dtype=torch.bfloat16)
        nat.check(L.mbs_pack_conv3x3_weight(w.data_ptr(), Cout, Cin, packed.data_ptr(), nat.stream_ptr()))
    bias = torch.randn(Cout, device=dev) * 0.1
    scale = torch.rand(Cout, device=dev) + 0.5
    shift = torch.randn(Cout, device=dev) * 0.1
    Ho, Wo = (H // 2, W // 2) if mode == 1 else ((2 * H, 2 * W) if mode == 2 else (H, W))
    out = torch.full((N, Ho, Wo, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
    hw = torch.randn(Cout, device=dev) * 0.1 if head else None
    hout = torch.full((N, Ho, Wo), float("nan"), device=dev) if head else None
    d = nat.ConvDesc()
    d.mode, d.N, d.H, d.W = mode, N, H, W
    d.src0, d.C0, d.ld0, d.coff0 = x0.data_ptr(), C0, C0, 0
    d.src1, d.C1, d.ld1, d.coff1 = (x1.data_ptr() if C1 else None), C1, (C1 if C1 else 0), 0
    d.weight, d.Cout = packed.data_ptr(), Cout
    d.bias, d.scale, d.shift, d.act = bias.data_ptr(), scale.data_ptr(), shift.data_ptr(), act
    d.dst, d.ldd, d.coffd = out.data_ptr(), Cout, 0
    d.head_w, d.head_n, d.head_out = (hw.data_ptr() if head else None), (1 if head else 0), (hout.data_ptr() if head else None)
    d.head_b[0] = 0.25
    nat.check(L.mbs_conv_gemm(ctypes.byref(d), nat.stream_ptr()), "conv_gemm")
    torch.cuda.synchronize()
    # reference
    xin = torch.cat([x0, x1], -1) if C1 else x0
    xin = xin.float().permute(0, 3, 1, 2)
    wf = w.bfloat16().float()
    if mode == 0:
        y = F.conv2d(xin, wf, bias, padding=1)
    elif mode == 1:
        y = F.conv2d(xin, wf, bias, stride=2, padding=1)
    else:
        y = F.conv_transpose2d(xin, wf, bias, stride=2)
    if act == 1:
        y = F.relu(y)
    y = y * scale[None, :, None, None] + shift[None, :, None, None]
    yr = y.permute(0, 2, 3, 1)
    err = (out.float() - yr).abs()
    bad = torch.isnan(out.float()).sum().item()
    tol = 0.02 * yr.abs().max().item() + 1e-3
    res = dict(mode=mode, N=N, H=H, W=W, C0=C0, C1=C1, Cout=Cout, max_err=err.max().item(), ref_max=yr.abs().max().item(),
               nan=bad, ok=bool(err.max().item() < tol and bad == 0))
    if head:
        href = (yr * hw).sum(-1) + 0.25
        res["head_err"] = (hout - href).abs().max().item()
        res["ok"] = res["ok"] and res["head_err"] < 1e-2
    res["timeout_flag"] = L.mbs_debug_flags(1)
    if bench:
        for _ in range(3):
            L.mbs_conv_gemm(ctypes.byref(d), nat.stream_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(bench):
            L.mbs_conv_gemm(ctypes.byref(d), nat.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / bench
        taps = 1 if mode == 2 else 9
        flops = 2.0 * N * (Ho * Wo if mode != 2 else H * W) * (Cout if mode != 2 else 4 * Cout) * taps * Cin
        res["ms"] = ms
        res["tflops"] = flops / ms / 1e9
    P(json.dumps(res))
    return res

cases = [
    (0, 1, 8, 16, 64, 0, 64),
    (0, 1, 32, 32, 64, 0, 64),
    (0, 1, 24, 40, 64, 0, 64),
    (0, 2, 32, 32, 128, 0, 128),
    (0, 1, 32, 32, 64, 64, 64),
    (0, 1, 32, 32, 256, 0, 256),
    (0, 1, 16, 16, 512, 512, 512),
    (1, 1, 32, 32, 64, 0, 64),
    (1, 1, 48, 80, 128, 0, 128),
    (2, 1, 16, 16, 128, 0, 64),
    (2, 1, 8, 24, 1024, 0, 512),
    (0, 1, 4, 4, 1024, 0, 1024),
    # large enough for the paired-tile (MT = 2) path, ragged edges
    (0, 1, 200, 200, 128, 0, 128),
    (1, 1, 400, 416, 128, 0, 128),
    (0, 1, 208, 200, 128, 128, 128),
    (0, 2, 168, 160, 64, 0, 128),
]
if os.environ.get("PROBE_ONLY"):       # e.g. PROBE_ONLY=2,1,1024,1024,128,0,64 : one shape (for ncu captures)
    c = tuple(int(v) for v in os.environ["PROBE_ONLY"].split(","))
    run_conv(*c, bench=int(os.environ.get("PROBE_REPS", "3")))
    sys.exit(0)
allok = True
for c in cases:
    try:
        r = run_conv(*c)
        allok &= r["ok"]
    except Exception as e:
        P("EXC", c, repr(e))
        allok = False
try:
    r = run_conv(0, 1, 32, 32, 64, 0, 64, head=True)
    allok &= r["ok"]
except Exception as e:
    P("EXC head", repr(e)); allok = False
P("ALL_OK", allok)
if allok or os.environ.get("FORCE_BENCH"):
    for c in [(0, 1, 2048, 2048, 64, 0, 64), (0, 1, 2048, 2048, 64, 64, 64), (0, 1, 1024, 1024, 128, 0, 128), (0, 1, 1024, 1024, 128, 128, 128), (0, 1, 1024, 1024, 64, 0, 128), (1, 1, 1024, 1024, 128, 0, 128),
              (0, 1, 512, 512, 256, 0, 256), (0, 1, 256, 256, 512, 0, 512), (0, 1, 256, 256, 512, 512, 512),
              (0, 1, 128, 128, 1024, 0, 1024), (1, 1, 2048, 2048, 64, 0, 64), (2, 1, 1024, 1024, 128, 0, 64),
              (2, 1, 128, 128, 1024, 0, 512)]:
        try:
            run_conv(*c, bench=10)
        except Exception as e:
            P("EXC bench", c, repr(e))
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/probe_conv.log", "w").write("\n".join(log) + "\n")
