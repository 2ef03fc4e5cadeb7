import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from microbeseg_b200 import synthetic as sy, _native as nat
from oracle import augment as oa
L = nat.lib(); dev = torch.device("cuda:0")
img = sy.synth_frame(96, 96, 200).astype(np.uint16)
for q in ((0.2, 99.8), (0.1, 99.9)):
    ref = oa.contrast_stretch(img[..., None], *q)[..., 0]
    t = torch.from_numpy(img.view(np.int16).copy()).to(dev)[None].contiguous()
    ws = torch.zeros(int(L.mbs_aug_workspace_bytes(1)), dtype=torch.uint8, device=dev)
    modes = torch.tensor([1], dtype=torch.int32, device=dev)
    pr = torch.tensor([[q[0], q[1], 1, 1]], dtype=torch.float32, device=dev)
    nat.check(L.mbs_aug_contrast(t.data_ptr(), 1, 96, 96, modes.data_ptr(), pr.data_ptr(), ws.data_ptr(), ws.numel(), nat.stream_ptr()))
    torch.cuda.synchronize()
    got = t[0].cpu().numpy().view(np.uint16)
    der = ws[65536 * 4:65536 * 4 + 64].cpu().numpy().view(np.float64)
    print(q, "derived", der[:5], "numpy", np.percentile(img, q), img.mean(), img.min(), img.max())
    bad = np.argwhere(got != ref)
    print("mismatches", len(bad))
    for y, x in bad[:6]:
        p0, p1 = np.percentile(img, q)
        v = (min(max(float(img[y, x]), p0), p1) - p0) / (p1 - p0) * 65535.0
        print(img[y, x], got[y, x], ref[y, x], repr(v))
