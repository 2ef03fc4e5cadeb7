"""max / mean error of every network golden (relative to scale), for A/B runs of kernel paths"""
import glob, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import net as onet
from microbeseg_b200.unets import build_unet
torch.set_grad_enabled(False)
for f in sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "net_*.npz"))):
    g = np.load(f)
    filters, act, seed = tuple(int(v) for v in g["filters"]), str(g["act"]), int(g["seed"])
    pool = str(g["pool"]) if "pool" in g.files else "conv"
    norm = str(g["norm"]) if "norm" in g.files else "bn"
    net = build_unet("DU", act, pool, norm, torch.device("cuda:0"), 1, filters=list(filters))
    net.load_state_dict(onet.seeded_state_dict(onet.reference_layout_template("DU", filters, pool_method=pool, normalization=norm), seed))
    net.eval()
    img = g["img"]; lo, hi = img.min(), img.max()
    x = torch.from_numpy((2 * (img.astype(np.float32) - lo) / (hi - lo) - 1)[None, None]).cuda()
    b, c = net(x)
    out = []
    for got, ref in ((b, g["border"]), (c, g["cell"])):
        e = np.abs(got[0, 0].cpu().numpy() - ref); s = max(1.0, float(np.abs(ref).max()))
        out.append(f"max {e.max() / s:.4f} mean {e.mean() / s:.5f}")
    print(os.path.basename(f), img.shape, " | ".join(out), flush=True)
