import sys, glob; sys.path.insert(0, ".")
import numpy as np, torch
from oracle import net as onet
from microbeseg_b200.unets import build_unet
torch.set_grad_enabled(False)
for f in sorted(glob.glob("tests/golden/net_*.npz")):
    g = np.load(f)
    filters, act, seed = tuple(int(v) for v in g["filters"]), str(g["act"]), int(g["seed"])
    pool = str(g["pool"]) if "pool" in g.files else "conv"; norm = str(g["norm"]) if "norm" in g.files else "bn"
    net = build_unet("DU", act, pool, norm, torch.device("cuda:0"), 1, filters=list(filters))
    sd = onet.seeded_state_dict(onet.reference_layout_template("DU", filters, pool_method=pool, normalization=norm), seed)
    net.load_state_dict(sd); net.eval()
    img = g["img"]; lo, hi = img.min(), img.max()
    x = torch.from_numpy((2 * (img.astype(np.float32) - lo) / (hi - lo) - 1)[None, None]).cuda()
    b, c = net(x)
    for nm, got, ref in (("border", b, g["border"]), ("cell", c, g["cell"])):
        e = np.abs(got[0, 0].cpu().numpy() - ref); s = max(1.0, np.abs(ref).max())
        print(f.split("/")[-1][:34], nm, "max %.4f mean %.5f p99.9 %.4f (rel to %.2f)" % (e.max() / s, e.mean() / s, np.quantile(e, 0.999) / s, s))
