"""Post-processing cost on maps produced by the network with fitted heads (what bench.py's frames look like)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microbeseg_b200 import postprocessing as pp, synthetic as sy, calibrate
from microbeseg_b200.unets import build_unet
torch.set_grad_enabled(False); torch.manual_seed(0)
dev = torch.device("cuda:0")
net = build_unet("DU", "relu", "conv", "bn", dev, 1, filters=[64, 1024]).eval()
calibrate.fit_heads(net, [calibrate.synthetic_training_pair(512, 512, 7000 + 10 * k)[:3] for k in range(3)])
out = torch.empty((2048, 2048), dtype=torch.int16, device=dev)
for t in range(3):
    img = sy.synth_frame(2048, 2048, 2000 + t)
    d = torch.from_numpy(img.view(np.int16)).to(dev)
    b, c = net.forward_frame(d, [0, 0], float(img.min()), float(img.max()))
    b, c = b[0, 0], c[0, 0]
    pp.distance_postprocessing_device(b, c, 0.45, 0.10, out=out, want_info=True)
    info = dict(pp.last_info)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pp.distance_postprocessing_device(b, c, 0.45, 0.10, out=out)
    e1.record(); torch.cuda.synchronize()
    print("frame", t, "postproc %.3f ms" % (e0.elapsed_time(e1) / 10), info, "objects", int(out.max()))
