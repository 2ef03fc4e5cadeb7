"""One post-processing call on network-produced maps (fitted heads) for ncu: the tail-sweep regime of flood_kernel."""
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
img = sy.synth_frame(2048, 2048, 2000)
d = torch.from_numpy(img.view(np.int16)).to(dev)
b, c = net.forward_frame(d, [0, 0], float(img.min()), float(img.max()))
b, c = b[0, 0].clone(), c[0, 0].clone()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    pp.distance_postprocessing_device(b, c, 0.45, 0.10, out=out)
torch.cuda.synchronize()
print("objects", int(out.cpu().numpy().view(np.uint16).max()))
