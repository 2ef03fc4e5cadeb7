"""One-frame driver for ncu: `python tools/profile_frame.py [size] [frames]` runs `frames` frames of the
config-2 frame path (net + post-processing) on cuda:0 with random-init weights."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microbeseg_b200 import postprocessing as pp, synthetic as sy
from microbeseg_b200.unets import build_unet
size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.set_grad_enabled(False)
torch.manual_seed(0)
dev = torch.device("cuda:0")
net = build_unet("DU", "relu", "conv", "bn", dev, 1, filters=[64, 1024]).eval()
img = sy.synth_frame(size, size, 2000)
d = torch.from_numpy(img.view(np.int16)).to(dev)
m = sy.synth_instance_mask(size, size, int(size * size * 0.4 / 330), 4096)
bm, cm = sy.synth_distance_maps(m, 4097)
bmd, cmd = torch.from_numpy(bm[..., 0]).to(dev), torch.from_numpy(cm[..., 0]).to(dev)
out = torch.empty((size, size), dtype=torch.int16, device=dev)
for _ in range(frames):
    b, c = net.forward_frame(d, [0, 0], float(img.min()), float(img.max()))
    pp.distance_postprocessing_device(bmd, cmd, 0.45, 0.10, out=out)   # realistic maps (random-init maps have no seeds)
torch.cuda.synchronize()
print("done", float(b.mean()), int(out.max()))
