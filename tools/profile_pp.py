"""Post-processing only (config 3 style) for ncu launch lists: python tools/profile_pp.py [size] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microbeseg_b200 import postprocessing as pp, synthetic as sy
size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
m = sy.synth_instance_mask(size, size, int(size * size * 20000 / 4096 ** 2), 4096)
b, c = sy.synth_distance_maps(m, 4097)
bd, cd = torch.from_numpy(b[..., 0]).to(dev), torch.from_numpy(c[..., 0]).to(dev)
out = torch.empty((size, size), dtype=torch.int16, device=dev)
for _ in range(reps):
    pp.distance_postprocessing_device(bd, cd, 0.45, 0.10, out=out)
torch.cuda.synchronize()
print("done", int(out.cpu().numpy().view(np.uint16).max()))
