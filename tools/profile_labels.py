"""Label generation (config 4 style) for ncu launch lists: python tools/profile_labels.py [crops] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microbeseg_b200 import labels as lab, synthetic as sy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
distinct = min(n, 50)
masks = np.stack([sy.synth_instance_mask(320, 320, 30 + (i * 7) % 90, 10000 + i, (9.0, 16.0), (7.0, 12.0)).astype(np.uint16) for i in range(distinct)])
masks = masks[np.arange(n) % distinct]
dev = torch.device("cuda:0")
d = torch.from_numpy(masks.view(np.int16)).to(dev)
max_id = int(masks.max())
hint = int(np.ceil(0.75 * int(lab.max_major_axis_lengths(masks[:distinct]).max())))
for _ in range(reps):
    cell, neigh, mal = lab.create_labels_device(d, max_id, hint)
torch.cuda.synchronize()
print("done", float(cell.max()), float(neigh.max()))
