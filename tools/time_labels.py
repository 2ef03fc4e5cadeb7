"""CUDA-event timing of label generation on a device-resident batch of 500 synthetic 320x320 crops."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microbeseg_b200 import labels as lab, synthetic as sy
masks = np.stack([sy.synth_instance_mask(320, 320, 30 + (i * 7) % 90, 10000 + i, (9.0, 16.0), (7.0, 12.0)).astype(np.uint16) for i in range(50)])
masks = masks[np.arange(500) % 50]
d = torch.from_numpy(masks.view(np.int16)).cuda()
mid = int(masks.max())
hint = int(np.ceil(0.75 * int(lab.max_major_axis_lengths(masks[:50]).max())))
for _ in range(2):
    lab.create_labels_device(d, mid, hint)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    lab.create_labels_device(d, mid, hint)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("500 crops: %.3f ms -> %.2f Gpx/s" % (ms, 500 * 320 * 320 / ms / 1e6))
