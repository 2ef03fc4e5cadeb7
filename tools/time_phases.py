"""Per-phase CUDA-event timing of one 2048^2 frame (device resident)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microbeseg_b200 import postprocessing as pp, synthetic as sy
from microbeseg_b200.unets import build_unet, frame_minmax
size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
torch.set_grad_enabled(False); torch.manual_seed(0)
dev = torch.device("cuda:0")
net = build_unet("DU", "relu", "conv", "bn", dev, 1, filters=[64, 1024]).eval()
img = sy.synth_frame(size, size, 2000)
d = torch.from_numpy(img.view(np.int16)).to(dev)
eng = net.engine()
out = torch.empty((size, size), dtype=torch.int16, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
acc = np.zeros(5)
R = 20
for it in range(R + 3):
    ev[0].record()
    lohi = frame_minmax(d)
    ev[1].record()
    outs = eng.run(d[None], 0, 0, 0.0, 0.0, events=[ev[1], ev[2], ev[3]], lohi_dev=lohi)
    b, c = outs[0][0, 0], outs[1][0, 0]
    pp.distance_postprocessing_device(b, c, 0.45, 0.10, out=out)
    ev[4].record()
    torch.cuda.synchronize()
    if it >= 3:
        acc += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[3].elapsed_time(ev[4]), ev[0].elapsed_time(ev[4])]
acc /= R
print("minmax %.3f  first_conv %.3f  convs %.3f  postproc(degenerate maps) %.3f  total %.3f ms" % tuple(acc))
