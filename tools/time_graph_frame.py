"""A/B: the config-2 frame (network + post-processing, device resident) launched eagerly vs replayed from one CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microbeseg_b200 import postprocessing as pp, synthetic as sy, calibrate
from microbeseg_b200.unets import build_unet, frame_minmax
from microbeseg_b200.utils import model_input_pads
torch.set_grad_enabled(False); torch.manual_seed(0)
dev = torch.device("cuda:0")
size = 2048
net = build_unet("DU", "relu", "conv", "bn", dev, 1, filters=[64, 1024]).eval()
calibrate.fit_heads(net, [calibrate.synthetic_training_pair(512, 512, 7000 + 10 * k)[:3] for k in range(3)])
img = sy.synth_frame(size, size, 2000)
d = torch.from_numpy(img.view(np.int16)).to(dev)
pads = model_input_pads(size, size)
out = torch.empty((size, size), dtype=torch.int16, device=dev)
lohi = torch.empty(2, dtype=torch.float32, device=dev)
scratch = torch.empty(8, dtype=torch.uint8, device=dev)


def frame():
    lh = frame_minmax(d, out=lohi, scratch=scratch)
    b, c = net.forward_frame(d, pads, lohi_dev=lh)
    pp.distance_postprocessing_device(b[0, 0, pads[0]:, pads[1]:], c[0, 0, pads[0]:, pads[1]:], 0.45, 0.10, out=out)


def timeit(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("eager  %.3f ms per frame" % timeit(frame, 50), int(out.cpu().numpy().view(np.uint16).max()), flush=True)
ref = out.clone()
side = torch.cuda.Stream(dev)
side.wait_stream(torch.cuda.current_stream())
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(side):
    frame()
    torch.cuda.synchronize()
    g.capture_begin()
    frame()
    g.capture_end()
torch.cuda.current_stream().wait_stream(side)
out.zero_()
print("graph  %.3f ms per frame" % timeit(g.replay, 50), "same mask:", bool(torch.equal(out, ref)), flush=True)
