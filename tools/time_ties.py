"""Cost of the exact fallback: tie-heavy (quantised) distance maps and the boundary method at 2048^2 next to the
tie-free time of the same maps (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy import ndimage
from microbeseg_b200 import postprocessing as pp, synthetic as sy
dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048


def t_dist(b, c, reps=5):
    bd, cd = torch.from_numpy(b[..., 0]).to(dev), torch.from_numpy(c[..., 0]).to(dev)
    out = torch.empty((S, S), dtype=torch.int16, device=dev)
    pp.distance_postprocessing_device(bd, cd, 0.45, 0.10, out=out, want_info=True)
    info = dict(pp.last_info)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        pp.distance_postprocessing_device(bd, cd, 0.45, 0.10, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, info


m = sy.synth_instance_mask(S, S, int(S * S * 2600 / 1024 ** 2), 31)          # dense: many touching cells
border, cell = sy.synth_distance_maps(m, 32)
ms, info = t_dist(border, cell)
print(f"tie-free {S}^2: {ms:.3f} ms {info}", flush=True)
b0, c0 = sy.synth_distance_maps(m, 32, noise=0.0)
for q in (32, 8):
    cq = (np.round(c0 * q) / q).astype(np.float32)
    bq = (np.round(b0 * 16) / 16).astype(np.float32)
    ms, info = t_dist(bq, cq, reps=2)
    print(f"quantised to 1/{q} {S}^2: {ms:.3f} ms {info}", flush=True)
# boundary method: flat flood image; variant 1: a boundary class separates touching cells (one seed per mask component)
inner = ndimage.binary_erosion(m > 0, iterations=2)
rng = np.random.default_rng(5)
logits = rng.normal(0, 0.3, (S, S, 3)).astype(np.float32)
logits[..., 0] += np.where(m == 0, 3.0, 0.0)
logits[..., 1] += np.where(inner, 3.0, 0.0)
logits[..., 2] += np.where((m > 0) & ~inner, 2.0, 0.0)
e = np.exp(logits - logits.max(-1, keepdims=True))
prob = torch.from_numpy((e / e.sum(-1, keepdims=True)).astype(np.float32)).to(dev)
out = pp.boundary_postprocessing_device(prob, want_info=True)
info = dict(pp.last_info)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2):
    pp.boundary_postprocessing_device(prob)
e1.record(); torch.cuda.synchronize()
print(f"boundary method {S}^2: {e0.elapsed_time(e1) / 2:.3f} ms {info}, objects {int(out.cpu().numpy().view(np.uint16).max())}", flush=True)

# variant 2: touching cells share one mask component with several seeds -> equal-valued markers of different labels ->
# the heap's internal order decides -> whole-image sequential flood (the worst case, stated honestly)
seeds = np.zeros_like(m, dtype=bool)
for it in (3,):
    seeds = ndimage.binary_erosion(m > 0, iterations=it) & (ndimage.minimum_filter(np.where(m > 0, m, 1 << 30), 7) == ndimage.maximum_filter(m, 7))
logits = rng.normal(0, 0.1, (S, S, 3)).astype(np.float32)
logits[..., 0] += np.where(m == 0, 3.0, 0.0)
logits[..., 1] += np.where(m > 0, 3.0, 0.0) + np.where(seeds, 4.0, 0.0)
logits[..., 2] += np.where((m > 0) & ~seeds, 2.5, 0.0)
e = np.exp(logits - logits.max(-1, keepdims=True))
prob = torch.from_numpy((e / e.sum(-1, keepdims=True)).astype(np.float32)).to(dev)
out = pp.boundary_postprocessing_device(prob, want_info=True)
info = dict(pp.last_info)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pp.boundary_postprocessing_device(prob)
e1.record(); torch.cuda.synchronize()
print(f"boundary method, touching cells in one mask component {S}^2: {e0.elapsed_time(e1):.3f} ms {info}, objects {int(out.cpu().numpy().view(np.uint16).max())}", flush=True)
