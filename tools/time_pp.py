"""Post-processing timings (CUDA events): config 3 (4096^2 synthetic maps) and 2048^2 network-produced maps.
MBS_PP_LEGACY=1 selects the round-1 streaming pipeline for A/B runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microbeseg_b200 import _native as nat, postprocessing as pp, synthetic as sy, calibrate
from microbeseg_b200.unets import build_unet
torch.set_grad_enabled(False); torch.manual_seed(0)
dev = torch.device("cuda:0")
L = nat.lib()


def timeit(b, c, out, reps=20):
    for _ in range(3):
        pp.distance_postprocessing_device(b, c, 0.45, 0.10, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        pp.distance_postprocessing_device(b, c, 0.45, 0.10, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def flood_profile():
    """per-sweep instrumentation written by flood_kernel into the Stats block at the head of the workspace"""
    ws = max(pp._WS.values(), key=lambda t: t.numel())
    raw = ws[:1040].cpu().numpy()
    maxclk = raw[912:1040].view(np.uint32)
    ph = raw[848:888].view(np.uint64).astype(np.float64)
    rounds, maxr, items = (raw[464 + 128 * k:592 + 128 * k].view(np.uint32) for k in range(3))
    tiles = raw[44:172].view(np.uint32)
    t = raw[176:464].view(np.uint64).astype(np.int64)
    sweeps = int(raw[32:36].view(np.uint32)[0])
    out = []
    for k in range(min(sweeps, 33)):
        out.append((int(tiles[k]) if k < 32 else -1, round(float(t[k + 1] - t[k]) / 1e3, 1),
                    f"rounds avg {rounds[k] / max(tiles[k], 1):.1f} max {maxr[k]}, items/tile {items[k] / max(tiles[k], 1):.0f}, longest visit {maxclk[k] / 1e3:.1f} kclk" if k < 32 else ""))
    fin = round(float(t[35] - t[min(sweeps, 34)]) / 1e3, 1)
    phs = ""
    if ph[4] > 0:
        phs = (f", block-wide visits in sweeps>=3: {int(ph[4] / 256) if False else int(ph[4])} visits, kclk per visit load {ph[0] / ph[4] / 1e3:.1f} "
               f"queue {ph[1] / ph[4] / 1e3:.1f} rounds {ph[2] / ph[4] / 1e3:.1f} store {ph[3] / ph[4] / 1e3:.1f}")
    return f"sweeps (tiles visited, us): {out}{phs}, final phase {fin} us, flood total {round(float(t[35] - t[0]) / 1e3, 1)} us"


tag = "legacy" if os.environ.get("MBS_PP_LEGACY") == "1" else "tiled"
for size, cells in ((4096, 20000), (2048, 5000)):
    m = sy.synth_instance_mask(size, size, cells, 4096)
    b, c = sy.synth_distance_maps(m, 4097)
    bd, cd = torch.from_numpy(b[..., 0]).to(dev), torch.from_numpy(c[..., 0]).to(dev)
    out = torch.empty((size, size), dtype=torch.int16, device=dev)
    L.mbs_launch_count(1)
    pp.distance_postprocessing_device(bd, cd, 0.45, 0.10, out=out, want_info=True)
    n_launch = int(L.mbs_launch_count(0))
    ms = timeit(bd, cd, out)
    print(f"{tag} synthetic {size}^2: {ms:.3f} ms = {size * size / 1e6 / ms:.2f} Gpx/s, launches {n_launch}, {dict(pp.last_info)}", flush=True)
    if tag == "tiled":
        print("   ", flood_profile(), flush=True)
net = build_unet("DU", "relu", "conv", "bn", dev, 1, filters=[64, 1024]).eval()
calibrate.fit_heads(net, [calibrate.synthetic_training_pair(512, 512, 7000 + 10 * k)[:3] for k in range(3)])
out = torch.empty((2048, 2048), dtype=torch.int16, device=dev)
for t in range(2):
    img = sy.synth_frame(2048, 2048, 2000 + t)
    d = torch.from_numpy(img.view(np.int16)).to(dev)
    b, c = net.forward_frame(d, [0, 0], float(img.min()), float(img.max()))
    b, c = b[0, 0].clone(), c[0, 0].clone()
    pp.distance_postprocessing_device(b, c, 0.45, 0.10, out=out, want_info=True)
    info = dict(pp.last_info)
    ms = timeit(b, c, out)
    print(f"{tag} network maps 2048^2 frame {t}: {ms:.3f} ms, {info}, objects {int(out.cpu().numpy().view(np.uint16).max())}", flush=True)
    if tag == "tiled":
        print("   ", flood_profile(), flush=True)
