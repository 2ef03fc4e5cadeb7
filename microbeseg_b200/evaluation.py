"""Evaluation-side consumers of the segmentation path, on libmbseg (CUDA).

Mirrors what the reference's ``EvalWorker`` does with the two operators of this path
(/root/reference/src/evaluation/eval.py):
  * ``threshold_sweep``   -- eval.py:128-129 + :395-412: ONE network prediction per image, post-processed for every
    (th_cell, th_seed) pair of the sweep ``product([0.05, 0.075, 0.10, 0.125], [0.35, 0.45])``; the maps stay on the GPU
    and only the uint16 masks come back;
  * ``label_instances``   -- ``skimage.measure.label`` of an integer instance image (eval.py:261,313);
  * ``aji_plus``          -- ``get_fast_aji_plus`` (src/evaluation/stats_utils.py:98-179): the pairwise intersection table
    is built on the GPU from one sort of the (true id, pred id) pixel pairs instead of one full-image mask per instance;
    the maximal unique pairing is the reference's own ``scipy.optimize.linear_sum_assignment`` call on the host.
No CPU fallback: a missing CUDA device / libmbseg.so raises RuntimeError.
"""
from itertools import product

import numpy as np
import torch

from . import _native as nat
from . import postprocessing as pp

DEFAULT_TH_CELL = (0.05, 0.075, 0.10, 0.125)      # eval.py:128
DEFAULT_TH_SEED = (0.35, 0.45)


def default_thresholds():
    """eval.py:128-129: list of (th_cell, th_seed)."""
    return list(product(DEFAULT_TH_CELL, DEFAULT_TH_SEED))


def threshold_sweep(border, cell, ths=None):
    """border / cell: (H,W) or (H,W,1) float32 maps (NumPy or CUDA tensors, pads already cropped).
    Returns {(th_cell, th_seed): uint16 (H,W) NumPy mask}, each identical to
    ``distance_postprocessing(border, cell, th_seed=..., th_cell=...)`` (eval.py:401-404)."""
    import ctypes
    ths = default_thresholds() if ths is None else [tuple(t) for t in ths]
    dev = pp._device_of(border, cell)
    b = pp._as_device_map(border, dev)
    c = pp._as_device_map(cell, dev)
    if b.stride(1) != 1 or c.stride(1) != 1 or b.stride(0) != c.stride(0):
        b, c = b.contiguous(), c.contiguous()
    H, W = c.shape
    # the smoothed map, the seed labelling and the area filter depend on th_seed only: one front end per seed threshold,
    # one flood per (seed, cell) pair (mbs_distance_postprocessing_sweep)
    seeds = sorted({float(t[1]) for t in ths})
    cells = sorted({float(t[0]) for t in ths})
    L = nat.lib()
    outs = {}
    with torch.cuda.device(dev):
        stage = torch.empty((len(seeds), len(cells), H, W), dtype=torch.int16, device=dev)
        ws = pp._workspace(dev, int(L.mbs_postproc_workspace_bytes(H, W)))
        a_s = (ctypes.c_float * len(seeds))(*seeds)
        a_c = (ctypes.c_float * len(cells))(*cells)
        nat.check(L.mbs_distance_postprocessing_sweep(b.data_ptr(), c.data_ptr(), H, W, c.stride(0) if H > 1 else W,
                                                      ctypes.cast(a_s, ctypes.c_void_p), len(seeds), ctypes.cast(a_c, ctypes.c_void_p),
                                                      len(cells), stage.data_ptr(), ws.data_ptr(), ws.numel(), None, nat.stream_ptr()),
                  "distance_postprocessing_sweep")
        want = sorted({(seeds.index(float(t[1])), cells.index(float(t[0]))) for t in ths})
        if len(want) == len(seeds) * len(cells):
            host = stage.cpu().numpy().view(np.uint16)           # one read-back for the whole sweep
        else:
            host = {k: stage[k[0], k[1]].cpu().numpy().view(np.uint16) for k in want}
    for th in ths:
        k = (seeds.index(float(th[1])), cells.index(float(th[0])))
        outs[tuple(th)] = host[k[0], k[1]] if isinstance(host, np.ndarray) else host[k]
    return outs


def sweep_batch(net, img_batch, pads, ths=None):
    """EvalWorker.inference for the distance method (eval.py:378-412) without the file I/O: ``img_batch`` is the
    normalised [N,1,H,W] float tensor of the reference's InferenceDataset, ``pads`` = [pad_y, pad_x].
    Returns a list (one per image) of {(th_cell, th_seed): mask}."""
    dev = next(net.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("microbeseg_b200.evaluation needs a CUDA device (no CPU fallback)")
    with torch.no_grad():
        border, cell = net(img_batch.to(dev))
    res = []
    for h in range(border.shape[0]):
        res.append(threshold_sweep(border[h, 0, pads[0]:, pads[1]:], cell[h, 0, pads[0]:, pads[1]:], ths))
    return res


def _to_u16_device(mask, dev):
    if isinstance(mask, torch.Tensor):
        t = mask.to(dev)
        if t.dtype != torch.int16:
            t = t.to(torch.int32).to(torch.int16)       # uint16 payload
        return t.contiguous()
    arr = np.ascontiguousarray(np.asarray(mask)).astype(np.uint16)
    return torch.from_numpy(arr.view(np.int16)).to(dev)


def label_instances_device(mask):
    """-> (int32 CUDA tensor (H,W) with labels 1..n, n).  ``mask``: (H,W) instance ids < 65536."""
    L = nat.lib()
    dev = pp._device_of(mask)
    with torch.cuda.device(dev):
        m = _to_u16_device(mask, dev)
        H, W = m.shape
        nbytes = int(L.mbs_postproc_workspace_bytes(H, W))
        ws = pp._workspace(dev, nbytes)
        labels = torch.empty((H, W), dtype=torch.int32, device=dev)
        n = torch.zeros(1, dtype=torch.int32, device=dev)
        nat.check(L.mbs_label8_instances(m.data_ptr(), H, W, labels.data_ptr(), n.data_ptr(), ws.data_ptr(), ws.numel(),
                                         nat.stream_ptr()), "label8_instances")
        return labels, int(n.item())


def label_instances(mask):
    """skimage.measure.label(mask) for a 2-D integer instance image -> int32 NumPy array."""
    return label_instances_device(mask)[0].cpu().numpy()


def aji_plus(true, pred, relabel=True):
    """AJI+ of two instance masks.  ``relabel=True`` applies measure.label to both first, as eval.py:261 does
    (``get_fast_aji_plus(true=measure.label(gt), pred=measure.label(prediction))``)."""
    from scipy.optimize import linear_sum_assignment
    dev = pp._device_of(true, pred)
    with torch.cuda.device(dev):
        if relabel:
            t, nt = label_instances_device(true)
            p, npred = label_instances_device(pred)
        else:
            t = _to_u16_device(true, dev).to(torch.int32) & 0xFFFF
            p = _to_u16_device(pred, dev).to(torch.int32) & 0xFFFF
            nt, npred = int(t.max().item()), int(p.max().item())
        if nt == 0 and npred == 0:
            return float("nan")                               # 0 / 0 in the reference
        t, p = t.reshape(-1).long(), p.reshape(-1).long()
        t_area = torch.bincount(t, minlength=nt + 1)[1:].double().cpu().numpy()
        p_area = torch.bincount(p, minlength=npred + 1)[1:].double().cpu().numpy()
        both = (t > 0) & (p > 0)
        key, cnt = torch.unique((t[both] - 1) * max(npred, 1) + (p[both] - 1), return_counts=True)
        key, cnt = key.cpu().numpy(), cnt.double().cpu().numpy()
    inter = np.zeros((nt, npred), np.float64)
    union = np.zeros((nt, npred), np.float64)
    if len(key):
        ti, pi = key // max(npred, 1), key % max(npred, 1)
        inter[ti, pi] = cnt
        union[ti, pi] = t_area[ti] + p_area[pi] - cnt
    iou = inter / (union + 1.0e-6)
    pt, pq = linear_sum_assignment(-iou)
    piou = iou[pt, pq]
    pt, pq = pt[piou > 0.0], pq[piou > 0.0]
    overall_inter = inter[pt, pq].sum()
    overall_union = union[pt, pq].sum()
    unpaired_t = np.ones(nt, bool)
    unpaired_t[pt] = False
    unpaired_p = np.ones(npred, bool)
    unpaired_p[pq] = False
    overall_union += t_area[unpaired_t].sum() + p_area[unpaired_p].sum()
    return float(overall_inter / overall_union)
